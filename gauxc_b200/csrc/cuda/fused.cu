// K_F: the fused, persistent, warp-specialised middle of the EXC/VXC path.  Per tile of 128 points
//
//     X = B P_sub (FP64 DMMA)  ->  rho, grad rho  ->  functional, weights, EXC/N_EL  ->  Z
//
// replacing the reference device path's pack_submat + per-task cuBLAS dgemm + uvvars kernels +
// ExchCXX device call + factor/inc kernels + zmat kernel (SURVEY.md 2.1 K2-K9;
// scheme1_base.cxx:1604-1664, uvvars_lda/gga.hpp, zmat_vxc.cu) and matching the HOST semantics
// (reference_local_host_work_driver.cxx eval_xmat :123-146, eval_uvvar_{lda,gga}_rks :150-163,
// 242-268, eval_zmat_{lda,gga}_vxc_rks :586-604, 678-713; host driver :446-502).
//
// One CTA per SM walks a host-balanced list of tiles.  Five roles, 17 warps:
//   producer (1 warp) : TMA box loads of B^T (16 basis rows x 128 points) + LDGSTS gather of the
//                       matching 16 x 64 block of P through the task's AO map, 4-stage mbarrier ring
//   MMA      (8 warps): 128 x 64 chunk of X on the DMMA pipe (m8n8k4), hands each finished chunk
//                       to the density warps through shared memory
//   density  (4 warps): thread = grid point; rho += X.B, grad rho += X.dB streamed from global
//   zmat     (4 warps): thread = grid point; functional, weight scaling, EXC/N_EL tile partials,
//                       Z = 1/2 vrho B + 2 vgamma (grad rho . dB) streamed to global
// so the HBM-bound streams of the density and Z stages run underneath the DMMA work of the next
// chunk / next tile instead of in kernels of their own, and X never leaves the SM.
#include "kernels.cuh"
#include "ptx.cuh"
#include "xc_functionals.cuh"

namespace gxb {

namespace {

constexpr int FK = 16;       // basis rows (K) per pipeline stage
constexpr int FN = 64;       // columns of X per chunk
constexpr int FSTAGES = 4;
constexpr int P_LD = FN + 4;   // (ld mod 16) == 4: conflict-free DMMA B-fragment loads
constexpr int X_LD = TP + 2;   // conflict-free C-fragment stores
constexpr int MMA_WARPS = 8, DEN_WARPS = 4, Z_WARPS = 4;
constexpr int MMA_THREADS = MMA_WARPS * 32, DEN_THREADS = DEN_WARPS * 32, Z_THREADS = Z_WARPS * 32;
constexpr int FUSED_THREADS = MMA_THREADS + DEN_THREADS + Z_THREADS + 32;

struct FusedSmem {
  double A[FSTAGES][FK][TP];    // TMA destination, dense + global XOR swizzle
  double P[FSTAGES][FK][P_LD];
  double X[FN][X_LD];
  double den[4][TP];
  double red[2][2][Z_WARPS];
  uint64_t full[FSTAGES], empty[FSTAGES];
  uint64_t xfull, xempty, denfull, denempty;
};
constexpr size_t FUSED_SMEM_BYTES = sizeof(FusedSmem) + 1024;

template <bool GGA>
__global__ void __launch_bounds__(FUSED_THREADS, 1)
fused_xmat_den_zmat_kernel(const __grid_constant__ CUtensorMap tmapA, PlanView pv,
                           const DevTile* __restrict__ tiles, const int* __restrict__ order,
                           const int* __restrict__ cta_begin, double* __restrict__ ws,
                           const double* __restrict__ P, int ldp, FunctionalDesc func,
                           double* __restrict__ exc_part, double* __restrict__ nel_part,
                           int part_off) {
  extern __shared__ uint8_t smem_raw[];
  FusedSmem& S = *reinterpret_cast<FusedSmem*>(
      (reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int q_begin = cta_begin[blockIdx.x], q_end = cta_begin[blockIdx.x + 1];

  if (tid == 0) {
    for (int s = 0; s < FSTAGES; ++s) {
      mbar_init(&S.full[s], 32);            // 32 producer lanes (cp.async arrivals) + TMA bytes
      mbar_init(&S.empty[s], MMA_THREADS);
    }
    mbar_init(&S.xfull, MMA_THREADS);
    mbar_init(&S.xempty, DEN_THREADS);
    mbar_init(&S.denfull, DEN_THREADS);
    mbar_init(&S.denempty, Z_THREADS);
    mbar_fence_init();
  }
  __syncthreads();

  if (warp < MMA_WARPS) {
    // ------------------------------------------------------------------ MMA warps
    const int g = lane >> 2, t = lane & 3;
    // warps sharing an SM sub-partition (warp & 3) take complementary row blocks so that
    // ragged tiles (npts < 128) leave no sub-partition without DMMA work
    const int wn = warp >> 2;
    const int wm = wn ? 3 - (warp & 3) : (warp & 3);
    int s = 0;
    uint32_t ph = 0, xph = 0;
    int a_idx[4];
#pragma unroll
    for (int mi = 0; mi < 4; ++mi) a_idx[mi] = (wm * 32 + mi * 8 + g) ^ (t << 2);

    for (int q = q_begin; q < q_end; ++q) {
      const DevTile tile = tiles[order[q]];
      const int nbe = pv.tasks[tile.task].nbe;
      const int nk = pad16(nbe) / FK;
      const int nn = (nbe + FN - 1) / FN;
      const int mi_cnt = min(4, max(0, (tile.npts - wm * 32 + 7) / 8));
      for (int c = 0; c < nn; ++c) {
        const int ni_cnt = min(4, max(0, (nbe - c * FN - wn * 32 + 7) / 8));
        const bool active = mi_cnt > 0 && ni_cnt > 0;
        double acc[4][4][2];
#pragma unroll
        for (int mi = 0; mi < 4; ++mi)
#pragma unroll
          for (int ni = 0; ni < 4; ++ni) acc[mi][ni][0] = acc[mi][ni][1] = 0.;

        for (int ks = 0; ks < nk; ++ks) {
          mbar_wait(&S.full[s], ph);
          if (active) {
            const double* as = &S.A[s][t][0];
            const double* ps = &S.P[s][t][wn * 32 + g];
#pragma unroll
            for (int kk = 0; kk < 4; ++kk) {
              double a[4], b[4];
#pragma unroll
              for (int mi = 0; mi < 4; ++mi) a[mi] = as[kk * 4 * TP + a_idx[mi]];
#pragma unroll
              for (int ni = 0; ni < 4; ++ni) b[ni] = ps[kk * 4 * P_LD + ni * 8];
#pragma unroll
              for (int mi = 0; mi < 4; ++mi)
#pragma unroll
                for (int ni = 0; ni < 4; ++ni)
                  if (mi < mi_cnt && ni < ni_cnt) dmma(acc[mi][ni][0], acc[mi][ni][1], a[mi], b[ni]);
            }
          }
          mbar_arrive(&S.empty[s]);
          if (++s == FSTAGES) { s = 0; ph ^= 1; }
        }
        // hand the chunk of X to the density warps
        mbar_wait(&S.xempty, xph ^ 1);
#pragma unroll
        for (int mi = 0; mi < 4; ++mi)
#pragma unroll
          for (int ni = 0; ni < 4; ++ni)
#pragma unroll
            for (int j = 0; j < 2; ++j)
              S.X[wn * 32 + ni * 8 + 2 * t + j][wm * 32 + mi * 8 + g] = acc[mi][ni][j];
        mbar_arrive(&S.xfull);
        xph ^= 1;
      }
    }
  } else if (warp < MMA_WARPS + DEN_WARPS) {
    // ------------------------------------------------------------------ density warps
    const int p = tid - MMA_THREADS;
    uint32_t xph = 0, dph = 0;
    for (int q = q_begin; q < q_end; ++q) {
      const DevTile tile = tiles[order[q]];
      const int nbe = pv.tasks[tile.task].nbe;
      const size_t ms = (size_t)pad16(nbe) * TP;
      const double* __restrict__ Bt = ws + tile.ws_off;
      const int nn = (nbe + FN - 1) / FN;
      double r0 = 0., r1 = 0., r2 = 0., r3 = 0.;
      for (int c = 0; c < nn; ++c) {
        const int n0 = c * FN;
        const int ncols = min(FN, nbe - n0);
        mbar_wait(&S.xfull, xph);
        int n = 0;
        for (; n + 8 <= ncols; n += 8) {
          double x[8], b0[8], b1[8], b2[8], b3[8];
#pragma unroll
          for (int u = 0; u < 8; ++u) {
            const size_t o = (size_t)(n0 + n + u) * TP + (p ^ ((u & 3) << 2));
            b0[u] = Bt[o];
            if (GGA) { b1[u] = Bt[o + ms]; b2[u] = Bt[o + 2 * ms]; b3[u] = Bt[o + 3 * ms]; }
            x[u] = S.X[n + u][p];
          }
#pragma unroll
          for (int u = 0; u < 8; ++u) {
            r0 = fma(x[u], b0[u], r0);
            if (GGA) { r1 = fma(x[u], b1[u], r1); r2 = fma(x[u], b2[u], r2); r3 = fma(x[u], b3[u], r3); }
          }
        }
        for (; n < ncols; ++n) {
          const size_t o = (size_t)(n0 + n) * TP + (p ^ ((n & 3) << 2));
          const double x = S.X[n][p];
          r0 = fma(x, Bt[o], r0);
          if (GGA) { r1 = fma(x, Bt[o + ms], r1); r2 = fma(x, Bt[o + 2 * ms], r2); r3 = fma(x, Bt[o + 3 * ms], r3); }
        }
        mbar_arrive(&S.xempty);
        xph ^= 1;
      }
      mbar_wait(&S.denempty, dph ^ 1);
      // X carries the RKS factor 2 (eval_xmat fac = 2), the gradient another 2
      S.den[0][p] = 2. * r0;
      if (GGA) { S.den[1][p] = 4. * r1; S.den[2][p] = 4. * r2; S.den[3][p] = 4. * r3; }
      mbar_arrive(&S.denfull);
      dph ^= 1;
    }
  } else if (warp < MMA_WARPS + DEN_WARPS + Z_WARPS) {
    // ------------------------------------------------------------------ functional + Z warps
    const int p = tid - MMA_THREADS - DEN_THREADS;
    const int zw = p >> 5;
    uint32_t dph = 0;
    int it = 0;
    for (int q = q_begin; q < q_end; ++q, ++it) {
      const int tile_idx = order[q];
      const DevTile tile = tiles[tile_idx];
      const int nbe = pv.tasks[tile.task].nbe;
      const int nbp = pad16(nbe);
      const size_t ms = (size_t)nbp * TP;
      const double* __restrict__ Bt = ws + tile.ws_off;
      double* __restrict__ Z = ws + tile.ws_off + (GGA ? 4 : 1) * ms;

      mbar_wait(&S.denfull, dph);
      const double rho = S.den[0][p];
      double dx = 0., dy = 0., dz = 0.;
      if (GGA) { dx = S.den[1][p]; dy = S.den[2][p]; dz = S.den[3][p]; }
      mbar_arrive(&S.denempty);
      dph ^= 1;

      const bool ok = p < tile.npts;
      double a = 0., fx = 0., fy = 0., fz = 0., e_loc = 0., n_loc = 0.;
      if (ok) {
        const double w = pv.w[tile.pt_off + p];
        const double sigma = GGA ? dx * dx + dy * dy + dz * dz : 0.;
        const XcOut xc = eval_functional(func, rho, sigma);
        const double eps = xc.eps * w;      // host driver :453-466
        const double vrho = xc.vrho * w;
        a = 0.5 * vrho;
        if (GGA) {
          const double gf = 2. * (xc.vsigma * w);
          fx = gf * dx; fy = gf * dy; fz = gf * dz;
        }
        e_loc = eps * rho;                  // :490-497
        n_loc = w * rho;
      }
      // fixed-order tile partials of EXC / N_EL
      {
        double e = e_loc, nn = n_loc;
#pragma unroll
        for (int d = 16; d > 0; d >>= 1) {
          e += __shfl_xor_sync(0xffffffffu, e, d);
          nn += __shfl_xor_sync(0xffffffffu, nn, d);
        }
        double* red = &S.red[it & 1][0][0];
        if (lane == 0) { red[zw] = e; red[Z_WARPS + zw] = nn; }
        named_bar_sync(1, Z_THREADS);
        if (p == 0) {
          exc_part[part_off + tile_idx] = (red[0] + red[1]) + (red[2] + red[3]);
          nel_part[part_off + tile_idx] = (red[4] + red[5]) + (red[6] + red[7]);
        }
      }
      // Z rows (pad rows are never read as valid output rows)
      int mu = 0;
      for (; mu + 4 <= nbe; mu += 4) {
        double b0[4], b1[4], b2[4], b3[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          const size_t o = (size_t)(mu + u) * TP + (p ^ (u << 2));
          b0[u] = Bt[o];
          if (GGA) { b1[u] = Bt[o + ms]; b2[u] = Bt[o + 2 * ms]; b3[u] = Bt[o + 3 * ms]; }
        }
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          double z = a * b0[u];
          if (GGA) { z = fma(fx, b1[u], z); z = fma(fy, b2[u], z); z = fma(fz, b3[u], z); }
          Z[(size_t)(mu + u) * TP + (p ^ (u << 2))] = z;
        }
      }
      for (; mu < nbe; ++mu) {
        const size_t o = (size_t)mu * TP + (p ^ ((mu & 3) << 2));
        double z = a * Bt[o];
        if (GGA) { z = fma(fx, Bt[o + ms], z); z = fma(fy, Bt[o + 2 * ms], z); z = fma(fz, Bt[o + 3 * ms], z); }
        Z[o] = z;
      }
    }
  } else {
    // ------------------------------------------------------------------ producer warp
    int s = 0;
    uint32_t ph = 0;
    if (lane == 0) tma_prefetch_desc(&tmapA);
    for (int q = q_begin; q < q_end; ++q) {
      const DevTile tile = tiles[order[q]];
      const DevTask task = pv.tasks[tile.task];
      const int nbe = task.nbe;
      const int nk = pad16(nbe) / FK;
      const int nn = (nbe + FN - 1) / FN;
      const int* __restrict__ ao = pv.task_ao + task.ao_off;
      const int rowB = (int)(tile.ws_off / TP);
      for (int c = 0; c < nn; ++c) {
        const int na = c * FN + 2 * lane;
        const bool va = na < nbe, vb = na + 1 < nbe;
        const int ca = va ? __ldg(ao + na) : 0, cb = vb ? __ldg(ao + na + 1) : 0;
        for (int ks = 0; ks < nk; ++ks) {
          const int k0 = ks * FK;
          const int kmine = k0 + (lane & 15);
          const long long rb_mine = kmine < nbe ? (long long)__ldg(ao + kmine) * ldp : -1;
          mbar_wait(&S.empty[s], ph ^ 1);
          if (lane == 0) {
            mbar_expect_tx(&S.full[s], FK * TP * sizeof(double));
            tma_load_2d(&S.A[s][0][0], &tmapA, &S.full[s], 0, rowB + k0);
          }
#pragma unroll
          for (int r = 0; r < FK; ++r) {
            const long long rb = __shfl_sync(0xffffffffu, rb_mine, r);
            const bool vr = rb >= 0;
            const double* src = P + (vr ? rb : 0);
            cp_async8_zfill(&S.P[s][r][2 * lane], src + ca, vr && va);
            cp_async8_zfill(&S.P[s][r][2 * lane + 1], src + cb, vr && vb);
          }
          cp_async_mbar_arrive_noinc(&S.full[s]);
          if (++s == FSTAGES) { s = 0; ph ^= 1; }
        }
      }
    }
  }
}

}  // namespace

int fused_threads() { return FUSED_THREADS; }

void launch_fused(const CUtensorMap& tmapA, const PlanView& pv, const DevTile* tiles, const int* order,
                  const int* cta_begin, int ncta, double* ws, const double* P, int ldp,
                  FunctionalDesc func, double* exc_part, double* nel_part, int part_off,
                  cudaStream_t s) {
  if (ncta <= 0) return;
  static bool attr_set = false;
  if (!attr_set) {
    cudaFuncSetAttribute(fused_xmat_den_zmat_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                         (int)FUSED_SMEM_BYTES);
    cudaFuncSetAttribute(fused_xmat_den_zmat_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                         (int)FUSED_SMEM_BYTES);
    attr_set = true;
  }
  if (func.is_gga)
    fused_xmat_den_zmat_kernel<true><<<ncta, FUSED_THREADS, FUSED_SMEM_BYTES, s>>>(
        tmapA, pv, tiles, order, cta_begin, ws, P, ldp, func, exc_part, nel_part, part_off);
  else
    fused_xmat_den_zmat_kernel<false><<<ncta, FUSED_THREADS, FUSED_SMEM_BYTES, s>>>(
        tmapA, pv, tiles, order, cta_begin, ws, P, ldp, func, exc_part, nel_part, part_off);
}

}  // namespace gxb
