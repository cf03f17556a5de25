// K_B / K_C / K_D: the dense FP64 contractions of the EXC/VXC path on the DMMA tensor pipe
// (mma.sync m8n8k4 f64 -> SASS DMMA.8x8x4; tcgen05 has no f64 kind) fed by cp.async
// multi-stage shared-memory pipelines, plus the fused pointwise stages between them.
//
// Host semantics being matched (reference_local_host_work_driver.cxx):
//   eval_xmat            :123-146   X = 2 * P_sub * B
//   eval_uvvar_lda/gga   :150-163, 242-268   rho = B.X ; drho = 2 dB.X ; gamma = |drho|^2
//   eval_zmat_lda/gga    :586-604, 678-713   Z = 1/2 vrho B + 2 vgamma (drho . dB)
//   inc_vxc              :1678-1692  VXC_sub(lower) = B Z^T + Z B^T, scatter-add via cut map
// and the reference device equivalents K2-K11 of SURVEY.md 2.1 (pack_submat.cu, cuBLAS
// dgemm/dsyr2k per task, uvvars*.hpp, zmat_vxc.cu, cuda_inc_potential.cu) which this
// replaces with three grouped launches per batch of tiles.
#include "kernels.cuh"
#include "xc_functionals.cuh"

namespace gxb {

namespace {

__device__ __forceinline__ void cp_async16(void* smem, const void* g, bool pred) {
  const unsigned s = (unsigned)__cvta_generic_to_shared(smem);
  const int sz = pred ? 16 : 0;
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;\n" ::"r"(s), "l"(g), "r"(sz));
}
__device__ __forceinline__ void cp_async8(void* smem, const void* g, bool pred) {
  const unsigned s = (unsigned)__cvta_generic_to_shared(smem);
  const int sz = pred ? 8 : 0;
  asm volatile("cp.async.ca.shared.global [%0], [%1], 8, %2;\n" ::"r"(s), "l"(g), "r"(sz));
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::); }
template <int N>
__device__ __forceinline__ void cp_async_wait() {
  asm volatile("cp.async.wait_group %0;\n" ::"n"(N));
}

// D(8x8) += A(8x4, row) * B(4x8, col); lane = 4*g + t holds A[g][t], B[t][g], C[g][2t..2t+1]
__device__ __forceinline__ void dmma(double& c0, double& c1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
               : "+d"(c0), "+d"(c1)
               : "d"(a), "d"(b));
}

// ---------------------------------------------------------------------------------------
// K_B: per tile, X = B(TP x nbe) * P_sub(nbe x nbe) in 64-column chunks; each chunk of X is
// consumed in registers for rho += X.B and drho += X.dB, so X never touches memory.
// CTA = 256 threads (8 warps as 4(M) x 2(N), warp tile 32x32), CTA tile 128 x 64, k-step 16.
// P_sub is gathered on the fly from the full P through the task's AO map (8-byte cp.async),
// replacing the reference's separate pack kernel + P_sub scratch.
// ---------------------------------------------------------------------------------------
constexpr int XB_N = 64, XB_K = 16, XB_STAGES = 3;
constexpr int XA_LD = TP + 4;    // 132: (ld mod 16) == 4 -> conflict-free DMMA fragment loads
constexpr int XP_LD = XB_N + 4;  // 68
constexpr int XB_SMEM = XB_STAGES * XB_K * (XA_LD + XP_LD) * 8;

template <bool GGA>
__global__ void __launch_bounds__(256) xmat_density_kernel(PlanView pv,
                                                            const DevTile* __restrict__ tiles,
                                                            double* __restrict__ ws,
                                                            const double* __restrict__ P, int ldp,
                                                            double* __restrict__ den, int ntiles) {
  extern __shared__ __align__(16) double smem[];
  double* As = smem;                               // [STAGES][XB_K][XA_LD]
  double* Ps = smem + XB_STAGES * XB_K * XA_LD;    // [STAGES][XB_K][XP_LD]

  const DevTile tile = tiles[blockIdx.x];
  const DevTask task = pv.tasks[tile.task];
  const int nbe = task.nbe;
  const int* __restrict__ ao = pv.task_ao + task.ao_off;
  const double* __restrict__ Bm = ws + tile.ws_off;
  const size_t ms = (size_t)nbe * TP;

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int g = lane >> 2, t = lane & 3;
  const int wm = warp & 3, wn = warp >> 2;

  const int nk = (nbe + XB_K - 1) / XB_K;
  const int nn = (nbe + XB_N - 1) / XB_N;
  const int total = nk * nn;
  // rows of this warp that hold real points (multiples of 8)
  const int mi_cnt = min(4, max(0, (tile.npts - wm * 32 + 7) / 8));

  auto load_stage = [&](int it, int buf) {
    const int n0 = (it / nk) * XB_N, k0 = (it % nk) * XB_K;
    double* as = As + buf * XB_K * XA_LD;
    double* ps = Ps + buf * XB_K * XP_LD;
#pragma unroll
    for (int r = 0; r < 4; ++r) {
      const int idx = tid + r * 256;
      const int row = idx >> 6, ch = idx & 63;
      const int k = k0 + row;
      const bool p = k < nbe;
      cp_async16(as + row * XA_LD + ch * 2, Bm + (size_t)(p ? k : 0) * TP + ch * 2, p);
    }
#pragma unroll
    for (int r = 0; r < 4; ++r) {
      const int idx = tid + r * 256;
      const int row = idx >> 6, col = idx & 63;
      const int k = k0 + row, n = n0 + col;
      const bool p = (k < nbe) && (n < nbe);
      const size_t off = p ? ((size_t)__ldg(ao + k) * ldp + __ldg(ao + n)) : 0;
      cp_async8(ps + row * XP_LD + col, P + off, p);
    }
  };

  double acc[4][4][2];
  double r0[4], r1[4], r2[4], r3[4];
#pragma unroll
  for (int mi = 0; mi < 4; ++mi) {
    r0[mi] = r1[mi] = r2[mi] = r3[mi] = 0.;
#pragma unroll
    for (int ni = 0; ni < 4; ++ni) acc[mi][ni][0] = acc[mi][ni][1] = 0.;
  }

  for (int s = 0; s < XB_STAGES - 1; ++s) {
    if (s < total) load_stage(s, s);
    cp_async_commit();
  }

  for (int it = 0; it < total; ++it) {
    cp_async_wait<XB_STAGES - 2>();
    __syncthreads();
    {
      const int nx = it + XB_STAGES - 1;
      if (nx < total) load_stage(nx, nx % XB_STAGES);
      cp_async_commit();
    }
    const int buf = it % XB_STAGES;
    const double* as = As + buf * XB_K * XA_LD + wm * 32 + g;
    const double* ps = Ps + buf * XB_K * XP_LD + wn * 32 + g;
    const int n0 = (it / nk) * XB_N;
    const int ni_cnt = min(4, max(0, (nbe - n0 - wn * 32 + 7) / 8));
    if (mi_cnt > 0 && ni_cnt > 0) {
#pragma unroll
      for (int kk = 0; kk < 4; ++kk) {
        double a[4], b[4];
#pragma unroll
        for (int mi = 0; mi < 4; ++mi) a[mi] = as[(kk * 4 + t) * XA_LD + mi * 8];
#pragma unroll
        for (int ni = 0; ni < 4; ++ni) b[ni] = ps[(kk * 4 + t) * XP_LD + ni * 8];
#pragma unroll
        for (int mi = 0; mi < 4; ++mi)
#pragma unroll
          for (int ni = 0; ni < 4; ++ni)
            if (mi < mi_cnt && ni < ni_cnt) dmma(acc[mi][ni][0], acc[mi][ni][1], a[mi], b[ni]);
      }
    }
    if ((it % nk) == nk - 1) {
      // chunk of X complete: fold into rho / drho and reset
#pragma unroll
      for (int mi = 0; mi < 4; ++mi)
#pragma unroll
        for (int ni = 0; ni < 4; ++ni) {
          const int m = wm * 32 + mi * 8 + g;
          const int n = n0 + wn * 32 + ni * 8 + 2 * t;
#pragma unroll
          for (int j = 0; j < 2; ++j) {
            const double xv = acc[mi][ni][j];
            acc[mi][ni][j] = 0.;
            if (mi < mi_cnt && n + j < nbe) {
              const size_t o = (size_t)(n + j) * TP + m;
              r0[mi] += xv * Bm[o];
              if (GGA) {
                r1[mi] += xv * Bm[o + ms];
                r2[mi] += xv * Bm[o + 2 * ms];
                r3[mi] += xv * Bm[o + 3 * ms];
              }
            }
          }
        }
    }
  }
  cp_async_wait<0>();
  __syncthreads();

  // reduce over the 4 lanes of a quad (columns), then over the two N-warps via smem
  double* red = smem;  // [2][4][TP]
#pragma unroll
  for (int mi = 0; mi < 4; ++mi) {
    double v0 = r0[mi], v1 = r1[mi], v2 = r2[mi], v3 = r3[mi];
    v0 += __shfl_xor_sync(0xffffffffu, v0, 1);
    v0 += __shfl_xor_sync(0xffffffffu, v0, 2);
    if (GGA) {
      v1 += __shfl_xor_sync(0xffffffffu, v1, 1);
      v1 += __shfl_xor_sync(0xffffffffu, v1, 2);
      v2 += __shfl_xor_sync(0xffffffffu, v2, 1);
      v2 += __shfl_xor_sync(0xffffffffu, v2, 2);
      v3 += __shfl_xor_sync(0xffffffffu, v3, 1);
      v3 += __shfl_xor_sync(0xffffffffu, v3, 2);
    }
    if (t == 0) {
      const int m = wm * 32 + mi * 8 + g;
      red[(wn * 4 + 0) * TP + m] = v0;
      if (GGA) {
        red[(wn * 4 + 1) * TP + m] = v1;
        red[(wn * 4 + 2) * TP + m] = v2;
        red[(wn * 4 + 3) * TP + m] = v3;
      }
    }
  }
  __syncthreads();
  if (tid < TP) {
    const size_t o = (size_t)blockIdx.x * TP + tid;
    const size_t ds = (size_t)ntiles * TP;
    // X carries the RKS factor 2 (eval_xmat fac = 2), the gradient another 2
    den[o] = 2. * (red[tid] + red[4 * TP + tid]);
    if (GGA) {
      den[o + ds] = 4. * (red[1 * TP + tid] + red[5 * TP + tid]);
      den[o + 2 * ds] = 4. * (red[2 * TP + tid] + red[6 * TP + tid]);
      den[o + 3 * ds] = 4. * (red[3 * TP + tid] + red[7 * TP + tid]);
    }
  }
}

// ---------------------------------------------------------------------------------------
// K_C: thread = point.  Functional, weight scaling (host driver :453-466), EXC / N_EL tile
// partials (:490-497, summed later in fixed order), and Z formation.  HBM-bound:
// reads k*8*nbe + writes 8*nbe bytes per point.
// ---------------------------------------------------------------------------------------
template <bool GGA>
__global__ void __launch_bounds__(TP) func_zmat_kernel(PlanView pv,
                                                        const DevTile* __restrict__ tiles,
                                                        double* __restrict__ ws,
                                                        const double* __restrict__ den, int ntiles,
                                                        FunctionalDesc func,
                                                        double* __restrict__ exc_part,
                                                        double* __restrict__ nel_part,
                                                        int part_off) {
  const DevTile tile = tiles[blockIdx.x];
  const DevTask task = pv.tasks[tile.task];
  const int i = threadIdx.x;
  const bool ok = i < tile.npts;
  const int nbe = task.nbe;
  const size_t ms = (size_t)nbe * TP;
  const size_t o = (size_t)blockIdx.x * TP + i;
  const size_t ds = (size_t)ntiles * TP;

  double a = 0., fx = 0., fy = 0., fz = 0., e_loc = 0., n_loc = 0.;
  if (ok) {
    const double w = pv.w[tile.pt_off + i];
    const double rho = den[o];
    double dx = 0., dy = 0., dz = 0., sigma = 0.;
    if (GGA) {
      dx = den[o + ds];
      dy = den[o + 2 * ds];
      dz = den[o + 3 * ds];
      sigma = dx * dx + dy * dy + dz * dz;
    }
    const XcOut xc = eval_functional(func, rho, sigma);
    const double eps = xc.eps * w;
    const double vrho = xc.vrho * w;
    a = 0.5 * vrho;
    if (GGA) {
      const double gf = 2. * (xc.vsigma * w);
      fx = gf * dx;
      fy = gf * dy;
      fz = gf * dz;
    }
    e_loc = eps * rho;
    n_loc = w * rho;
  }

  if (blockIdx.y == 0) {
    __shared__ double se[TP / 32], sn[TP / 32];
    double e = e_loc, n = n_loc;
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) {
      e += __shfl_xor_sync(0xffffffffu, e, d);
      n += __shfl_xor_sync(0xffffffffu, n, d);
    }
    if ((i & 31) == 0) {
      se[i >> 5] = e;
      sn[i >> 5] = n;
    }
    __syncthreads();
    if (i == 0) {
      exc_part[part_off + blockIdx.x] = (se[0] + se[1]) + (se[2] + se[3]);
      nel_part[part_off + blockIdx.x] = (sn[0] + sn[1]) + (sn[2] + sn[3]);
    }
  }

  const double* __restrict__ Bm = ws + tile.ws_off;
  double* __restrict__ Z = ws + tile.ws_off + (GGA ? 4 : 1) * ms;
  const int per = (nbe + gridDim.y - 1) / gridDim.y;
  const int m0 = blockIdx.y * per, m1 = min(nbe, m0 + per);
  for (int mu = m0; mu < m1; ++mu) {
    const size_t q = (size_t)mu * TP + i;
    double z = a * Bm[q];
    if (GGA) {
      z = fma(fx, Bm[q + ms], z);
      z = fma(fy, Bm[q + 2 * ms], z);
      z = fma(fz, Bm[q + 3 * ms], z);
    }
    Z[q] = z;
  }
}

// ---------------------------------------------------------------------------------------
// K_D: M = B^T Z per task (K = points), output block 128 x 64 per CTA accumulated over a run
// of tiles, then VXC_sub = M + M^T is scatter-added into the LOWER triangle of the full VXC
// with FP64 atomics (RED.ADD.F64).  Same warp layout as K_B.
// ---------------------------------------------------------------------------------------
constexpr int VB_M = 128, VB_N = 64, VB_K = 16, VB_STAGES = 3;
constexpr int V_LD = VB_K + 4;  // 20: (ld mod 16) == 4
constexpr int VB_SMEM = VB_STAGES * (VB_M + VB_N) * V_LD * 8;

__global__ void __launch_bounds__(256) vxc_kernel(PlanView pv, const DevTile* __restrict__ tiles,
                                                   const VxcItem* __restrict__ items,
                                                   const double* __restrict__ ws, int zmat,
                                                   double* __restrict__ VXC, int ldv) {
  extern __shared__ __align__(16) double smem[];
  double* As = smem;                              // [STAGES][VB_M][V_LD]
  double* Zs = smem + VB_STAGES * VB_M * V_LD;    // [STAGES][VB_N][V_LD]

  const VxcItem item = items[blockIdx.x];
  const DevTask task = pv.tasks[item.task];
  const int nbe = task.nbe;
  const int* __restrict__ ao = pv.task_ao + task.ao_off;
  const size_t ms = (size_t)nbe * TP;
  const int m0 = item.mblk * VB_M, n0 = item.nblk * VB_N;

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int g = lane >> 2, t = lane & 3;
  const int wm = warp & 3, wn = warp >> 2;
  const int mi_cnt = min(4, max(0, (nbe - m0 - wm * 32 + 7) / 8));
  const int ni_cnt = min(4, max(0, (nbe - n0 - wn * 32 + 7) / 8));

  // flattened (tile, k-chunk) sequence
  const int ntile = item.tile_end - item.tile_begin;
  int total = 0;
  for (int q = 0; q < ntile; ++q) total += (tiles[item.tile_begin + q].npts + VB_K - 1) / VB_K;

  // iterator state for the producer side
  int ld_tile = 0, ld_k = 0;
  auto load_stage = [&](int buf) {
    const DevTile tl = tiles[item.tile_begin + ld_tile];
    const double* __restrict__ Bm = ws + tl.ws_off;
    const double* __restrict__ Zm = Bm + (size_t)zmat * ms;
    const int k0 = ld_k * VB_K;
    double* as = As + buf * VB_M * V_LD;
    double* zs = Zs + buf * VB_N * V_LD;
#pragma unroll
    for (int r = 0; r < 4; ++r) {
      const int idx = tid + r * 256;
      const int row = idx >> 3, ch = idx & 7;
      const int mu = m0 + row;
      const bool p = mu < nbe;
      cp_async16(as + row * V_LD + ch * 2, Bm + (size_t)(p ? mu : 0) * TP + k0 + ch * 2, p);
    }
#pragma unroll
    for (int r = 0; r < 2; ++r) {
      const int idx = tid + r * 256;
      const int row = idx >> 3, ch = idx & 7;
      const int nu = n0 + row;
      const bool p = nu < nbe;
      cp_async16(zs + row * V_LD + ch * 2, Zm + (size_t)(p ? nu : 0) * TP + k0 + ch * 2, p);
    }
    if (++ld_k >= (tl.npts + VB_K - 1) / VB_K) {
      ld_k = 0;
      ++ld_tile;
    }
  };

  double acc[4][4][2];
#pragma unroll
  for (int mi = 0; mi < 4; ++mi)
#pragma unroll
    for (int ni = 0; ni < 4; ++ni) acc[mi][ni][0] = acc[mi][ni][1] = 0.;

  for (int s = 0; s < VB_STAGES - 1; ++s) {
    if (s < total) load_stage(s);
    cp_async_commit();
  }
  for (int it = 0; it < total; ++it) {
    cp_async_wait<VB_STAGES - 2>();
    __syncthreads();
    {
      const int nx = it + VB_STAGES - 1;
      if (nx < total) load_stage(nx % VB_STAGES);
      cp_async_commit();
    }
    if (mi_cnt > 0 && ni_cnt > 0) {
      const int buf = it % VB_STAGES;
      const double* as = As + buf * VB_M * V_LD + (wm * 32 + g) * V_LD + t;
      const double* zs = Zs + buf * VB_N * V_LD + (wn * 32 + g) * V_LD + t;
#pragma unroll
      for (int kk = 0; kk < 4; ++kk) {
        double a[4], b[4];
#pragma unroll
        for (int mi = 0; mi < 4; ++mi) a[mi] = as[mi * 8 * V_LD + kk * 4];
#pragma unroll
        for (int ni = 0; ni < 4; ++ni) b[ni] = zs[ni * 8 * V_LD + kk * 4];
#pragma unroll
        for (int mi = 0; mi < 4; ++mi)
#pragma unroll
          for (int ni = 0; ni < 4; ++ni)
            if (mi < mi_cnt && ni < ni_cnt) dmma(acc[mi][ni][0], acc[mi][ni][1], a[mi], b[ni]);
      }
    }
  }
  cp_async_wait<0>();

  // scatter: VXC_sub = M + M^T, only the lower triangle of the full matrix is accumulated
#pragma unroll
  for (int mi = 0; mi < 4; ++mi)
#pragma unroll
    for (int ni = 0; ni < 4; ++ni) {
      if (mi >= mi_cnt || ni >= ni_cnt) continue;
      const int mu = m0 + wm * 32 + mi * 8 + g;
      if (mu >= nbe) continue;
      const int gm = __ldg(ao + mu);
#pragma unroll
      for (int j = 0; j < 2; ++j) {
        const int nu = n0 + wn * 32 + ni * 8 + 2 * t + j;
        if (nu >= nbe) continue;
        const int gn = __ldg(ao + nu);
        double v = acc[mi][ni][j];
        if (mu == nu) v *= 2.;
        const int hi = max(gm, gn), lo = min(gm, gn);
        atomicAdd(VXC + (size_t)lo * ldv + hi, v);
      }
    }
}

// ---------------------------------------------------------------------------------------
// finalisation kernels
// ---------------------------------------------------------------------------------------
__device__ __forceinline__ void two_sum(double& s, double& c, double x) {
  // Neumaier compensated accumulation
  const double tsum = s + x;
  if (fabs(s) >= fabs(x)) c += (s - tsum) + x;
  else c += (x - tsum) + s;
  s = tsum;
}

__global__ void __launch_bounds__(256) reduce_partials_kernel(const double* __restrict__ e,
                                                               const double* __restrict__ n,
                                                               int cnt, double* __restrict__ out2) {
  __shared__ double sh[4][256];
  double es = 0, ec = 0, ns = 0, nc = 0;
  for (int i = threadIdx.x; i < cnt; i += 256) {
    two_sum(es, ec, e[i]);
    two_sum(ns, nc, n[i]);
  }
  sh[0][threadIdx.x] = es; sh[1][threadIdx.x] = ec;
  sh[2][threadIdx.x] = ns; sh[3][threadIdx.x] = nc;
  __syncthreads();
  if (threadIdx.x == 0) {
    double Es = 0, Ec = 0, Ns = 0, Nc = 0;
    for (int i = 0; i < 256; ++i) {
      two_sum(Es, Ec, sh[0][i]); Ec += sh[1][i];
      two_sum(Ns, Nc, sh[2][i]); Nc += sh[3][i];
    }
    out2[0] = Es + Ec;
    out2[1] = Ns + Nc;
  }
}

// upper <- lower (host driver :577-583; device K12 symmetrize_mat.cu)
__global__ void symmetrize_kernel(double* __restrict__ A, int n, int ld) {
  __shared__ double tile[32][33];
  const int bi = blockIdx.x, bj = blockIdx.y;
  if (bj > bi) return;  // only blocks on/below the diagonal are sources
  const int i = bi * 32 + threadIdx.x, j = bj * 32 + threadIdx.y;
  // A(i,j) col-major, i>=j in the lower triangle
  if (i < n && j < n) tile[threadIdx.y][threadIdx.x] = A[(size_t)j * ld + i];
  __syncthreads();
  // write A(j', i') = A(i', j') transposed block
  const int ti = bj * 32 + threadIdx.x;  // row index in the upper block
  const int tj = bi * 32 + threadIdx.y;  // col index in the upper block
  if (ti < n && tj < n && tj > ti) A[(size_t)tj * ld + ti] = tile[threadIdx.x][threadIdx.y];
}

}  // namespace

void launch_xmat_density(const PlanView& pv, const DevTile* tiles, int ntiles, double* ws,
                         const double* P, int ldp, double* den, bool gga, cudaStream_t s) {
  if (ntiles <= 0) return;
  static bool attr_set = false;
  if (!attr_set) {
    cudaFuncSetAttribute(xmat_density_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, XB_SMEM);
    cudaFuncSetAttribute(xmat_density_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, XB_SMEM);
    attr_set = true;
  }
  if (gga) xmat_density_kernel<true><<<ntiles, 256, XB_SMEM, s>>>(pv, tiles, ws, P, ldp, den, ntiles);
  else xmat_density_kernel<false><<<ntiles, 256, XB_SMEM, s>>>(pv, tiles, ws, P, ldp, den, ntiles);
}

void launch_func_zmat(const PlanView& pv, const DevTile* tiles, int ntiles, double* ws,
                      const double* den, FunctionalDesc func, double* exc_part, double* nel_part,
                      int part_off, cudaStream_t s) {
  if (ntiles <= 0) return;
  // split the mu loop so that small batches still fill the machine
  int ysplit = 1;
  while (ntiles * ysplit < 148 * 8 && ysplit < 8) ysplit *= 2;
  dim3 grid(ntiles, ysplit);
  if (func.is_gga)
    func_zmat_kernel<true><<<grid, TP, 0, s>>>(pv, tiles, ws, den, ntiles, func, exc_part, nel_part, part_off);
  else
    func_zmat_kernel<false><<<grid, TP, 0, s>>>(pv, tiles, ws, den, ntiles, func, exc_part, nel_part, part_off);
}

void launch_vxc(const PlanView& pv, const DevTile* tiles, const VxcItem* items, int nitems,
                const double* ws, bool gga, double* VXC, int ldv, cudaStream_t s) {
  if (nitems <= 0) return;
  static bool attr_set = false;
  if (!attr_set) {
    cudaFuncSetAttribute(vxc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, VB_SMEM);
    attr_set = true;
  }
  vxc_kernel<<<nitems, 256, VB_SMEM, s>>>(pv, tiles, items, ws, gga ? 4 : 1, VXC, ldv);
}

void launch_reduce_partials(const double* exc_part, const double* nel_part, int n, double* out2,
                            cudaStream_t s) {
  reduce_partials_kernel<<<1, 256, 0, s>>>(exc_part, nel_part, n, out2);
}

void launch_symmetrize(double* VXC, int nbf, int ldv, cudaStream_t s) {
  const int nb = (nbf + 31) / 32;
  symmetrize_kernel<<<dim3(nb, nb), dim3(32, 32), 0, s>>>(VXC, nbf, ldv);
}

}  // namespace gxb
