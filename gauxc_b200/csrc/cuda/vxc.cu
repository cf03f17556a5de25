// K_D: VXC_sub = B^T Z + Z^T B on the FP64 DMMA pipe with Z formed on the fly and a scatter-add into the
// full matrix, plus the finalisation kernels (partials reduction, symmetrise).
//
// Host semantics matched: eval_zmat_{lda,gga}_vxc_rks + inc_vxc (reference_local_host_work_driver.cxx:
// 586-604, 678-713, 1678-1692: Z = 1/2 vrho B + 2 vgamma grad rho . grad B, dsyr2k lower +
// inc_by_submat_atomic, host/util.hpp:130-168) and the symmetrise loop of the host driver
// (reference_replicated_xc_host_integrator_exc_vxc.hpp:577-583).  Replaces the reference device path's
// zmat kernel + per-task cuBLAS dsyr2k + sym_task_inc_potential (zmat_vxc.cu, scheme1_base.cxx:1711-1766,
// cuda_inc_potential.cu) by one grouped launch per batch.
//
// An item = one 128 x 64 output block (mblk, nblk) of M = B^T Z of one task, accumulated over a run of
// that task's tiles (K = points, 16 per pipeline stage).  Z is NEVER stored in HBM: the fused kernel leaves
// four factors per point (a, fx, fy, fz) in the tile's factor rows and this kernel forms
// Z = a B + fx dBx + fy dBy + fz dBz per stage (LDA: Z = a B).  One persistent CTA per SM, four roles:
//   producer (1 thread)  item queue (atomic counter, task order -> the VXC region being scattered into stays
//                        in L2) + one TMA box (128 basis rows x 16 points of B) per stage, 5-stage mbarrier ring
//   combiners (8 warps)  the 64 x 16 Z block of the stage: 128-bit streaming loads of B / dB straight from the
//                        swizzled workspace (thread = two rows x two points, 8 loads in flight), four FMAs,
//                        one 128-bit store into the stage -- in the layout a TMA box would have produced
//   MMA (8 warps)        4 x 2 warps, warp tile 32 x 32, two MMA warps per SM sub-partition (enough to
//                        saturate the DMMA pipe); a finished block is dumped to shared memory
//   scatter (4 warps)    M + M^T folded into the LOWER triangle of VXC with FP64 reductions (RED.ADD.F64,
//                        lane = row: coalesced) WHILE the MMA warps run the next item -- the scatter of small
//                        tasks (nbe^2 REDs per ~100 points) used to serialise with the K loop
// For LDA Z = diag(a) B makes M symmetric: only blocks that touch the lower triangle are scheduled (`sym`).
#include "kernels.cuh"
#include "ptx.cuh"

namespace gxb {

namespace {

constexpr int VK = 16;  // points (K) per stage
constexpr int VSTAGES = 5;
constexpr int VQ = 4;   // item-queue ring slots
constexpr int V_MMA_WARPS = 8, V_COMB_WARPS = 8, V_SCAT_WARPS = 4;
constexpr int V_MMA_THREADS = V_MMA_WARPS * 32, V_COMB_THREADS = V_COMB_WARPS * 32;
// warps 0-7 MMA, 8-15 combiners, 16-19 scatter, 20 producer, 21-23 idle (complete the producer's
// warpgroup for setmaxnreg)
constexpr int V_THREADS = (V_MMA_WARPS + V_COMB_WARPS + V_SCAT_WARPS + 4) * 32;
// launch allocation 24 warps x 80; after re-partitioning 8 x 112 + 8 x 96 + 4 x 40 + 4 x 24
constexpr int V_MMA_REGS = 112, V_COMB_REGS = 96, V_SCAT_REGS = 40, V_PROD_REGS = 24;
static_assert(8 * V_MMA_REGS + 8 * V_COMB_REGS + 4 * V_SCAT_REGS + 4 * V_PROD_REGS <= 24 * 80, "register pool");
constexpr int OUT_LD = VXC_BLK + 2;  // conflict-free accumulator dump (C fragment: rows g, columns 2t, 2t+1)

struct VxcSlot {
  int nbe, ao_off, m0, n0, row0, ntiles, nks_last, diag;  // ntiles < 0: queue drained
};

struct VxcSmem {
  double A[VSTAGES][VXC_BLK][VK];
  double Z[VSTAGES][VXC_BLN][VK];
  double out[VXC_BLN][OUT_LD];  // out[nu][mu]
  double fac[2][4][TP];         // factor rows of the tile the combiners work on (double-buffered)
  uint64_t full[VSTAGES], empty[VSTAGES];
  uint64_t qfull[VQ], qempty[VQ];
  uint64_t ofull, oempty;
  VxcSlot q[VQ];
};
constexpr size_t VXC_SMEM_BYTES = sizeof(VxcSmem) + 1024;
static_assert(VXC_SMEM_BYTES <= 232448, "vxc kernel shared memory");

// One K step (16 points) of a warp's 32 x 32 tile restricted to MI x NI 8 x 8 blocks: straight-line,
// UNPREDICATED DMMAs (a predicated mma.sync costs a WARPSYNC each); ragged blocks dispatch once per
// K step to the variant with the counts rounded up to even -- rows beyond nbe are zero pad rows or
// rows of the neighbouring matrix (finite), and their accumulators are never scattered.
template <int MI, int NI>
__device__ __forceinline__ void vxc_step(double (&acc)[4][4][2], const double* __restrict__ as,
                                         const double* __restrict__ zs, int t, int sw) {
#pragma unroll
  for (int kk = 0; kk < 4; ++kk) {
    const int col = (kk * 4 + t) ^ sw;
    double a[MI], b[NI];
#pragma unroll
    for (int mi = 0; mi < MI; ++mi) a[mi] = as[mi * 8 * VK + col];
#pragma unroll
    for (int ni = 0; ni < NI; ++ni) b[ni] = zs[ni * 8 * VK + col];
#pragma unroll
    for (int mi = 0; mi < MI; ++mi)
#pragma unroll
      for (int ni = 0; ni < NI; ++ni) dmma(acc[mi][ni][0], acc[mi][ni][1], a[mi], b[ni]);
  }
}

// GGA: Z = a B + fx dBx + fy dBy + fz dBz (matrices 0..3 of a tile); LDA: Z = a B (matrix 0)
template <bool GGA>
__global__ void __launch_bounds__(V_THREADS, 1)
vxc_kernel(const __grid_constant__ CUtensorMap tmapV, PlanView pv, const VxcItem* __restrict__ items,
           int nitems, int* __restrict__ counter, const double* __restrict__ ws, int fac_row, int sym,
           double* __restrict__ VXC, int ldv) {
  // no pointer arithmetic on the base: the compiler must see shared-space accesses (LDS/STS, not
  // generic LD/ST) in the fragment loads
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  VxcSmem& S = *reinterpret_cast<VxcSmem*>(smem_raw);
  if (threadIdx.x == 0 && (smem_u32(smem_raw) & 127u)) __trap();  // TMA destinations need 128 B

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  constexpr int NMAT = GGA ? 4 : 1;

  if (tid == 0) {
    for (int s = 0; s < VSTAGES; ++s) {
      mbar_init(&S.full[s], 1 + V_COMB_WARPS);  // producer (TMA bytes of B^T) + one elected arrive per combiner warp
      mbar_init(&S.empty[s], V_MMA_WARPS);      // one elected arrive per MMA warp
    }
    for (int i = 0; i < VQ; ++i) {
      mbar_init(&S.qfull[i], 1);
      mbar_init(&S.qempty[i], V_MMA_WARPS + V_COMB_WARPS + V_SCAT_WARPS);
    }
    mbar_init(&S.ofull, V_MMA_WARPS);
    mbar_init(&S.oempty, V_SCAT_WARPS);
    mbar_fence_init();
  }
  __syncthreads();

  // consumer side of the item queue: iteration `it` reads slot it % VQ
  auto next_item = [&](int it) {
    const int slot = it & (VQ - 1);
    mbar_wait(&S.qfull[slot], (it / VQ) & 1);
    const VxcSlot sl = S.q[slot];
    __syncwarp();
    if (lane == 0) mbar_arrive(&S.qempty[slot]);
    return sl;
  };

  if (warp >= V_MMA_WARPS + V_COMB_WARPS + V_SCAT_WARPS) {
    // ---------------------------------------------------------------- producer (one thread)
    reg_dec<V_PROD_REGS>();
    if (warp != V_MMA_WARPS + V_COMB_WARPS + V_SCAT_WARPS || lane != 0) return;
    tma_prefetch_desc(&tmapV);
    int s = 0;
    uint32_t ph = 0;
    // the item record is self-contained (no dependent task / tile loads) and fetched ONE ITEM AHEAD,
    // so the queue pop + metadata latency hides behind the TMA issue loop of the current item
    auto fetch = [&](int idx) {
      VxcItem r;
      if (idx < nitems) {
        const int4* src = reinterpret_cast<const int4*>(items + idx);
        const int4 a = __ldg(src), b = __ldg(src + 1);
        r.nbe = a.x; r.ao_off = a.y; r.mblk = a.z; r.nblk = a.w;
        r.row0 = b.x; r.ntiles = b.y; r.nks_last = b.z; r.pad = b.w;
      } else {
        r.nbe = r.ao_off = r.mblk = r.nblk = r.row0 = r.ntiles = r.nks_last = r.pad = 0;
      }
      return r;
    };
    // queue pops run two items ahead and record loads one item ahead, so neither latency is on the
    // critical path of an item -- except near the end of the queue (fewer than `margin` items left),
    // where holding popped items would unbalance the tail: there the next item is popped only after
    // the current one has been issued
    const int margin = 4 * (int)gridDim.x;
    auto pop = [&]() { return atomicAdd(counter, 1); };
    int idx = pop();
    int idx1 = (idx + margin < nitems) ? pop() : -1;
    VxcItem item = fetch(idx);
    for (int it = 0;; ++it) {
      const int slot = it & (VQ - 1);
      mbar_wait(&S.qempty[slot], ((it / VQ) & 1) ^ 1);
      if (idx >= nitems) {
        S.q[slot].ntiles = -1;
        mbar_arrive(&S.qfull[slot]);
        break;
      }
      VxcItem nxt = item;
      int idx2 = -1;
      if (idx1 >= 0) {
        nxt = fetch(idx1);
        if (idx1 + margin < nitems) idx2 = pop();
      }
      const int m0 = item.mblk * VXC_BLK, n0 = item.nblk * VXC_BLN;
      VxcSlot sl;
      sl.nbe = item.nbe; sl.ao_off = item.ao_off; sl.m0 = m0; sl.n0 = n0;
      sl.row0 = item.row0; sl.ntiles = item.ntiles; sl.nks_last = item.nks_last;
      sl.diag = (sym && n0 + VXC_BLN - 1 > m0) ? 1 : 0;  // the block reaches above the diagonal
      S.q[slot] = sl;
      mbar_arrive(&S.qfull[slot]);
      const int stride = tile_rows(NMAT, item.nbe);
      for (int q = 0; q < item.ntiles; ++q) {
        const int rowB = item.row0 + q * stride;
        const int nks = (q + 1 < item.ntiles) ? TP / VK : item.nks_last;
        for (int ks = 0; ks < nks; ++ks) {
          mbar_wait(&S.empty[s], ph ^ 1);
          mbar_arrive_expect_tx(&S.full[s], VXC_BLK * VK * sizeof(double));
          tma_load_2d(&S.A[s][0][0], &tmapV, &S.full[s], ks * VK, rowB + m0);
          if (++s == VSTAGES) { s = 0; ph ^= 1; }
        }
      }
      if (idx1 < 0) {  // tail of the queue: lazy pop
        idx1 = pop();
        nxt = fetch(idx1);
      }
      item = nxt;
      idx = idx1;
      idx1 = idx2;
    }
    return;
  }

  if (warp >= V_MMA_WARPS + V_COMB_WARPS) {
    // ---------------------------------------------------------------- scatter warps
    reg_dec<V_SCAT_REGS>();
    const int sw = warp - (V_MMA_WARPS + V_COMB_WARPS);
    uint32_t oph = 0;
    for (int it = 0;; ++it) {
      const VxcSlot sl = next_item(it);
      if (sl.ntiles < 0) break;
      const int nbe = sl.nbe, m0 = sl.m0, n0 = sl.n0;
      const int* __restrict__ ao = pv.task_ao + sl.ao_off;
      // global AO index of the lane's row in each of the four 32-row groups of the block
      int gm[4];
#pragma unroll
      for (int rb = 0; rb < 4; ++rb) {
        const int mu = m0 + rb * 32 + lane;
        gm[rb] = mu < nbe ? __ldg(ao + mu) : -1;
      }
      const int ncols = min(VXC_BLN, nbe - n0);
      mbar_wait(&S.ofull, oph);
      // VXC_sub = M + M^T, only the lower triangle of the full matrix is accumulated
      for (int c = sw; c < ncols; c += V_SCAT_WARPS) {
        const int nu = n0 + c;
        const int gn = __ldg(ao + nu);
#pragma unroll
        for (int rb = 0; rb < 4; ++rb) {
          const int mu = m0 + rb * 32 + lane;
          if (gm[rb] < 0) continue;  // mu >= nbe
          double v = S.out[c][rb * 32 + lane];
          if (sym) {
            if (nu > mu) continue;  // M symmetric: the mirror entry carries it
            v *= 2.;
          } else if (mu == nu) {
            v *= 2.;
          }
          const int hi = max(gm[rb], gn), lo = min(gm[rb], gn);
          atomicAdd(VXC + (size_t)lo * ldv + hi, v);
        }
      }
      __syncwarp();
      if (lane == 0) mbar_arrive(&S.oempty);
      oph ^= 1;
    }
    return;
  }

  if (warp >= V_MMA_WARPS) {
    // ---------------------------------------------------------------- combiner warps
    // The stream of (tile, 16-point stage) steps of an item is software-pipelined: the 8 (GGA) 128-bit loads
    // of step j + 1 are in flight while step j is combined and stored, so 64 KB per SM are on their way at any
    // time -- the L2 / HBM latency (~1 us loaded) times the 31 GB/s per SM this stream must sustain at the
    // DMMA peak.  The factor rows of a tile (4 x 1 KB) are staged in shared memory once per tile.
    reg_inc<V_COMB_REGS>();  // 96 > the 80 of the launch allocation
    const int c = tid - V_MMA_THREADS;
    const int pq = c & 7;    // physical 16-byte pair inside the 128-byte line of a row
    const int rr = c >> 3;   // rows rr and rr + 32 of the 64-row block (same row & 3: same swizzle)
    // logical points of the physical pair (swizzle: column = point ^ ((row & 3) << 2))
    const int i0 = (2 * pq) ^ ((rr & 3) << 2);
    int s = 0;
    uint32_t ph = 0;
    int fbuf = 0;
    for (int it = 0;; ++it) {
      const VxcSlot sl = next_item(it);
      if (sl.ntiles < 0) break;
      const int nbp = pad16(sl.nbe);
      const size_t stride = (size_t)tile_rows(NMAT, sl.nbe) * TP;
      const size_t ms = (size_t)nbp * TP;
      const bool v0 = sl.n0 + rr < nbp, v1 = sl.n0 + rr + 32 < nbp;  // rows beyond the pad rows: zeros
      const int nsteps = (sl.ntiles - 1) * (TP / VK) + sl.nks_last;
      const double* __restrict__ src0 = ws + (size_t)(sl.row0 + sl.n0 + rr) * TP + 2 * pq;
      const double* __restrict__ fac0 = ws + (size_t)(sl.row0 + NMAT * nbp + fac_row) * TP;
      auto load = [&](double (&b)[2][NMAT][2], int j) {
        const double* src = src0 + (size_t)(j >> 3) * stride + (j & 7) * VK;
#pragma unroll
        for (int m = 0; m < NMAT; ++m) {
          if (v0) ldg128_stream(b[0][m], src + m * ms);
          else b[0][m][0] = b[0][m][1] = 0.;
          if (v1) ldg128_stream(b[1][m], src + m * ms + (size_t)32 * TP);
          else b[1][m][0] = b[1][m][1] = 0.;
        }
      };
      // factor rows of tile q -> S.fac[fbuf]: 4 KB (GGA) = one 128-bit load + store per combiner thread
      auto stage_factors = [&](int q) {
        if (c < NMAT * (TP / 2)) {
          double f[2];
          ldg128_stream(f, fac0 + (size_t)q * stride + (size_t)(c >> 6) * TP + 2 * (c & 63));
          *reinterpret_cast<double2*>(&S.fac[fbuf][c >> 6][2 * (c & 63)]) = make_double2(f[0], f[1]);
        }
        named_bar_sync(1, V_COMB_THREADS);
      };
      double bn[2][NMAT][2];
      load(bn, 0);
      for (int j = 0; j < nsteps; ++j) {
        if ((j & 7) == 0) {
          fbuf ^= 1;  // the previous tile's rows may still be read by slower warps of this stage
          stage_factors(j >> 3);
        }
        double bc[2][NMAT][2];
#pragma unroll
        for (int u = 0; u < 2; ++u)
#pragma unroll
          for (int m = 0; m < NMAT; ++m) { bc[u][m][0] = bn[u][m][0]; bc[u][m][1] = bn[u][m][1]; }
        if (j + 1 < nsteps) load(bn, j + 1);
        double z[2][2];
        const int col = (j & 7) * VK + i0;
#pragma unroll
        for (int u = 0; u < 2; ++u) z[u][0] = z[u][1] = 0.;
#pragma unroll
        for (int m = 0; m < NMAT; ++m) {
          const double2 f = *reinterpret_cast<const double2*>(&S.fac[fbuf][m][col]);
#pragma unroll
          for (int u = 0; u < 2; ++u) {
            z[u][0] = (m == 0) ? f.x * bc[u][0][0] : fma(f.x, bc[u][m][0], z[u][0]);
            z[u][1] = (m == 0) ? f.y * bc[u][0][1] : fma(f.y, bc[u][m][1], z[u][1]);
          }
        }
        mbar_wait(&S.empty[s], ph ^ 1);
        *reinterpret_cast<double2*>(&S.Z[s][rr][2 * pq]) = make_double2(z[0][0], z[0][1]);
        *reinterpret_cast<double2*>(&S.Z[s][rr + 32][2 * pq]) = make_double2(z[1][0], z[1][1]);
        __syncwarp();
        if (lane == 0) mbar_arrive(&S.full[s]);
        if (++s == VSTAGES) { s = 0; ph ^= 1; }
      }
    }
    return;
  }

  // ------------------------------------------------------------------ MMA warps
  reg_inc<V_MMA_REGS>();
  const int g = lane >> 2, t = lane & 3;
  // 4 x 2 warps; sub-partition = warp & 3 = (wm + wn) & 3: idle row blocks of ragged output blocks are
  // spread over the DMMA pipes
  const int wn = warp >> 2;
  const int wm = ((warp & 3) - wn) & 3;
  const int sw = (g & 3) << 2;
  int s = 0;
  uint32_t ph = 0, oph = 0;

  for (int it = 0;; ++it) {
    const VxcSlot sl = next_item(it);
    if (sl.ntiles < 0) break;
    const int nbe = sl.nbe, m0 = sl.m0, n0 = sl.n0;
    const int nks = (sl.ntiles - 1) * (TP / VK) + sl.nks_last;
    const int mi_cnt = min(4, max(0, (nbe - m0 - wm * 32 + 7) / 8));
    const int ni_cnt = min(4, max(0, (nbe - n0 - wn * 32 + 7) / 8));
    // symmetric M: a warp tile entirely above the diagonal is carried by its mirror
    const bool above = sl.diag && (n0 + wn * 32 > m0 + wm * 32 + 31);
    const bool active = mi_cnt > 0 && ni_cnt > 0 && !above;
    const int var = (mi_cnt > 2 ? 2 : 0) | (ni_cnt > 2 ? 1 : 0);

    double acc[4][4][2];
#pragma unroll
    for (int mi = 0; mi < 4; ++mi)
#pragma unroll
      for (int ni = 0; ni < 4; ++ni) acc[mi][ni][0] = acc[mi][ni][1] = 0.;

    for (int ks = 0; ks < nks; ++ks) {
      mbar_wait(&S.full[s], ph);
      if (active) {
        const double* as = &S.A[s][wm * 32 + g][0];
        const double* zs = &S.Z[s][wn * 32 + g][0];
        switch (var) {
          case 3: vxc_step<4, 4>(acc, as, zs, t, sw); break;
          case 2: vxc_step<4, 2>(acc, as, zs, t, sw); break;
          case 1: vxc_step<2, 4>(acc, as, zs, t, sw); break;
          default: vxc_step<2, 2>(acc, as, zs, t, sw); break;
        }
      }
      // one arrive per warp: per-thread arrives on one mbarrier would serialise in the shared-memory
      // atomic unit every K step
      __syncwarp();
      if (lane == 0) mbar_arrive(&S.empty[s]);
      if (++s == VSTAGES) { s = 0; ph ^= 1; }
    }
    // hand the block to the scatter warps (they finished the previous one long ago unless the kernel is
    // RED-bound, in which case this wait is the bound)
    mbar_wait(&S.oempty, oph ^ 1);
    if (active) {
#pragma unroll
      for (int mi = 0; mi < 4; ++mi)
#pragma unroll
        for (int ni = 0; ni < 4; ++ni)
#pragma unroll
          for (int j = 0; j < 2; ++j)
            S.out[wn * 32 + ni * 8 + 2 * t + j][wm * 32 + mi * 8 + g] = acc[mi][ni][j];
    }
    __syncwarp();
    if (lane == 0) mbar_arrive(&S.ofull);
    oph ^= 1;
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
  const int ti = bj * 32 + threadIdx.x;  // row index in the upper block
  const int tj = bi * 32 + threadIdx.y;  // col index in the upper block
  if (ti < n && tj < n && tj > ti) A[(size_t)tj * ld + ti] = tile[threadIdx.x][threadIdx.y];
}

// lower triangle (column-major, column j holds rows j..n-1) <-> packed vector, for the single allreduce of
// [tril(VXC) | EXC | N_EL]; the unpack also mirrors (upper <- lower)
__device__ __forceinline__ size_t tril_off(int j, int n) { return (size_t)j * n - (size_t)j * (j - 1) / 2; }
__global__ void pack_tril_kernel(const double* __restrict__ A, int n, int ld, double* __restrict__ out) {
  const int j = blockIdx.y;
  const int i = j + blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) out[tril_off(j, n) + (i - j)] = A[(size_t)j * ld + i];
}
__global__ void unpack_tril_sym_kernel(const double* __restrict__ in, double* __restrict__ A, int n, int ld) {
  __shared__ double tile[32][33];
  const int bi = blockIdx.x, bj = blockIdx.y;
  if (bj > bi) return;  // only blocks on/below the diagonal are sources
  const int i = bi * 32 + threadIdx.x, j = bj * 32 + threadIdx.y;
  double v = 0.;
  if (i < n && j < n && i >= j) {
    v = in[tril_off(j, n) + (i - j)];
    A[(size_t)j * ld + i] = v;
  }
  tile[threadIdx.y][threadIdx.x] = v;
  __syncthreads();
  const int ti = bj * 32 + threadIdx.x;  // row index in the upper block
  const int tj = bi * 32 + threadIdx.y;  // col index in the upper block
  if (ti < n && tj < n && tj > ti) A[(size_t)tj * ld + ti] = tile[threadIdx.x][threadIdx.y];
}

// LDA density operand: P' = lower triangle of (P + P^T)/2 with the diagonal halved, so that
// rho = B.(P B) = 2 B.(P' B) is summed over k <= n only (fused.cu, LDA K loop).  Exact for a
// symmetric P ((a + a)/2 == a); for a non-symmetric P it is the same quadratic form as the host's.
__global__ void sym_half_kernel(const double* __restrict__ P, int ldp, double* __restrict__ out, int n) {
  __shared__ double tile[32][33];
  const int bi = blockIdx.x, bj = blockIdx.y;
  if (bj > bi) return;
  {
    const int r = bj * 32 + threadIdx.x, c = bi * 32 + threadIdx.y;  // mirror block, element (r, c)
    tile[threadIdx.y][threadIdx.x] = (r < n && c < n) ? P[(size_t)c * ldp + r] : 0.;
  }
  __syncthreads();
  const int i = bi * 32 + threadIdx.x, j = bj * 32 + threadIdx.y;
  if (i < n && j < n && i >= j) {
    const double a = P[(size_t)j * ldp + i], b = tile[threadIdx.x][threadIdx.y];  // P(i,j), P(j,i)
    out[(size_t)j * n + i] = (i == j) ? 0.5 * a : 0.5 * (a + b);
  }
}

}  // namespace

void launch_sym_half(const double* P, int ldp, double* out, int nbf, cudaStream_t s) {
  const int nb = (nbf + 31) / 32;
  if (nb > 0) sym_half_kernel<<<dim3(nb, nb), dim3(32, 32), 0, s>>>(P, ldp, out, nbf);
}

cudaError_t launch_vxc(const CUtensorMap& tmapV, const PlanView& pv, const VxcItem* items, int nitems,
                       int* counter, int ncta, const double* ws, bool gga, int fac_row, bool sym, double* VXC,
                       int ldv, cudaStream_t s) {
  if (nitems <= 0 || ncta <= 0) return cudaSuccess;
  // the attribute is per device and cheap to set: no process-wide "done" flag
  cudaError_t e = gga ? cudaFuncSetAttribute(vxc_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                             (int)VXC_SMEM_BYTES)
                      : cudaFuncSetAttribute(vxc_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                             (int)VXC_SMEM_BYTES);
  if (e != cudaSuccess) return e;
  ncta = ncta < nitems ? ncta : nitems;
  if (gga)
    vxc_kernel<true><<<ncta, V_THREADS, VXC_SMEM_BYTES, s>>>(tmapV, pv, items, nitems, counter, ws, fac_row,
                                                             sym ? 1 : 0, VXC, ldv);
  else
    vxc_kernel<false><<<ncta, V_THREADS, VXC_SMEM_BYTES, s>>>(tmapV, pv, items, nitems, counter, ws, fac_row,
                                                              sym ? 1 : 0, VXC, ldv);
  return cudaGetLastError();
}

void launch_pack_tril(const double* A, int n, int ld, double* out, cudaStream_t s) {
  if (n > 0) pack_tril_kernel<<<dim3((n + 255) / 256, n), 256, 0, s>>>(A, n, ld, out);
}
void launch_unpack_tril_sym(const double* in, double* A, int n, int ld, cudaStream_t s) {
  const int nb = (n + 31) / 32;
  if (nb > 0) unpack_tril_sym_kernel<<<dim3(nb, nb), dim3(32, 32), 0, s>>>(in, A, n, ld);
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
