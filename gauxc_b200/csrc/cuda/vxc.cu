// K_D: VXC_sub = B^T Z + Z^T B on the FP64 DMMA pipe with a scatter-add into the full matrix,
// plus the finalisation kernels (partials reduction, pack / unpack, symmetrise).
//
// Host semantics matched: inc_vxc (reference_local_host_work_driver.cxx:1678-1692: dsyr2k lower +
// inc_by_submat_atomic, host/util.hpp:130-168) and the symmetrise loop of the host driver
// (reference_replicated_xc_host_integrator_exc_vxc.hpp:577-583).  Replaces the reference device
// path's per-task cuBLAS dsyr2k + sym_task_inc_potential (scheme1_base.cxx:1711-1766,
// cuda_inc_potential.cu) by one grouped launch per batch.
//
// An item = one 128 x 64 output block (mblk, nblk) of M = B^T Z of one task, accumulated over a
// run of that task's tiles (K = points).  Both operands arrive as TMA boxes (128 resp. 64 basis rows x
// 16 points) straight from the swizzled workspace (dense + conflict-free, device_plan.hpp) through
// a 4-stage mbarrier ring fed by one producer thread; 8 MMA warps (4 x 2, warp tile 32 x 32).
// The kernel is PERSISTENT and runs TWO CTAs PER SM (384 threads, 97 KB of shared memory each): every CTA pulls items from a device-side queue (atomic counter, task order -> the
// VXC region being scattered into stays in L2), its producer runs ahead into the next item's loads, and
// while the MMA warps of one CTA scatter a finished block with FP64 reductions (RED.ADD.F64: nbe^2 of
// them per task and tile run, as expensive as the K loop itself for the ~100-point tasks that dominate
// large molecules) the other CTA's K loop keeps the DMMA pipe busy -- 16 MMA warps per SM in all, the
// occupancy at which the pipe saturates.  M + M^T is folded into the LOWER triangle of VXC.  For LDA
// Z = diag(1/2 w vrho) B makes M symmetric: only blocks that reach the lower triangle are scheduled
// (`sym`).
//
// Measured alternative (profiles/r02b_*): forming Z = a B + f . dB on the fly from per-point factors
// (dedicated combiner warps streaming B / dB, scatter delegated to further warps, no Z in HBM) removed
// the Z pass from the fused kernel (32 -> 15 GB per launch) but tripled the bytes per flop of this one
// (28 GB per launch, 47 % L2 hits) and starved its DMMA pipe: taxol VXC 100 -> 133 ms, ubiquitin
// 286 -> 436 ms.  The step is bound by feeding the DMMA pipe, not by HBM; Z stays materialised.
#include "kernels.cuh"
#include "ptx.cuh"
#include <cstdio>
#include <cstdlib>

namespace gxb {

namespace {

constexpr int VK = 16;  // points (K) per stage
constexpr int VSTAGES = 4;
constexpr int VQ = 4;   // item-queue ring slots
#ifndef GXB_VPF
#define GXB_VPF 0
#endif
constexpr int VPF = GXB_VPF;  // stages prefetched into L2 ahead of the ring (0: none)
constexpr int V_MMA_WARPS = 8;
constexpr int V_MMA_THREADS = V_MMA_WARPS * 32;
// warps 0-7 MMA, warp 8 producer (one thread), 9-11 idle (complete the producer's warpgroup: registers are
// allocated in units of four warps, and setmaxnreg is a warpgroup instruction).  Launch allocation 12 warps
// x 80; after re-partitioning 8 x 104 + 4 x 32 -- twice per SM.
constexpr int V_THREADS = V_MMA_THREADS + 128;
constexpr int V_MMA_REGS = 104, V_PROD_REGS = 32;
static_assert(8 * V_MMA_REGS + 4 * V_PROD_REGS <= 12 * 80, "register pool of one CTA");

struct VxcSlot {
  int nbe, ao_off, m0, n0, nks, diag, pad0, pad1;  // nks < 0: queue drained
};

struct VxcSmem {
  double A[VSTAGES][VXC_BLK][VK];
  double Z[VSTAGES][VXC_BLN][VK];
  uint64_t full[VSTAGES], empty[VSTAGES];
  uint64_t qfull[VQ], qempty[VQ];
  VxcSlot q[VQ];
};
constexpr size_t VXC_SMEM_BYTES = sizeof(VxcSmem) + 1024;
static_assert(2 * (VXC_SMEM_BYTES + 1024) <= 233472, "two VXC CTAs per SM");

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

__global__ void __launch_bounds__(V_THREADS, 2)
vxc_kernel(const __grid_constant__ CUtensorMap tmapA, const __grid_constant__ CUtensorMap tmapZ, PlanView pv,
           const VxcItem* __restrict__ items, int nitems, int* __restrict__ counter, int zmat, int nmat,
           int sym, double* __restrict__ VXC, int ldv) {
  // no pointer arithmetic on the base: the compiler must see shared-space accesses (LDS/STS, not
  // generic LD/ST) in the fragment loads
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  VxcSmem& S = *reinterpret_cast<VxcSmem*>(smem_raw);
  if (threadIdx.x == 0 && (smem_u32(smem_raw) & 127u)) __trap();  // TMA destinations need 128 B

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;

  if (tid == 0) {
    for (int s = 0; s < VSTAGES; ++s) {
      mbar_init(&S.full[s], 1);
      mbar_init(&S.empty[s], V_MMA_WARPS);  // one elected arrive per MMA warp
    }
    for (int i = 0; i < VQ; ++i) {
      mbar_init(&S.qfull[i], 1);
      mbar_init(&S.qempty[i], V_MMA_WARPS);
    }
    mbar_fence_init();
  }
  __syncthreads();

  if (warp >= V_MMA_WARPS) {
    // ---------------------------------------------------------------- producer (one thread)
    reg_dec<V_PROD_REGS>();
    if (warp != V_MMA_WARPS || lane != 0) return;
    tma_prefetch_desc(&tmapA);
    tma_prefetch_desc(&tmapZ);
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
        S.q[slot].nks = -1;
        mbar_arrive(&S.qfull[slot]);
        break;
      }
      VxcItem nxt = item;
      int idx2 = -1;
      if (idx1 >= 0) {
        nxt = fetch(idx1);
        if (idx1 + margin < nitems) idx2 = pop();
      }
      const int nbp = pad16(item.nbe);
      const int m0 = item.mblk * VXC_BLK, n0 = item.nblk * VXC_BLN;
      VxcSlot sl;
      sl.nbe = item.nbe; sl.ao_off = item.ao_off; sl.m0 = m0; sl.n0 = n0;
      sl.nks = (item.ntiles - 1) * (TP / VK) + item.nks_last;
      sl.diag = (sym && n0 + VXC_BLN - 1 > m0) ? 1 : 0;  // the block reaches above the diagonal
      sl.pad0 = sl.pad1 = 0;
      S.q[slot] = sl;
      mbar_arrive(&S.qfull[slot]);
      const int stride = nmat * nbp;  // workspace rows of one tile
      // The boxes of stage j + VPF are prefetched into L2 (TMA prefetch: no shared memory, no completion)
      // when stage j is issued: the ring holds only 4 stages per CTA, HBM latency under load is several of
      // them, and the ring cannot grow (two CTAs share the SM's shared memory).
      const int nst = (item.ntiles - 1) * (TP / VK) + item.nks_last;
      auto prefetch = [&](int j) {
        const int rowB = item.row0 + (j >> 3) * stride, ks = j & 7;
        tma_prefetch_2d(&tmapA, ks * VK, rowB + m0);
        tma_prefetch_2d(&tmapZ, ks * VK, rowB + zmat * nbp + n0);
      };
      for (int j = 0; j < VPF && j < nst; ++j) prefetch(j);
      for (int j = 0; j < nst; ++j) {
        const int rowB = item.row0 + (j >> 3) * stride, ks = j & 7;
        if (j + VPF < nst) prefetch(j + VPF);
        mbar_wait(&S.empty[s], ph ^ 1);
        mbar_arrive_expect_tx(&S.full[s], (VXC_BLK + VXC_BLN) * VK * sizeof(double));
        tma_load_2d(&S.A[s][0][0], &tmapA, &S.full[s], ks * VK, rowB + m0);
        tma_load_2d(&S.Z[s][0][0], &tmapZ, &S.full[s], ks * VK, rowB + zmat * nbp + n0);
        if (++s == VSTAGES) { s = 0; ph ^= 1; }
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

  // ------------------------------------------------------------------ MMA warps
  reg_inc<V_MMA_REGS>();
  const int g = lane >> 2, t = lane & 3;
  // 4 x 2 warps; sub-partition = warp & 3 = (wm + wn) & 3: idle row blocks of ragged output blocks are
  // spread over the DMMA pipes
  const int wn = warp >> 2;
  const int wm = ((warp & 3) - wn) & 3;
  const int sw = (g & 3) << 2;
  int s = 0;
  uint32_t ph = 0;

  for (int it = 0;; ++it) {
    const int slot = it & (VQ - 1);
    mbar_wait(&S.qfull[slot], (it / VQ) & 1);
    const VxcSlot sl = S.q[slot];
    __syncwarp();
    if (lane == 0) mbar_arrive(&S.qempty[slot]);
    if (sl.nks < 0) break;
    const int nbe = sl.nbe, m0 = sl.m0, n0 = sl.n0;
    const int mi_cnt = min(4, max(0, (nbe - m0 - wm * 32 + 7) / 8));
    const int ni_cnt = min(4, max(0, (nbe - n0 - wn * 32 + 7) / 8));
    // symmetric M: a warp tile entirely above the diagonal is carried by its mirror
    const bool above = sl.diag && (n0 + wn * 32 > m0 + wm * 32 + 31);
    const bool active = mi_cnt > 0 && ni_cnt > 0 && !above;
    const int var = (mi_cnt > 2 ? 2 : 0) | (ni_cnt > 2 ? 1 : 0);
    // global AO indices of the warp's 32 rows / 32 columns, lane = row (column): requested now, read
    // through shuffles in the scatter, so their latency hides behind the K loop
    const int* __restrict__ ao = pv.task_ao + sl.ao_off;
    int ao_r = -1, ao_c = -1;
    {
      const int r = m0 + wm * 32 + lane, c = n0 + wn * 32 + lane;
      if (r < nbe) ao_r = __ldg(ao + r);
      if (c < nbe) ao_c = __ldg(ao + c);
    }

    double acc[4][4][2];
#pragma unroll
    for (int mi = 0; mi < 4; ++mi)
#pragma unroll
      for (int ni = 0; ni < 4; ++ni) acc[mi][ni][0] = acc[mi][ni][1] = 0.;

    for (int ks = 0; ks < sl.nks; ++ks) {
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
    if (!active) continue;

    // scatter: VXC_sub = M + M^T, only the lower triangle of the full matrix is accumulated
    int gm[4], gn[4][2];
#pragma unroll
    for (int mi = 0; mi < 4; ++mi) gm[mi] = __shfl_sync(0xffffffffu, ao_r, mi * 8 + g);
#pragma unroll
    for (int ni = 0; ni < 4; ++ni)
#pragma unroll
      for (int j = 0; j < 2; ++j) gn[ni][j] = __shfl_sync(0xffffffffu, ao_c, ni * 8 + 2 * t + j);
#pragma unroll
    for (int mi = 0; mi < 4; ++mi)
#pragma unroll
      for (int ni = 0; ni < 4; ++ni) {
        if (mi >= mi_cnt || ni >= ni_cnt) continue;
        const int mu = m0 + wm * 32 + mi * 8 + g;
        if (gm[mi] < 0) continue;  // mu >= nbe
#pragma unroll
        for (int j = 0; j < 2; ++j) {
          const int nu = n0 + wn * 32 + ni * 8 + 2 * t + j;
          if (gn[ni][j] < 0) continue;  // nu >= nbe
          double v = acc[mi][ni][j];
          if (sym) {
            if (nu > mu) continue;  // M symmetric: the mirror entry carries it
            v *= 2.;
          } else if (mu == nu) {
            v *= 2.;
          }
          const int hi = max(gm[mi], gn[ni][j]), lo = min(gm[mi], gn[ni][j]);
          atomicAdd(VXC + (size_t)lo * ldv + hi, v);
        }
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

cudaError_t launch_vxc(const CUtensorMap& tmapA, const CUtensorMap& tmapZ, const PlanView& pv, const VxcItem* items,
                       int nitems, int* counter, int nsm, int zmat, int nmat, bool sym, double* VXC, int ldv,
                       cudaStream_t s) {
  if (nitems <= 0 || nsm <= 0) return cudaSuccess;
  // the attribute is per device and cheap to set: no process-wide "done" flag
  cudaError_t e = cudaFuncSetAttribute(vxc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)VXC_SMEM_BYTES);
  if (e != cudaSuccess) return e;
  // two CTAs of 97 KB each must be co-resident: ask for the largest shared-memory carveout (the default
  // heuristic sizes the carveout for ONE block)
  e = cudaFuncSetAttribute(vxc_kernel, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
  if (e != cudaSuccess) return e;
  if (std::getenv("GAUXC_B200_DEBUG")) {
    static bool said = false;
    if (!said) {
      int nb = 0;
      cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, vxc_kernel, V_THREADS, VXC_SMEM_BYTES);
      std::fprintf(stderr, "[gauxc_b200] vxc_kernel: %d resident CTAs per SM (%zu B shared each)\n", nb, VXC_SMEM_BYTES);
      said = true;
    }
  }
  int ncta = 2 * nsm;  // two co-resident CTAs per SM
  ncta = ncta < nitems ? ncta : nitems;
  vxc_kernel<<<ncta, V_THREADS, VXC_SMEM_BYTES, s>>>(tmapA, tmapZ, pv, items, nitems, counter, zmat, nmat,
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
