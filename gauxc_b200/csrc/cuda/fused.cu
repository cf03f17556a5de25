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
// One persistent CTA per SM pulls tiles from a device-side queue (atomic counter, task order: the
// tiles in flight share P_sub / VXC regions in L2; no static partition to go out of balance) and
// broadcasts them to its roles through a 4-slot shared-memory ring.  Five roles, 20 warps:
//   producer (4 warps): TMA box loads of B^T (16 basis rows x 128 points) + LDGSTS gather of the
//                       matching 16 x 64 block of P through the task's AO map (4 rows per warp),
//                       5-stage mbarrier ring
//   MMA      (8 warps): 128 x 64 chunk of X on the DMMA pipe (m8n8k4); warp tile 64 x 16 with the two
//                       row halves of a column strip on the SAME SM sub-partition, so ragged tiles
//                       (npts < 128) load the four DMMA pipes evenly.  Each finished chunk is handed
//                       to the density warps through shared memory
//   density  (4 warps): lane = 4 consecutive grid points, warp = every 4th basis row:
//                       rho += X.B, grad rho += X.dB with 256-bit streaming loads
//   zmat     (4 warps): functional, weight scaling, EXC/N_EL tile partials (thread = point), then
//                       Z = 1/2 vrho B + 2 vgamma (grad rho . dB) with 256-bit loads/stores, rows walked
//                       in reverse so the most recently streamed rows are still in L2
// so the HBM-bound streams of the density and Z stages run underneath the DMMA work of the next
// chunk / next tile instead of in kernels of their own, and X never leaves the SM.
#include "kernels.cuh"
#include "ptx.cuh"
#include "xc_functionals.cuh"
#include "xc_functionals_pol_gga.cuh"

// GXB_KNOCKOUT (diagnostic builds only, results are garbage): bit 0 removes the density / Z global
// streams, bit 1 the P gather, bit 2 the B^T bulk loads -- to time what each costs the DMMA pipe.
#ifndef GXB_KNOCKOUT
#define GXB_KNOCKOUT 0
#endif
// stages of B^T prefetched into L2 ahead of the ring on the first pass over a tile's rows (0: none)
#ifndef GXB_FPF
#define GXB_FPF 0
#endif
// GXB_TIMING (diagnostic builds): MMA warp 0 of CTA 0 accumulates the clocks it spends waiting on each barrier
// and prints them when the launch ends
#ifndef GXB_TIMING
#define GXB_TIMING 0
#endif
#if GXB_TIMING
#include <cstdio>
#define GXB_T0() const long long _t0 = clock64()
#define GXB_T1(acc) acc += clock64() - _t0
#else
#define GXB_T0()
#define GXB_T1(acc)
#endif

namespace gxb {

namespace {

constexpr int FK = 16;       // basis rows (K) per pipeline stage
constexpr int FN = 64;       // columns of X per chunk
constexpr int FSTAGES = 5;
constexpr int TQ = 4;         // tile-queue ring slots
constexpr int P_LD = FN + 4;   // (ld mod 16) == 4: conflict-free DMMA B-fragment loads
constexpr int X_LD = TP + 2;   // conflict-free C-fragment stores, rows stay 16-byte aligned
constexpr int MMA_WARPS = 8, DEN_WARPS = 4, Z_WARPS = 4, PROD_WARPS = 4;
constexpr int MMA_THREADS = MMA_WARPS * 32, DEN_THREADS = DEN_WARPS * 32, Z_THREADS = Z_WARPS * 32;
// warps 0-7 MMA, 8-11 density, 12-15 functional+Z, 16-19 producers (one warpgroup per role family so
// that setmaxnreg can move registers from the producers to the MMA warps)
constexpr int PROD_THREADS = PROD_WARPS * 32;
constexpr int FUSED_THREADS = MMA_THREADS + DEN_THREADS + Z_THREADS + PROD_THREADS;
// launch allocation 20 warps x 96; after re-partitioning 8 x 112 (MMA) + 4 x 112 (density) + 4 x 96 (Z)
// + 4 x 48 (producers) (must not exceed it)
constexpr int MMA_REGS = 112, DEN_REGS = 112, PROD_REGS = 48;

struct FusedSmem {
  double A[FSTAGES][FK][TP];    // TMA destination, dense + global XOR swizzle
  double P[FSTAGES][FK][P_LD];
  double X[FN][X_LD];
  double dpart[DEN_WARPS][4][TP];  // per density warp partial sums
  double den[4][TP];
  double fac[2][8][TP];            // 1/2 w vrho, 2 w vsigma grad rho per point (UKS GGA: s and z channel)
  double red[2][2][Z_WARPS];
  uint64_t full[FSTAGES], empty[FSTAGES];
  uint64_t xfull, xempty, denfull, denempty;
  uint64_t tqfull[TQ], tqempty[TQ];
  int tq[TQ];  // tile index inside the batch, -1 = queue drained
};
constexpr size_t FUSED_SMEM_BYTES = sizeof(FusedSmem) + 1024;
static_assert(FUSED_SMEM_BYTES <= 232448, "fused kernel shared memory");

// L2 priorities of the two streaming passes over a tile's B / dB rows: the density pass marks them
// evict_last (the Z pass of the same tile re-reads them a few hundred microseconds later), the Z
// pass reads and writes evict_first (nothing touches those lines again before the VXC kernel).
#ifndef GXB_L2_HINTS
#define GXB_L2_HINTS 1
#endif
__device__ __forceinline__ void ldg_stream(double (&v)[4], const double* p, uint64_t pol) {
  if (GXB_KNOCKOUT & 1) { v[0] = v[1] = v[2] = v[3] = 1.; return; }
  if (GXB_L2_HINTS) ldg256_stream_hint(v, p, pol);
  else ldg256_stream(v, p);
}
__device__ __forceinline__ void stg_stream(double* p, const double (&v)[4], uint64_t pol) {
  if (GXB_KNOCKOUT & 1) { if (v[0] == 1.2345e-300) stg256(p, v); return; }
  if (GXB_L2_HINTS) stg256_hint(p, v, pol);
  else stg256(p, v);
}
__device__ __forceinline__ void lds4(double (&v)[4], const double* p) {
  const double2 a = *reinterpret_cast<const double2*>(p);
  const double2 b = *reinterpret_cast<const double2*>(p + 2);
  v[0] = a.x; v[1] = a.y; v[2] = b.x; v[3] = b.y;
}
__device__ __forceinline__ void sts4(double* p, const double (&v)[4]) {
  *reinterpret_cast<double2*>(p) = make_double2(v[0], v[1]);
  *reinterpret_cast<double2*>(p + 2) = make_double2(v[2], v[3]);
}

// One pipeline stage (16 K rows) of a warp's 64 x 16 tile restricted to its first MI row blocks.
// Straight-line, UNPREDICATED DMMAs: a predicated mma.sync costs a WARPSYNC per instruction, so
// ragged tiles dispatch (warp-uniformly, once per stage) to the variant with MI rounded up to even;
// rows / columns beyond the tile multiply stale-but-finite or zero-filled operands and are never read.
// KB..KE: the 4-row K steps of the stage this warp takes (all four, or one half in the split-K mode
// of short tiles).
// PAIRED (GGA kernels): the 8 x 8 blocks of a warp tile are laid over the points / columns in PAIRS -- row g of blocks
// 2q and 2q + 1 are the adjacent points 16 q + 2 g and 16 q + 2 g + 1, column g of the two column blocks the adjacent columns
// 2 g and 2 g + 1 of the warp's strip -- so one LDS.128 fetches the fragments of two blocks (6 fragment loads per K step
// instead of 10; the XOR swizzle of the tile rows only touches bits 2-3 of a column and keeps pairs together).  Ragged tiles
// then skip whole pairs (16 points), which is what the even-count variants do anyway; the LDA kernel keeps the plain
// mapping, whose exact odd counts in the split-K halves are worth more there.
template <int MI, int KB = 0, int KE = 4, bool PAIRED = false>
__device__ __forceinline__ void mma_stage(double (&acc)[8][2][2], const double* __restrict__ as,
                                          const double* __restrict__ ps, int a_ev, int a_od, int ap) {
#pragma unroll
  for (int kk = KB; kk < KE; ++kk) {
    double a[MI], b[2];
    if (PAIRED) {
      static_assert(!PAIRED || (MI % 2 == 0), "paired blocks come in twos");
#pragma unroll
      for (int q = 0; q < MI / 2; ++q) {
        const double2 v = *reinterpret_cast<const double2*>(as + kk * 4 * ap + a_ev + q * 16);
        a[2 * q] = v.x;
        a[2 * q + 1] = v.y;
      }
      const double2 w = *reinterpret_cast<const double2*>(ps + kk * 4 * P_LD);
      b[0] = w.x;
      b[1] = w.y;
    } else {
#pragma unroll
      for (int mi = 0; mi < MI; ++mi) a[mi] = as[kk * 4 * ap + ((mi & 1) ? a_od : a_ev) + (mi & ~1) * 8];
#pragma unroll
      for (int ni = 0; ni < 2; ++ni) b[ni] = ps[kk * 4 * P_LD + ni * 8];
    }
#pragma unroll
    for (int mi = 0; mi < MI; ++mi)
#pragma unroll
      for (int ni = 0; ni < 2; ++ni) dmma(acc[mi][ni][0], acc[mi][ni][1], a[mi], b[ni]);
  }
}

// SPIN = 0: RKS (P = P_alpha, factor 2 inside).  UKS runs the kernel twice per batch:
// SPIN = 1 with P = Ps = P_alpha + P_beta only stores rho_s (GGA: and grad n) per point (uks_den, one array
// of uks_stride doubles per component); SPIN = 2 with P = Pz forms rho_z (grad M_z), rho_+- = (rho_s +- rho_z)/2
// (and the three gammas), evaluates the polarised functional and writes BOTH Z_s and Z_z (the two matrices
// after B / dB) -- eval_uvvar_{lda,gga}_uks / eval_zmat_{lda,gga}_vxc_uks of the reference host driver
// (reference_local_host_work_driver.cxx:166-188, 270-328, 607-634, 715-773; X factor 1.0, driver :387-396).
// func.nkern == 0: density only (integrate_den): no functional, no Z pass.
// DUAL: the functional contains kernels evaluated with dual numbers (B88, LYP; every polarised GGA) -- kept out
// of the instantiations that do not need them (xc_functionals.cuh)
// XOUT (EXC gradient, eval_xmat of reference_replicated_xc_host_integrator_exc_grad.hpp:370-373): the kernel only
// forms X = 2 A P_sub for ONE matrix A of the tile (B or one of dB/dx, dB/dy, dB/dz) and the density warps write it
// to another matrix slot of the tile instead of contracting it -- the gradient kernel (exc_grad.cu) needs X itself.
// The two slots travel in uks_stride (unused for SPIN == 0): low 16 bits = A, next 16 bits = X slot, bit 32 set = UKS
// (eval_xmat factor 1.0 instead of the RKS 2.0).
template <bool GGA, int SPIN, bool DUAL, bool XOUT = false>
__global__ void __launch_bounds__(FUSED_THREADS, 1)
fused_xmat_den_zmat_kernel(const __grid_constant__ TmapSet tmaps, PlanView pv,
                           const DevTile* __restrict__ tiles, int ntiles, int* __restrict__ counter,
                           double* __restrict__ ws,
                           const double* __restrict__ P, int ldp, FunctionalDesc func,
                           double* __restrict__ exc_part, double* __restrict__ nel_part,
                           int part_off, double* __restrict__ uks_den, size_t uks_stride) {
  // no pointer arithmetic on the base: the compiler must see shared-space accesses (LDS/STS, not
  // generic LD/ST) in the fragment loads
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  FusedSmem& S = *reinterpret_cast<FusedSmem*>(smem_raw);
  if (threadIdx.x == 0 && (smem_u32(smem_raw) & 127u)) __trap();  // TMA destinations need 128 B

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  // consumer side of the tile queue: iteration `it` reads slot it % TQ
  auto next_tile = [&](int it) {
    const int slot = it & (TQ - 1);
    mbar_wait(&S.tqfull[slot], (it / TQ) & 1);
    const int idx = S.tq[slot];
    mbar_arrive(&S.tqempty[slot]);
    return idx;
  };

  if (tid == 0) {
    for (int s = 0; s < FSTAGES; ++s) {
      mbar_init(&S.full[s], PROD_THREADS);  // producer lanes (cp.async arrivals) + TMA bytes
      mbar_init(&S.empty[s], MMA_WARPS);  // one elected arrive per MMA warp
    }
    mbar_init(&S.xfull, MMA_THREADS);
    mbar_init(&S.xempty, DEN_THREADS);
    mbar_init(&S.denfull, DEN_THREADS);
    mbar_init(&S.denempty, Z_THREADS);
    for (int i = 0; i < TQ; ++i) {
      mbar_init(&S.tqfull[i], 1);
      mbar_init(&S.tqempty[i], MMA_THREADS + DEN_THREADS + Z_THREADS + PROD_THREADS - 32);
    }
    mbar_fence_init();
  }
  __syncthreads();

  if (warp < MMA_WARPS) {
    // ------------------------------------------------------------------ MMA warps
    reg_inc<MMA_REGS>();
    const int g = lane >> 2, t = lane & 3;
    const int wm = warp >> 2;  // row half (64 points)
    const int wn = warp & 3;   // column strip (16 columns) == SM sub-partition
    int s = 0;
    uint32_t ph = 0, xph = 0;
    // swizzled column of point (wm*64 + mi*8 + g) in a row with (row & 3) == t:
    // a0 ^ (mi << 3), i.e. a0 + 8*mi for even mi and (a0 ^ 8) + 8*(mi - 1) for odd mi
    constexpr bool PAIRED = GGA;  // fragment layout of mma_stage
    const int a_ev_full = (wm * 64 + (PAIRED ? 2 * g : g)) ^ (t << 2), a_ev_short = (PAIRED ? 2 * g : g) ^ (t << 2);

    long long w_tq = 0, w_full = 0, w_full_first = 0, w_xempty = 0, w_store = 0, n_stage = 0, n_tile = 0;
#if GXB_TIMING
    long long b_clk[4] = {0, 0, 0, 0}, b_stage[4] = {0, 0, 0, 0};  // per tile-fill bucket (<= 32, 64, 96, 128 points)
    long long b_wait[4] = {0, 0, 0, 0};                              // of which: waiting on any barrier
#endif
    const long long t_begin = clock64();
    for (int it = 0;; ++it) {
      int tile_idx;
      { GXB_T0(); tile_idx = next_tile(it); GXB_T1(w_tq); }
      if (tile_idx < 0) break;
      ++n_tile;
      const DevTile tile = tiles[tile_idx];
      const int nbe = tile.nbe;
#if GXB_TIMING
      const long long t_tile = clock64(), st_tile = n_stage;
      const long long w_tile = w_full + w_full_first + w_xempty + w_store;
      const int bucket = (tile.npts - 1) >> 5;
#endif
      const int nk = pad16(nbe) / FK;
      const int nn = (nbe + FN - 1) / FN;
      // Short tiles (<= 64 points) would leave the wm = 1 warps idle and every sub-partition with a
      // single MMA warp: there both row halves work on the SAME 64 rows and split each stage's K
      // steps (wm = 0: rows 0-7 of the stage, wm = 1: rows 8-15); the two partial X are added in
      // shared memory when the chunk is handed over.
      const bool split = tile.npts <= 64;
      const int a_ev = split ? a_ev_short : a_ev_full, a_od = a_ev ^ 8;
      const int mi_cnt = split ? (tile.npts + 7) / 8 : min(8, max(0, (tile.npts - wm * 64 + 7) / 8));
      const int ap = tile_width(tile.npts);  // shared-memory pitch of the B^T box of this tile
      for (int c = 0; c < nn; ++c) {
        const int ni_cnt = min(2, max(0, (nbe - c * FN - wn * 16 + 7) / 8));
        // LDA: only rho = sum_n B_n X_n is needed, a quadratic form in B, so the K loop of column
        // chunk c stops at the diagonal block (P' = lower triangle of P with a halved diagonal)
        const int nkc = GGA ? nk : min(nk, (FN / FK) * (c + 1));
        const bool active = mi_cnt > 0 && ni_cnt > 0;
        const int mi_var = (mi_cnt + 1) >> 1;
        double acc[8][2][2];
#pragma unroll
        for (int mi = 0; mi < 8; ++mi)
#pragma unroll
          for (int ni = 0; ni < 2; ++ni) acc[mi][ni][0] = acc[mi][ni][1] = 0.;

        for (int ks = 0; ks < nkc; ++ks) {
          { GXB_T0(); mbar_wait(&S.full[s], ph); if (ks < FSTAGES) { GXB_T1(w_full_first); } else { GXB_T1(w_full); } }
          ++n_stage;
          const double* as = &S.A[s][0][0] + t * ap;
          const double* ps = &S.P[s][t][wn * 16 + (PAIRED ? 2 * g : g)];
          if (active) {
            if (!split) {
              switch (mi_var) {
                case 4: mma_stage<8, 0, 4, PAIRED>(acc, as, ps, a_ev, a_od, ap); break;
                case 3: mma_stage<6, 0, 4, PAIRED>(acc, as, ps, a_ev, a_od, ap); break;
                case 2: mma_stage<4, 0, 4, PAIRED>(acc, as, ps, a_ev, a_od, ap); break;
                default: mma_stage<2, 0, 4, PAIRED>(acc, as, ps, a_ev, a_od, ap); break;
              }
            } else if (GGA) {
              // split-K halves, row blocks rounded up to even (the GGA kernel measures faster with
              // the smaller set of variants: taxol fused 123 ms against 132 ms with exact counts)
              if (wm == 0) {
                switch (mi_var) {
                  case 4: mma_stage<8, 0, 2, PAIRED>(acc, as, ps, a_ev, a_od, ap); break;
                  case 3: mma_stage<6, 0, 2, PAIRED>(acc, as, ps, a_ev, a_od, ap); break;
                  case 2: mma_stage<4, 0, 2, PAIRED>(acc, as, ps, a_ev, a_od, ap); break;
                  default: mma_stage<2, 0, 2, PAIRED>(acc, as, ps, a_ev, a_od, ap); break;
                }
              } else {
                switch (mi_var) {
                  case 4: mma_stage<8, 2, 4, PAIRED>(acc, as, ps, a_ev, a_od, ap); break;
                  case 3: mma_stage<6, 2, 4, PAIRED>(acc, as, ps, a_ev, a_od, ap); break;
                  case 2: mma_stage<4, 2, 4, PAIRED>(acc, as, ps, a_ev, a_od, ap); break;
                  default: mma_stage<2, 2, 4, PAIRED>(acc, as, ps, a_ev, a_od, ap); break;
                }
              }
            } else if (wm == 0) {
              // LDA kernel (large molecules on coarse grids: tasks of 20-60 points dominate): exact
              // row-block count in the split-K halves (ubiquitin fused 339 -> 326 ms)
              switch (mi_cnt) {
                case 8: mma_stage<8, 0, 2>(acc, as, ps, a_ev, a_od, ap); break;
                case 7: mma_stage<7, 0, 2>(acc, as, ps, a_ev, a_od, ap); break;
                case 6: mma_stage<6, 0, 2>(acc, as, ps, a_ev, a_od, ap); break;
                case 5: mma_stage<5, 0, 2>(acc, as, ps, a_ev, a_od, ap); break;
                case 4: mma_stage<4, 0, 2>(acc, as, ps, a_ev, a_od, ap); break;
                case 3: mma_stage<3, 0, 2>(acc, as, ps, a_ev, a_od, ap); break;
                case 2: mma_stage<2, 0, 2>(acc, as, ps, a_ev, a_od, ap); break;
                default: mma_stage<1, 0, 2>(acc, as, ps, a_ev, a_od, ap); break;
              }
            } else {
              switch (mi_cnt) {
                case 8: mma_stage<8, 2, 4>(acc, as, ps, a_ev, a_od, ap); break;
                case 7: mma_stage<7, 2, 4>(acc, as, ps, a_ev, a_od, ap); break;
                case 6: mma_stage<6, 2, 4>(acc, as, ps, a_ev, a_od, ap); break;
                case 5: mma_stage<5, 2, 4>(acc, as, ps, a_ev, a_od, ap); break;
                case 4: mma_stage<4, 2, 4>(acc, as, ps, a_ev, a_od, ap); break;
                case 3: mma_stage<3, 2, 4>(acc, as, ps, a_ev, a_od, ap); break;
                case 2: mma_stage<2, 2, 4>(acc, as, ps, a_ev, a_od, ap); break;
                default: mma_stage<1, 2, 4>(acc, as, ps, a_ev, a_od, ap); break;
              }
            }
          }
          // release the stage: one arrive per warp (256 per-thread arrives on one mbarrier would
          // serialise in the shared-memory atomic unit every stage)
          __syncwarp();
          if (lane == 0) mbar_arrive(&S.empty[s]);
          if (++s == FSTAGES) { s = 0; ph ^= 1; }
        }
        // hand the chunk of X to the density warps
        { GXB_T0(); mbar_wait(&S.xempty, xph ^ 1); GXB_T1(w_xempty); }
        GXB_T0();
        if (PAIRED) {
          // element (mi, ni, j) of the warp tile is point 16 (mi >> 1) + 2 g + (mi & 1), column 2 (2 t + j) + ni
          const int prow = (split ? 0 : wm * 64) + 2 * g;
          if (!split || wm == 1) {
#pragma unroll
            for (int q = 0; q < 4; ++q)
#pragma unroll
              for (int ni = 0; ni < 2; ++ni)
#pragma unroll
                for (int j = 0; j < 2; ++j)
                  *reinterpret_cast<double2*>(&S.X[wn * 16 + 2 * (2 * t + j) + ni][prow + 16 * q]) =
                      make_double2(acc[2 * q][ni][j], acc[2 * q + 1][ni][j]);
          }
          if (split) {
            // split-K: the wm = 1 partial went to shared memory first, wm = 0 adds its own on top
            named_bar_sync(3, MMA_THREADS);
            if (wm == 0) {
#pragma unroll
              for (int q = 0; q < 4; ++q)
#pragma unroll
                for (int ni = 0; ni < 2; ++ni)
#pragma unroll
                  for (int j = 0; j < 2; ++j) {
                    double2* dst = reinterpret_cast<double2*>(&S.X[wn * 16 + 2 * (2 * t + j) + ni][prow + 16 * q]);
                    const double2 o = *dst;
                    *dst = make_double2(o.x + acc[2 * q][ni][j], o.y + acc[2 * q + 1][ni][j]);
                  }
            }
          }
        } else if (!split) {
#pragma unroll
          for (int mi = 0; mi < 8; ++mi)
#pragma unroll
            for (int ni = 0; ni < 2; ++ni)
#pragma unroll
              for (int j = 0; j < 2; ++j)
                S.X[wn * 16 + ni * 8 + 2 * t + j][wm * 64 + mi * 8 + g] = acc[mi][ni][j];
        } else {
          // split-K: the wm = 1 partial goes to shared memory first, wm = 0 adds its own on top
          if (wm == 1) {
#pragma unroll
            for (int mi = 0; mi < 8; ++mi)
#pragma unroll
              for (int ni = 0; ni < 2; ++ni)
#pragma unroll
                for (int j = 0; j < 2; ++j) S.X[wn * 16 + ni * 8 + 2 * t + j][mi * 8 + g] = acc[mi][ni][j];
          }
          named_bar_sync(3, MMA_THREADS);
          if (wm == 0) {
#pragma unroll
            for (int mi = 0; mi < 8; ++mi)
#pragma unroll
              for (int ni = 0; ni < 2; ++ni)
#pragma unroll
                for (int j = 0; j < 2; ++j) S.X[wn * 16 + ni * 8 + 2 * t + j][mi * 8 + g] += acc[mi][ni][j];
          }
        }
        mbar_arrive(&S.xfull);
        xph ^= 1;
        GXB_T1(w_store);
      }
#if GXB_TIMING
      b_clk[bucket] += clock64() - t_tile;
      b_stage[bucket] += n_stage - st_tile;
      b_wait[bucket] += w_full + w_full_first + w_xempty + w_store - w_tile;
#endif
    }
#if GXB_TIMING
    if (blockIdx.x == 0 && warp == 0 && lane == 0) {
      const double tot = (double)(clock64() - t_begin);
      printf("[fused timing] tiles %lld stages %lld total %.0f clk: wait tile-queue %.1f%% full(first 5 stages of a chunk) %.1f%% "
             "full(later) %.1f%% xempty %.1f%% X hand-over %.1f%%; per stage %.0f clk\n",
             n_tile, n_stage, tot, 100. * w_tq / tot, 100. * w_full_first / tot, 100. * w_full / tot, 100. * w_xempty / tot,
             100. * w_store / tot, tot / (double)(n_stage ? n_stage : 1));
      printf("[fused fill] clk per stage by tile fill: <=32 pts %lld stages %.0f clk | <=64 %lld %.0f | <=96 %lld %.0f | <=128 %lld %.0f\n",
             b_stage[0], b_stage[0] ? (double)b_clk[0] / b_stage[0] : 0., b_stage[1], b_stage[1] ? (double)b_clk[1] / b_stage[1] : 0.,
             b_stage[2], b_stage[2] ? (double)b_clk[2] / b_stage[2] : 0., b_stage[3], b_stage[3] ? (double)b_clk[3] / b_stage[3] : 0.);
      printf("[fused fillwait] barrier-wait clk per stage by tile fill: %.0f | %.0f | %.0f | %.0f\n",
             b_stage[0] ? (double)b_wait[0] / b_stage[0] : 0., b_stage[1] ? (double)b_wait[1] / b_stage[1] : 0.,
             b_stage[2] ? (double)b_wait[2] / b_stage[2] : 0., b_stage[3] ? (double)b_wait[3] / b_stage[3] : 0.);
    }
#else
    (void)w_tq; (void)w_full; (void)w_full_first; (void)w_xempty; (void)w_store; (void)n_stage; (void)n_tile; (void)t_begin;
#endif
  } else if (warp < MMA_WARPS + DEN_WARPS) {
    // ------------------------------------------------------------------ density warps
    reg_inc<DEN_REGS>();
    const int p = tid - MMA_THREADS;       // point owned in the cross-warp reduction
    const int dw = warp - MMA_WARPS;       // rows with (row & 3) == dw
    const int p4 = lane * 4;               // 4 consecutive points
    const int cofs = p4 ^ (dw << 2);       // their (swizzled) column in every row of this warp
    uint32_t xph = 0, dph = 0;
    const uint64_t pol_keep = l2_policy_evict_last();
    for (int it = 0;; ++it) {
      const int tile_idx = next_tile(it);
      if (tile_idx < 0) break;
      const DevTile tile = tiles[tile_idx];
      const int nbe = tile.nbe;
      const size_t ms = (size_t)pad16(nbe) * TP;
      const double* __restrict__ Bt = ws + tile.ws_off + cofs;
      const int nn = (nbe + FN - 1) / FN;
      double r0[4] = {0., 0., 0., 0.}, r1[4] = {0., 0., 0., 0.}, r2[4] = {0., 0., 0., 0.},
             r3[4] = {0., 0., 0., 0.};
      const bool lane_on = p4 < tile_width(tile.npts);  // columns beyond the tile width do not exist
      if (XOUT) {
        double* __restrict__ Xo = ws + tile.ws_off + (size_t)((uks_stride >> 16) & 0xffff) * ms + cofs;
        const double xfac = ((uks_stride >> 32) & 1) ? 1. : 2.;  // eval_xmat fac: RKS 2, UKS 1
        for (int c = 0; c < nn; ++c) {
          const int n0 = c * FN;
          const int ncols = lane_on ? min(FN, nbe - n0) : 0;
          mbar_wait(&S.xfull, xph);
          for (int n = dw; n < ncols; n += 4) {
            double x[4];
            lds4(x, &S.X[n][p4]);
#pragma unroll
            for (int j = 0; j < 4; ++j) x[j] *= xfac;
            stg256(Xo + (size_t)(n0 + n) * TP, x);
          }
          mbar_arrive(&S.xempty);
          xph ^= 1;
        }
        continue;
      }
      for (int c = 0; c < nn; ++c) {
        const int n0 = c * FN;
        const int ncols = lane_on ? min(FN, nbe - n0) : 0;
        mbar_wait(&S.xfull, xph);
        int n = dw;
        for (; n + 4 < ncols; n += 8) {
          double x[2][4], b0[2][4], b1[2][4], b2[2][4], b3[2][4];
#pragma unroll
          for (int u = 0; u < 2; ++u) {
            const double* src = Bt + (size_t)(n0 + n + 4 * u) * TP;
            ldg_stream(b0[u], src, pol_keep);
            if (GGA) {
              ldg_stream(b1[u], src + ms, pol_keep);
              ldg_stream(b2[u], src + 2 * ms, pol_keep);
              ldg_stream(b3[u], src + 3 * ms, pol_keep);
            }
            lds4(x[u], &S.X[n + 4 * u][p4]);
          }
#pragma unroll
          for (int u = 0; u < 2; ++u)
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              r0[j] = fma(x[u][j], b0[u][j], r0[j]);
              if (GGA) {
                r1[j] = fma(x[u][j], b1[u][j], r1[j]);
                r2[j] = fma(x[u][j], b2[u][j], r2[j]);
                r3[j] = fma(x[u][j], b3[u][j], r3[j]);
              }
            }
        }
        if (n < ncols) {
          double x[4], b0[4], b1[4], b2[4], b3[4];
          const double* src = Bt + (size_t)(n0 + n) * TP;
          ldg_stream(b0, src, pol_keep);
          if (GGA) {
            ldg_stream(b1, src + ms, pol_keep);
            ldg_stream(b2, src + 2 * ms, pol_keep);
            ldg_stream(b3, src + 3 * ms, pol_keep);
          }
          lds4(x, &S.X[n][p4]);
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            r0[j] = fma(x[j], b0[j], r0[j]);
            if (GGA) {
              r1[j] = fma(x[j], b1[j], r1[j]);
              r2[j] = fma(x[j], b2[j], r2[j]);
              r3[j] = fma(x[j], b3[j], r3[j]);
            }
          }
        }
        mbar_arrive(&S.xempty);
        xph ^= 1;
      }
      // cross-warp reduction in fixed order
      sts4(&S.dpart[dw][0][p4], r0);
      if (GGA) {
        sts4(&S.dpart[dw][1][p4], r1);
        sts4(&S.dpart[dw][2][p4], r2);
        sts4(&S.dpart[dw][3][p4], r3);
      }
      named_bar_sync(2, DEN_THREADS);
      mbar_wait(&S.denempty, dph ^ 1);
      // X carries the RKS factor 2 (eval_xmat fac = 2), the gradient another 2; the LDA path sums
      // the lower triangle of the quadratic form only (another 2)
#pragma unroll
      for (int qn = 0; qn < (GGA ? 4 : 1); ++qn) {
        const double v = (S.dpart[0][qn][p] + S.dpart[1][qn][p]) + (S.dpart[2][qn][p] + S.dpart[3][qn][p]);
        // RKS: X carries 2, the gradient another 2, the LDA triangle another 2; UKS: X factor 1.0
        S.den[qn][p] = (SPIN != 0 ? (GGA ? (qn == 0 ? 1. : 2.) : 2.) : ((qn == 0 && GGA) ? 2. : 4.)) * v;
      }
      mbar_arrive(&S.denfull);
      dph ^= 1;
      named_bar_sync(2, DEN_THREADS);  // dpart may be overwritten by the next tile
    }
  } else if (warp < MMA_WARPS + DEN_WARPS + Z_WARPS) {
    // ------------------------------------------------------------------ functional + Z warps
    const int p = tid - MMA_THREADS - DEN_THREADS;
    const int zw = p >> 5;
    const int p4 = lane * 4;
    const int cofs = p4 ^ (zw << 2);
    uint32_t dph = 0;
    const uint64_t pol_drop = l2_policy_evict_first();
    for (int it = 0;; ++it) {
      const int tile_idx = next_tile(it);
      if (tile_idx < 0) break;
      if (XOUT) continue;  // X is written by the density warps; no functional, no Z
      const DevTile tile = tiles[tile_idx];
      const int nbe = tile.nbe;
      const int nbp = pad16(nbe);
      const size_t ms = (size_t)nbp * TP;
      const double* __restrict__ Bt = ws + tile.ws_off + cofs;
      double* __restrict__ Z = ws + tile.ws_off + (GGA ? 4 : 1) * ms + cofs;
      double* __restrict__ Zz = Z + ms;  // UKS: Z_z follows Z_s

      mbar_wait(&S.denfull, dph);
      const double rho = S.den[0][p];
      double dx = 0., dy = 0., dz = 0.;
      if (GGA) { dx = S.den[1][p]; dy = S.den[2][p]; dz = S.den[3][p]; }
      mbar_arrive(&S.denempty);
      dph ^= 1;

      const bool ok = p < tile.npts;
      if (SPIN == 1) {  // UKS pass over Ps: keep rho_s (grad n) for the pass over Pz, nothing else to do
        if (ok) {
          uks_den[tile.pt_off + p] = rho;
          if (GGA) {
            uks_den[uks_stride + tile.pt_off + p] = dx;
            uks_den[2 * uks_stride + tile.pt_off + p] = dy;
            uks_den[3 * uks_stride + tile.pt_off + p] = dz;
          }
        }
        continue;
      }
      double a = 0., fx = 0., fy = 0., fz = 0., e_loc = 0., n_loc = 0.;
      double az = 0., gx = 0., gy = 0., gz = 0.;  // UKS GGA: factors of Z_z
      if (ok && SPIN == 2 && GGA && DUAL) {
        // eval_uvvar_gga_uks :270-328, weights :453-466, eval_zmat_gga_vxc_uks :715-773
        const double w = pv.w[tile.pt_off + p];
        const size_t ip = tile.pt_off + p;
        const double rho_s = uks_den[ip], nx = uks_den[uks_stride + ip], ny = uks_den[2 * uks_stride + ip],
                     nz = uks_den[3 * uks_stride + ip];
        const double rho_z = rho, mx = dx, my = dy, mz = dz;
        const double dn_sq = nx * nx + ny * ny + nz * nz, dm_sq = mx * mx + my * my + mz * mz,
                     dn_dm = nx * mx + ny * my + nz * mz;
        const double gpp = 0.25 * (dn_sq + dm_sq) + 0.5 * dn_dm, gpm = 0.25 * (dn_sq - dm_sq),
                     gmm = 0.25 * (dn_sq + dm_sq) - 0.5 * dn_dm;
        const XcOutPolGga xc = eval_functional_pol(func, 0.5 * (rho_s + rho_z), 0.5 * (rho_s - rho_z), gpp, gpm, gmm);
        const double eps = xc.eps * w;
        const double factp = 0.5 * (xc.va * w), factm = 0.5 * (xc.vb * w);
        const double vpp = xc.vaa * w, vpm = xc.vab * w, vmm = xc.vbb * w;
        const double g1 = 0.5 * (vpp + vpm + vmm), g2 = 0.5 * (vpp - vmm), g3 = 0.5 * (vpp - vpm + vmm);
        a = 0.5 * (factp + factm);
        az = 0.5 * (factp - factm);
        fx = g1 * nx + g2 * mx; fy = g1 * ny + g2 * my; fz = g1 * nz + g2 * mz;
        gx = g3 * mx + g2 * nx; gy = g3 * my + g2 * ny; gz = g3 * mz + g2 * nz;
        e_loc = eps * rho_s;
        n_loc = w * rho_s;
      } else if (ok && SPIN == 2) {
        const double w = pv.w[tile.pt_off + p];
        const double rho_s = uks_den[tile.pt_off + p], rho_z = rho;
        const XcOutPol xc = eval_functional_pol_lda(func, 0.5 * (rho_s + rho_z), 0.5 * (rho_s - rho_z));
        const double eps = xc.eps * w;  // host driver :453-457
        const double factp = 0.5 * (xc.va * w), factm = 0.5 * (xc.vb * w);
        a = 0.5 * (factp + factm);   // Z_s = a B
        fx = 0.5 * (factp - factm);  // Z_z = fx B
        e_loc = eps * rho_s;         // :490-497, total density
        n_loc = w * rho_s;
      } else if (ok) {
        const double w = pv.w[tile.pt_off + p];
        const double sigma = GGA ? dx * dx + dy * dy + dz * dz : 0.;
        const XcOut xc = eval_functional_t<DUAL>(func, rho, sigma);
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
      double(*fac)[TP] = S.fac[it & 1];
      fac[0][p] = a;
      if (GGA || SPIN == 2) fac[1][p] = fx;
      if (GGA) { fac[2][p] = fy; fac[3][p] = fz; }
      if (GGA && SPIN == 2) { fac[4][p] = az; fac[5][p] = gx; fac[6][p] = gy; fac[7][p] = gz; }
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
      if (func.nkern == 0) continue;  // integrate_den: nothing reads a Z
      // Z rows (pad rows are never read as valid output rows); warp zw owns rows = zw (mod 4),
      // last rows first: they were streamed most recently by the density warps
      double a4[4], fx4[4], fy4[4], fz4[4];
      lds4(a4, &fac[0][p4]);
      if (GGA || SPIN == 2) lds4(fx4, &fac[1][p4]);
      if (GGA) { lds4(fy4, &fac[2][p4]); lds4(fz4, &fac[3][p4]); }
      double az4[4], gx4[4], gy4[4], gz4[4];
      if (GGA && SPIN == 2) {
        lds4(az4, &fac[4][p4]); lds4(gx4, &fac[5][p4]); lds4(gy4, &fac[6][p4]); lds4(gz4, &fac[7][p4]);
      }
      int mu = ((nbe - 1 - zw) & ~3) + zw;  // largest row <= nbe-1 with (row & 3) == zw
      if (p4 >= tile_width(tile.npts)) mu = -1;  // columns beyond the tile width do not exist
      for (; mu - 4 >= 0; mu -= 8) {
        double b0[2][4], b1[2][4], b2[2][4], b3[2][4];
#pragma unroll
        for (int u = 0; u < 2; ++u) {
          const double* src = Bt + (size_t)(mu - 4 * u) * TP;
          ldg_stream(b0[u], src, pol_drop);
          if (GGA) {
            ldg_stream(b1[u], src + ms, pol_drop);
            ldg_stream(b2[u], src + 2 * ms, pol_drop);
            ldg_stream(b3[u], src + 3 * ms, pol_drop);
          }
        }
#pragma unroll
        for (int u = 0; u < 2; ++u) {
          double z[4];
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            z[j] = a4[j] * b0[u][j];
            if (GGA) {
              z[j] = fma(fx4[j], b1[u][j], z[j]);
              z[j] = fma(fy4[j], b2[u][j], z[j]);
              z[j] = fma(fz4[j], b3[u][j], z[j]);
            }
          }
          stg_stream(Z + (size_t)(mu - 4 * u) * TP, z, pol_drop);
          if (SPIN == 2) {  // Z_z from the same rows of B (dB)
            double zz[4];
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              if (GGA) {
                zz[j] = az4[j] * b0[u][j];
                zz[j] = fma(gx4[j], b1[u][j], zz[j]);
                zz[j] = fma(gy4[j], b2[u][j], zz[j]);
                zz[j] = fma(gz4[j], b3[u][j], zz[j]);
              } else {
                zz[j] = fx4[j] * b0[u][j];
              }
            }
            stg_stream(Zz + (size_t)(mu - 4 * u) * TP, zz, pol_drop);
          }
        }
      }
      if (mu >= 0) {
        double b0[4], b1[4], b2[4], b3[4], z[4];
        const double* src = Bt + (size_t)mu * TP;
        ldg_stream(b0, src, pol_drop);
        if (GGA) {
          ldg_stream(b1, src + ms, pol_drop);
          ldg_stream(b2, src + 2 * ms, pol_drop);
          ldg_stream(b3, src + 3 * ms, pol_drop);
        }
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          z[j] = a4[j] * b0[j];
          if (GGA) {
            z[j] = fma(fx4[j], b1[j], z[j]);
            z[j] = fma(fy4[j], b2[j], z[j]);
            z[j] = fma(fz4[j], b3[j], z[j]);
          }
        }
        stg_stream(Z + (size_t)mu * TP, z, pol_drop);
        if (SPIN == 2) {
          double zz[4];
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            if (GGA) {
              zz[j] = az4[j] * b0[j];
              zz[j] = fma(gx4[j], b1[j], zz[j]);
              zz[j] = fma(gy4[j], b2[j], zz[j]);
              zz[j] = fma(gz4[j], b3[j], zz[j]);
            } else {
              zz[j] = fx4[j] * b0[j];
            }
          }
          stg_stream(Zz + (size_t)mu * TP, zz, pol_drop);
        }
      }
    }
  } else {
    // ------------------------------------------------------------------ producer warps
    // warp pw gathers rows 4pw..4pw+3 of every 16-row stage of P; warp 0 also owns the tile queue
    // and the TMA loads of B^T
    reg_dec<PROD_REGS>();
    const int pw = warp - (MMA_WARPS + DEN_WARPS + Z_WARPS);
    int s = 0;
    uint32_t ph = 0;
    if (pw == 0 && lane < 4) tma_prefetch_desc(&tmaps.m[lane]);
    int popped = -2;  // lane 0 of warp 0: the NEXT tile, popped one tile ahead (-2: nothing popped yet)
    for (int it = 0;; ++it) {
      int tile_idx;
      if (pw == 0) {
        const int slot = it & (TQ - 1);
        mbar_wait(&S.tqempty[slot], ((it / TQ) & 1) ^ 1);
        tile_idx = -1;
        if (lane == 0) {
          tile_idx = popped == -2 ? atomicAdd(counter, 1) : popped;
          if (tile_idx >= ntiles) tile_idx = -1;
          S.tq[slot] = tile_idx;
          mbar_arrive(&S.tqfull[slot]);
          // pop the next tile now: the atomic's round trip and the tile record's first touch then overlap this
          // tile's loads instead of opening the next tile (the producer's lead over the MMA warps is only the
          // 5-stage ring)
          popped = tile_idx >= 0 ? atomicAdd(counter, 1) : -1;
          if (popped >= 0 && popped < ntiles) prefetch_l2(tiles + popped);
        }
        tile_idx = __shfl_sync(0xffffffffu, tile_idx, 0);
      } else {
        tile_idx = next_tile(it);
      }
      if (tile_idx < 0) break;
      const DevTile tile = tiles[tile_idx];
      const int nbe = tile.nbe;
      const int nk = pad16(nbe) / FK;
      const int nn = (nbe + FN - 1) / FN;
      const int* __restrict__ ao = pv.task_ao + tile.ao_off;
      const int rowB = (int)(tile.ws_off / TP) + (XOUT ? (int)(uks_stride & 0xffff) * pad16(nbe) : 0);
      const int W = tile_width(tile.npts);
      for (int c = 0; c < nn; ++c) {
        // lane -> columns (lane, lane + 32): each warp-wide LDGSTS covers 32 consecutive local AOs
        const int na = c * FN + lane, nb = na + 32;
        const bool va = na < nbe, vb = nb < nbe;
        const int ca = va ? __ldg(ao + na) : 0, cb = vb ? __ldg(ao + nb) : 0;
        // row bases of the NEXT stage are fetched while this stage is issued (the AO map lookup is a
        // dependent global load)
        const int kl = pw * 4 + (lane & 3);
        int ao_next = kl < nbe ? __ldg(ao + kl) : -1;
        const int nkc = GGA ? nk : min(nk, (FN / FK) * (c + 1));
        for (int ks = 0; ks < nkc; ++ks) {
          const int k0 = ks * FK;
          const long long rb_mine = ao_next >= 0 ? (long long)ao_next * ldp : -1;
          {
            const int kn = k0 + FK + kl;
            ao_next = (ks + 1 < nkc && kn < nbe) ? __ldg(ao + kn) : -1;
          }
          // first pass over the rows of B (chunk 0): they come from HBM, several ring depths of latency away --
          // pull the box of a later stage into L2 now (TMA prefetch: no shared memory, no completion)
          if (GXB_FPF > 0 && pw == 0 && lane == 0 && c == 0 && ks + GXB_FPF < nkc)
            tma_prefetch_2d(&tmaps.m[W / 32 - 1], 0, rowB + k0 + GXB_FPF * FK);
          mbar_wait(&S.empty[s], ph ^ 1);
          if (pw == 0) {
            // B^T: ONE tensor copy of 16 rows x the W columns the tile owns (one descriptor per
            // width; the box lands dense with pitch W, which the MMA warps address with)
            if (lane == 0 && !(GXB_KNOCKOUT & 4)) {
              mbar_expect_tx(&S.full[s], FK * W * sizeof(double));
              tma_load_2d(&S.A[s][0][0], &tmaps.m[W / 32 - 1], &S.full[s], 0, rowB + k0);
            }
          }
#pragma unroll
          for (int r = 0; r < 4; ++r) {
            const long long rb = __shfl_sync(0xffffffffu, rb_mine, r);
            const bool vr = rb >= 0;
            const double* src = P + (vr ? rb : 0);
            const int k = k0 + pw * 4 + r;  // LDA: rows k <= n only (see the MMA warps)
            if (GXB_KNOCKOUT & 2) continue;
            cp_async8_zfill(&S.P[s][pw * 4 + r][lane], src + ca, vr && va && (GGA || k <= na));
            cp_async8_zfill(&S.P[s][pw * 4 + r][lane + 32], src + cb, vr && vb && (GGA || k <= nb));
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

cudaError_t launch_fused(const TmapSet& tmapA, const PlanView& pv, const DevTile* tiles, int ntiles,
                         int* counter, int ncta, double* ws, const double* P, int ldp,
                         FunctionalDesc func, double* exc_part, double* nel_part, int part_off,
                         cudaStream_t s, int spin, double* uks_den, size_t uks_stride) {
  if (ncta <= 0 || ntiles <= 0) return cudaSuccess;
  ncta = ncta < ntiles ? ncta : ntiles;
  // the attribute is per device and cheap to set: no process-wide "done" flag
  auto launch = [&](auto kern) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)FUSED_SMEM_BYTES);
    if (e != cudaSuccess) return e;
    kern<<<ncta, FUSED_THREADS, FUSED_SMEM_BYTES, s>>>(tmapA, pv, tiles, ntiles, counter, ws, P, ldp, func,
                                                       exc_part, nel_part, part_off, uks_den, uks_stride);
    return cudaGetLastError();
  };
  if (spin == 3) return launch(fused_xmat_den_zmat_kernel<true, 0, false, true>);  // XOUT
  if (spin == 1 && func.is_gga) return launch(fused_xmat_den_zmat_kernel<true, 1, false>);
  if (spin == 2 && func.is_gga) return launch(fused_xmat_den_zmat_kernel<true, 2, true>);
  if (spin == 1) return launch(fused_xmat_den_zmat_kernel<false, 1, false>);
  if (spin == 2) return launch(fused_xmat_den_zmat_kernel<false, 2, false>);
  if (func.is_gga && functional_needs_dual(func)) return launch(fused_xmat_den_zmat_kernel<true, 0, true>);
  if (func.is_gga) return launch(fused_xmat_den_zmat_kernel<true, 0, false>);
  return launch(fused_xmat_den_zmat_kernel<false, 0, false>);
}

}  // namespace gxb
