// SSF (Stratmann-Scuseria-Frisch) molecular partition weights, thread = grid point.
//
// Arithmetic follows the HOST reference reference_ssf_weights_host
// (src/xc_integrator/local_work_driver/host/reference/weights.cxx:117-237), not the
// reference's CUDA kernel (cuda_ssf_1d.cu:24-128), whose algorithm differs:
//   * cutoff: r_parent < 0.5*(1-0.64)*dist_nearest  -> weight unchanged
//   * P_C = prod_{B != C} s(mu_CB) accumulated in ascending B, with the host's operand order
//     (B < C: g(mu_CB);  B > C: 1 - g(mu_BC)),  mu = (r_i - r_j) * (1 / RAB) with the reciprocal
//     precomputed on the host and g() multiplying by 1 / 0.64 = 1.5625: the kernel is FP64-ALU bound
//     and the host's two divisions per pair were half of its FP64 work; mu moves by at most an ulp,
//     which the continuous s(mu) turns into O(1e-16) of a weight (golden test: 1e-13 relative)
//   * w *= P_parent / sum_C P_C, sum in ascending C
// Instead of materialising natoms distances per point (the reference device path stores a
// natoms x npts scratch, xc_device_data.hpp:449-451) distances are recomputed, and every loop over
// atoms is cut off EXACTLY with the triangle inequality around an anchor atom A = the atom nearest
// to the tile's first point (found by a block-wide scan; tiles are spatially compact, so r_A is
// small for every point of the tile), whose neighbours arrive sorted by R_AB (nbr_idx / nbr_dist,
// built once on the host).  With kappa = (1 + 0.64) / (1 - 0.64) and r_X the distance of the point
// to atom X (r_X >= R_AX - r_A):
//   * the nearest atom lies within R_AC <= 2 r_A;
//   * r_C >= kappa r_min  =>  mu(C, nearest) >= 0.64  =>  P_C = 0 exactly: no candidate C beyond
//     R_AC - r_A >= kappa r_min;
//   * r_B >= kappa r_C  =>  mu(C, B) <= -0.64  =>  the factor s(mu_CB) is exactly 1: the product over
//     B stops at R_AB - r_A >= kappa r_C (and earlier as soon as it reaches 0).
// That is what the host's full pair loop produces for those terms, so the result is unchanged while
// the cost per point drops from O(natoms^2) to O(near atoms^2) -- 1231-atom ubiquitin and the
// 2499-atom water cluster need this.  Products / sums run in neighbour order instead of index
// order (a reordering of exact factors 0 and 1 plus O(1e-16) rounding).  The host additionally
// skips pairs whose two partials are both <= 1e-13; that only perturbs terms below 1e-13 of the sum.
//
// Measured and dropped (round 2, profiles/r02_ssf_ncu.txt): ncu shows 6.5 active threads per instruction --
// for a given candidate C only the ~15-20 % of a warp's points that keep C enter the pair loop.  (1) An FP32
// screen of mu (exactly-0 / exactly-1 factors decided in single precision, FP64 only for the active pairs,
// bit-identical weights) changed nothing (ubiquitin 1096 -> 1114 ms): 25-33 % of the visited pairs ARE active
// and the loop is issue-bound, not FP64-bound.  (2) Per-lane candidate lists (every lane walks its OWN surviving
// candidates, so all lanes sit in the pair loop together) were 5-7x SLOWER (ubiquitin 7.4 s, (H2O)833 32 s): a
// candidate is competitive for all points of a tile or for none, so the shared-candidate loop is long for ~8
// candidates per warp, the per-lane one for every list position; and 1 / R_CB becomes a 32-row gather.
//
// Round 2, later: (3) a squared-distance pre-test of the pair loop (no square root / reciprocal for the pairs whose factor
// is exactly 1; identical weights) is in: taxol 52.1 -> 49.2 ms, ubiquitin 1097 -> 974 ms, (H2O)833 6378 -> 5556 ms.
// (4) A warp-per-point kernel for the points far from every nucleus (nearest atom as the anchor, the kappa^2 r_min
// neighbourhood staged in shared memory with its distances, 32 lanes over the atoms B of one candidate at a time, lane
// products multiplied by shuffles) was built and swept over the hand-over radius (gpurun_out/r02v_ssf_sweep.log): same
// weights, but slower at every radius -- taxol 73 ms with all points, ubiquitin 1103 ms, (H2O)833 14.5 s (2 warps per CTA:
// 30 KB of lists per warp for 2499 atoms) -- the thread-per-point loop visits fewer pairs per point than its 6.5 active
// lanes suggest, because dead candidates die on their first few B and the warp's union of live (C, B) pairs is small.
#include "kernels.cuh"

namespace gxb {

namespace {

constexpr double magic_ssf = 0.64;

__device__ __forceinline__ double g_frisch(double mu) {
  const double s = mu * 1.5625;  // 1 / 0.64 (exactly representable; the host divides by 0.64)
  const double s2 = s * s, s3 = s * s2, s5 = s3 * s2, s7 = s5 * s2;
  return (35. * (s - s3) + 21. * s5 - 5. * s7) / 16.;
}

// Margin of the squared-distance pre-test of the pair loop: a pair is skipped only when (r_B - r_C) / R_CB exceeds 0.64
// by more than 1e-10 relative -- eight orders above the rounding of either evaluation -- so the host's own test
// (mu <= -0.64 resp. mu >= 0.64 on its rounded mu) takes the same branch and the factor is exactly 1 on both sides.
constexpr double skip_fac = magic_ssf * (1. + 1e-10);

__global__ void __launch_bounds__(TP) ssf_kernel(PlanView pv, const DevTile* __restrict__ tiles,
                                                  const double* __restrict__ atoms,
                                                  const double* __restrict__ rab,
                                                  const double* __restrict__ rab_inv,
                                                  const double* __restrict__ dist_nearest,
                                                  const int* __restrict__ nbr_idx,
                                                  const double* __restrict__ nbr_dist,
                                                  int natoms) {
  const DevTile tile = tiles[blockIdx.x];
  const int i = threadIdx.x;

  // anchor atom of the tile: nearest atom to the tile's first point
  __shared__ double s_best[TP / 32];
  __shared__ int s_arg[TP / 32];
  int anchor;
  {
    const double ax = pv.px[tile.pt_off], ay = pv.py[tile.pt_off], az = pv.pz[tile.pt_off];
    double best = 1e300;
    int arg = 0;
    for (int A = i; A < natoms; A += TP) {
      const double dx = ax - atoms[3 * A], dy = ay - atoms[3 * A + 1], dz = az - atoms[3 * A + 2];
      const double d2 = dx * dx + dy * dy + dz * dz;
      if (d2 < best) { best = d2; arg = A; }
    }
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) {
      const double ob = __shfl_xor_sync(0xffffffffu, best, d);
      const int oa = __shfl_xor_sync(0xffffffffu, arg, d);
      if (ob < best || (ob == best && oa < arg)) { best = ob; arg = oa; }
    }
    if ((i & 31) == 0) { s_best[i >> 5] = best; s_arg[i >> 5] = arg; }
    __syncthreads();
    best = s_best[0]; arg = s_arg[0];
#pragma unroll
    for (int w = 1; w < TP / 32; ++w)
      if (s_best[w] < best || (s_best[w] == best && s_arg[w] < arg)) { best = s_best[w]; arg = s_arg[w]; }
    anchor = arg;
  }

  if (i >= tile.npts) return;
  const int ip = tile.pt_off + i;
  const int par = pv.tasks[tile.task].iParent;
  const double px = pv.px[ip], py = pv.py[ip], pz = pv.pz[ip];

  auto dist = [&](int A) {
    const double dx = px - atoms[3 * A], dy = py - atoms[3 * A + 1], dz = pz - atoms[3 * A + 2];
    return sqrt(dx * dx + dy * dy + dz * dz);
  };

  const double r_par = dist(par);
  if (r_par < 0.5 * (1. - magic_ssf) * dist_nearest[par]) return;

  const int* __restrict__ nb = nbr_idx + (size_t)anchor * natoms;     // nb[0] == anchor
  const double* __restrict__ nd = nbr_dist + (size_t)anchor * natoms;  // ascending R(anchor, .)
  const double r_anc = dist(anchor);
  constexpr double kappa = (1. + magic_ssf) / (1. - magic_ssf) * (1. + 1e-9);  // + rounding margin

  // nearest atom: r_C >= R(anchor, C) - r_anc > r_anc once R(anchor, C) > 2 r_anc
  double rmin = r_anc;
  int imin = anchor;
  for (int k = 1; k < natoms; ++k) {
    if (nd[k] > 2. * r_anc * (1. + 1e-9)) break;
    const int A = nb[k];
    const double r = dist(A);
    if (r < rmin) { rmin = r; imin = A; }
  }

  double sum = 0., p_par = 0.;
  const double c_cut = kappa * rmin;
  for (int kc = 0; kc < natoms; ++kc) {
    if (nd[kc] - r_anc >= c_cut) break;  // every remaining C has P_C == 0
    const int C = nb[kc];
    const double rC = dist(C);
    if (C != imin) {
      const double Rinv = rab_inv[(size_t)C * natoms + imin];
      // host evaluates this pair as (iA,jA) = (max,min); mu is exactly antisymmetric
      const double mu = (C > imin) ? (rC - rmin) * Rinv : -((rmin - rC) * Rinv);
      if (mu >= magic_ssf) continue;  // P_C == 0
    }
    double Pc = 1.;
    const double* __restrict__ rabC = rab_inv + (size_t)C * natoms;
    const double* __restrict__ rabR = rab + (size_t)C * natoms;
    const double b_cut = kappa * rC;
    for (int kb = 0; kb < natoms; ++kb) {
      if (nd[kb] - r_anc >= b_cut) break;  // every remaining factor is exactly 1
      const int Bq = nb[kb];
      if (Bq == C) continue;
      // Most B inside the cut-off sphere still give a factor of exactly 1 (r_B - r_C >= 0.64 R_CB): decide that on
      // squared distances -- no square root, no reciprocal -- and evaluate mu only for the pairs that compete.
      const double bx = px - atoms[3 * Bq], by = py - atoms[3 * Bq + 1], bz = pz - atoms[3 * Bq + 2];
      const double rB2 = bx * bx + by * by + bz * bz;
      const double T = fma(skip_fac, rabR[Bq], rC);
      if (rB2 >= T * T) continue;
      const double rB = sqrt(rB2);
      const double Rinv = rabC[Bq];
      if (Bq < C) {
        const double mu = (rC - rB) * Rinv;
        if (mu <= -magic_ssf) continue;
        if (mu >= magic_ssf) { Pc = 0.; break; }
        Pc *= 0.5 * (1. - g_frisch(mu));
      } else {
        const double mu = (rB - rC) * Rinv;
        if (mu <= -magic_ssf) { Pc = 0.; break; }
        if (mu >= magic_ssf) continue;
        const double gq = 0.5 * (1. - g_frisch(mu));
        Pc *= 1. - gq;
      }
    }
    sum += Pc;
    if (C == par) p_par = Pc;
  }
  pv.w[ip] *= p_par / sum;
}

}  // namespace

void launch_ssf_weights(const PlanView& pv, const DevTile* tiles, int ntiles, const double* atoms,
                        const double* rab, const double* rab_inv, const double* dist_nearest, const int* nbr_idx,
                        const double* nbr_dist, int natoms, cudaStream_t s) {
  if (ntiles <= 0) return;
  ssf_kernel<<<ntiles, TP, 0, s>>>(pv, tiles, atoms, rab, rab_inv, dist_nearest, nbr_idx, nbr_dist, natoms);
}

}  // namespace gxb
