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
// FP32 screen.  Of the pairs (C, B) the two loops visit, only a handful per point are "active"
// (|mu_CB| < 0.64: a factor strictly between 0 and 1); all others contribute EXACTLY 1 (mu <= -0.64) or make
// P_C EXACTLY 0 (mu >= 0.64).  Which of the three a pair is follows from a single-precision mu (distances
// from FP32 atom coordinates, an FP32 copy of 1 / R_AB: ~8 FP32 instructions, no FP64 square root) whenever
// that mu is farther than 2e-3 from +-0.64 -- three orders of magnitude more than its rounding error; only
// the rest takes the FP64 path, with the operand order above.  The set of active factors and the order they
// are multiplied in are unchanged, so the weights are bit-identical to the all-FP64 kernel; the FP64 work per
// point drops from O(near atoms^2) to O(active pairs).
#include "kernels.cuh"

namespace gxb {

namespace {

constexpr double magic_ssf = 0.64;
constexpr float screen_lo = 0.64f - 2e-3f, screen_hi = 0.64f + 2e-3f;
constexpr int SSF_MAXC = 40;  // surviving candidates kept per point before they are flushed (20 KB of shared memory)

__device__ __forceinline__ double g_frisch(double mu) {
  const double s = mu * 1.5625;  // 1 / 0.64 (exactly representable; the host divides by 0.64)
  const double s2 = s * s, s3 = s * s2, s5 = s3 * s2, s7 = s5 * s2;
  return (35. * (s - s3) + 21. * s5 - 5. * s7) / 16.;
}

__global__ void __launch_bounds__(TP) ssf_kernel(PlanView pv, const DevTile* __restrict__ tiles,
                                                  const double* __restrict__ atoms,
                                                  const float4* __restrict__ atoms_f,
                                                  const double* __restrict__ rab_inv,
                                                  const float* __restrict__ rab_inv_f,
                                                  const double* __restrict__ dist_nearest,
                                                  const int* __restrict__ nbr_idx,
                                                  const double* __restrict__ nbr_dist,
                                                  int natoms) {
  const DevTile tile = tiles[blockIdx.x];
  const int i = threadIdx.x;

  // anchor atom of the tile: nearest atom to the tile's first point
  __shared__ double s_best[TP / 32];
  __shared__ int s_arg[TP / 32];
  __shared__ int s_cand[SSF_MAXC][TP];  // per lane: surviving candidates (column = thread: conflict-free)
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
  const float pxf = (float)px, pyf = (float)py, pzf = (float)pz;

  auto dist = [&](int A) {
    const double dx = px - atoms[3 * A], dy = py - atoms[3 * A + 1], dz = pz - atoms[3 * A + 2];
    return sqrt(dx * dx + dy * dy + dz * dz);
  };
  auto dist_f = [&](int A) {
    const float4 a = __ldg(atoms_f + A);
    const float dx = pxf - a.x, dy = pyf - a.y, dz = pzf - a.z;
    return sqrtf(dx * dx + dy * dy + dz * dz);
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
  const float rminf = (float)rmin;

  // P_C = prod_{B != C} s(mu_CB) for one candidate: neighbour order, FP32 screen, FP64 for the active pairs
  auto partition_function = [&](int C) {
    const double rC = dist(C);
    const float4 cf = __ldg(atoms_f + C);
    const float rCf = dist_f(C);
    double Pc = 1.;
    const double* __restrict__ rabC = rab_inv + (size_t)C * natoms;
    const double b_lim = kappa * rC + r_anc;  // factors beyond R(anchor, B) >= b_lim are exactly 1
    for (int kb = 0; kb < natoms; ++kb) {
      if (nd[kb] >= b_lim) break;
      const int Bq = nb[kb];
      if (Bq == C) continue;
      // FP32 screen: mu_CB = (r_C - r_B) / R_CB, everything from FP32 coordinates (no table gather: the lanes of
      // a warp work on different C)
      const float4 bf = __ldg(atoms_f + Bq);
      const float bx = pxf - bf.x, by = pyf - bf.y, bz = pzf - bf.z;
      const float cx = cf.x - bf.x, cy = cf.y - bf.y, cz = cf.z - bf.z;
      const float muf = (rCf - sqrtf(bx * bx + by * by + bz * bz)) * rsqrtf(cx * cx + cy * cy + cz * cz);
      if (muf <= -screen_hi) continue;              // factor exactly 1
      if (muf >= screen_hi) { Pc = 0.; break; }     // factor exactly 0
      // FP64 path, host operand order
      const double rB = dist(Bq);
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
    return Pc;
  };

  // Pass 1 (cheap, per lane): the candidates C that the nearest atom does not already zero.  Pass 2 walks each
  // lane's OWN list, so all lanes of a warp sit in the pair loop together -- with one shared candidate index
  // per warp iteration only ~6 of 32 lanes had work (ncu: 6.5 active threads per instruction), because each
  // point keeps a different ~15 % of the candidates.
  double sum = 0., p_par = 0.;
  int cnt = 0;
  const double c_lim = kappa * rmin + r_anc;  // candidates: R(anchor, C) < c_lim
  for (int kc = 0; kc < natoms; ++kc) {
    if (nd[kc] >= c_lim) break;  // every remaining C has P_C == 0
    const int C = nb[kc];
    if (C != imin) {
      // mu(C, nearest) >= 0.64  =>  P_C == 0: decided in FP32 unless too close to the threshold
      const float muf = (dist_f(C) - rminf) * __ldg(rab_inv_f + (size_t)C * natoms + imin);
      if (muf >= screen_hi) continue;
      if (muf > screen_lo) {
        const double rC = dist(C);
        const double Rinv = rab_inv[(size_t)C * natoms + imin];
        // host evaluates this pair as (iA,jA) = (max,min); mu is exactly antisymmetric
        const double mu = (C > imin) ? (rC - rmin) * Rinv : -((rmin - rC) * Rinv);
        if (mu >= magic_ssf) continue;  // P_C == 0
      }
    }
    if (cnt < SSF_MAXC) {
      s_cand[cnt++][i] = C;
    } else {  // list full (rare): evaluate in place, same order of the sum
      for (int j = 0; j < cnt; ++j) {
        const int Cj = s_cand[j][i];
        const double Pc = partition_function(Cj);
        sum += Pc;
        if (Cj == par) p_par = Pc;
      }
      cnt = 0;
      s_cand[cnt++][i] = C;
    }
  }
  for (int j = 0; j < cnt; ++j) {
    const int C = s_cand[j][i];
    const double Pc = partition_function(C);
    sum += Pc;
    if (C == par) p_par = Pc;
  }
  pv.w[ip] *= p_par / sum;
}

}  // namespace

void launch_ssf_weights(const PlanView& pv, const DevTile* tiles, int ntiles, const double* atoms,
                        const float* atoms_f4, const double* rab_inv, const float* rab_inv_f,
                        const double* dist_nearest, const int* nbr_idx, const double* nbr_dist, int natoms,
                        cudaStream_t s) {
  if (ntiles <= 0) return;
  ssf_kernel<<<ntiles, TP, 0, s>>>(pv, tiles, atoms, reinterpret_cast<const float4*>(atoms_f4), rab_inv, rab_inv_f,
                                   dist_nearest, nbr_idx, nbr_dist, natoms);
}

}  // namespace gxb
