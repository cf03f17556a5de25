// SSF (Stratmann-Scuseria-Frisch) molecular partition weights, thread = grid point.
//
// Arithmetic follows the HOST reference reference_ssf_weights_host
// (src/xc_integrator/local_work_driver/host/reference/weights.cxx:117-237), not the
// reference's CUDA kernel (cuda_ssf_1d.cu:24-128), whose algorithm differs:
//   * cutoff: r_parent < 0.5*(1-0.64)*dist_nearest  -> weight unchanged
//   * P_C = prod_{B != C} s(mu_CB) accumulated in ascending B, with the host's operand order
//     (B < C: g(mu_CB);  B > C: 1 - g(mu_BC)),  mu = (r_i - r_j) / RAB (a true division)
//   * w *= P_parent / sum_C P_C, sum in ascending C
// Instead of materialising natoms distances per point (the reference device path stores a
// natoms x npts scratch, xc_device_data.hpp:449-451) distances are recomputed, and a
// candidate C is dropped at once when the atom nearest to the point already forces
// P_C = 0 exactly (mu >= 0.64 against the nearest atom), which is what the host's pair
// loop produces for that C.  The host additionally skips pairs whose two partials are both
// <= 1e-13; that only perturbs terms below 1e-13 of the sum.
#include "kernels.cuh"

namespace gxb {

namespace {

constexpr double magic_ssf = 0.64;

__device__ __forceinline__ double g_frisch(double mu) {
  const double s = mu / magic_ssf;
  const double s2 = s * s, s3 = s * s2, s5 = s3 * s2, s7 = s5 * s2;
  return (35. * (s - s3) + 21. * s5 - 5. * s7) / 16.;
}

__global__ void __launch_bounds__(TP) ssf_kernel(PlanView pv, const DevTile* __restrict__ tiles,
                                                  const double* __restrict__ atoms,
                                                  const double* __restrict__ rab,
                                                  const double* __restrict__ dist_nearest,
                                                  int natoms) {
  const DevTile tile = tiles[blockIdx.x];
  const int i = threadIdx.x;
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

  // nearest atom
  double rmin = r_par;
  int imin = par;
  for (int A = 0; A < natoms; ++A) {
    const double r = dist(A);
    if (r < rmin) { rmin = r; imin = A; }
  }

  double sum = 0., p_par = 0.;
  for (int C = 0; C < natoms; ++C) {
    const double rC = dist(C);
    if (C != imin) {
      const double R = rab[(size_t)C * natoms + imin];
      // host evaluates this pair as (iA,jA) = (max,min); mu is exactly antisymmetric
      const double mu = (C > imin) ? (rC - rmin) / R : -((rmin - rC) / R);
      if (mu >= magic_ssf) continue;  // P_C == 0
    }
    double Pc = 1.;
    const double* __restrict__ rabC = rab + (size_t)C * natoms;
    for (int Bq = 0; Bq < natoms; ++Bq) {
      if (Bq == C) continue;
      const double rB = dist(Bq);
      const double R = rabC[Bq];
      if (Bq < C) {
        const double mu = (rC - rB) / R;
        if (mu <= -magic_ssf) continue;
        if (mu >= magic_ssf) { Pc = 0.; break; }
        Pc *= 0.5 * (1. - g_frisch(mu));
      } else {
        const double mu = (rB - rC) / R;
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
                        const double* rab, const double* dist_nearest, int natoms,
                        cudaStream_t s) {
  if (ntiles <= 0) return;
  ssf_kernel<<<ntiles, TP, 0, s>>>(pv, tiles, atoms, rab, dist_nearest, natoms);
}

}  // namespace gxb
