// EXC gradient kernels: Hessian collocation, the gradient assembly and the SSF weight
// derivatives.  Host semantics followed:
//   * gau2grid_collocation_hessian (local_work_driver/host/reference/gau2grid_collocation.cxx:153-216)
//   * the shell / point loops of reference_replicated_xc_host_integrator_exc_grad.hpp:404-590 (RKS)
//   * reference_ssf_weights_1std_contraction_host (host/reference/weights.cxx:806-984)
// replacing the reference device kernels increment_exc_grad_{lda,gga} (kernels/increment_exc_grad.cu) and
// eval_weight_1st_deriv_contracted_ssf_kernel_1d (kernels/cuda_ssf_1d.cu:146-350).
//
// Per batch of tiles the integrator runs: collocation (gradient for LDA, Hessian for GGA) -> the fused DMMA kernel in
// XOUT mode for X = 2 B P_sub -> exc_grad_kernel.  LDA: one pass (PHASE 2).  GGA: the host needs X_a = 2 (dB/da) P_sub only
// inside d11 = sum_a (d rho / da) X_a, which is LINEAR in the per-point vector grad rho: PHASE 0 evaluates the densities
// and the functional and writes U = 2 w vgamma sum_a (d rho / da) dB/da (one matrix), the DMMA kernel forms Y = 2 U P_sub,
// PHASE 1 assembles with Y in place of 2 w vgamma d11 -- TWO contractions instead of the reference's four
// (eval_xmat over xmat_len = 4 blocks, reference_replicated_xc_host_integrator_exc_grad.hpp:370-373), same sums up to the
// order of the additions.  UKS likewise: U_s, U_z carry the four vgamma combinations, four contractions instead of eight.
// Tile matrices ([pad16(nbe)][TP], swizzled as everywhere, device_plan.hpp): LDA  B dx dy dz | X (UKS: XN XZ);
// GGA  B dx dy dz xx xy xz yy yz zz | X | U | Y | F  (UKS: XN XZ | U_s U_z | Y_s Y_z | F), F = per-point factors.
#include <algorithm>
#include <initializer_list>

#include "kernels.cuh"
#include "ptx.cuh"
#include "xc_functionals.cuh"
#include "xc_functionals_pol_gga.cuh"

namespace gxb {

namespace {

// ------------------------------------------------------------------------------------------------
// Hessian collocation: thread = point, ten output matrices.  phi = f(x,y,z) S(r^2), f a monomial:
//   d_i phi  = f_i S0 + f x_i S1
//   d_ij phi = f_ij S0 + (f_i x_j + f_j x_i + f delta_ij) S1 + f x_i x_j S2
// with S0 = sum c e, S1 = sum -2 a c e, S2 = sum 4 a^2 c e.  Pure shells: sparse real solid harmonics (CCA order
// m = -l..l, gau2grid normalisation) as combinations of monomials.
// ------------------------------------------------------------------------------------------------
struct SphTerm { signed char a, b, c, n; double f; };  // n: terms in this row (stored in the row's first entry)
__constant__ SphTerm c_sph[5][9][6];

template <int NM>
__device__ __forceinline__ void store_n(double* __restrict__ B, size_t ms, int row, int i, bool ok, const double (&o)[10]) {
  const size_t off = (size_t)row * TP + swz(row, i);
#pragma unroll
  for (int q = 0; q < NM; ++q) B[off + q * ms] = ok ? o[q] : 0.;
}

__global__ void __launch_bounds__(TP) collocation_hessian_kernel(PlanView pv, const DevTile* __restrict__ tiles,
                                                                 double* __restrict__ ws) {
  const DevTile tile = tiles[blockIdx.x];
  const DevTask task = pv.tasks[tile.task];
  const int i = threadIdx.x;
  if (i >= tile_width(tile.npts)) return;
  const bool ok = i < tile.npts;
  const int ip = tile.pt_off + (ok ? i : 0);
  const double px = pv.px[ip], py = pv.py[ip], pz = pv.pz[ip];
  double* __restrict__ B = ws + tile.ws_off;
  const int nbp = pad16(task.nbe);
  const size_t ms = (size_t)nbp * TP;
  {
    const double zero[10] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0};
    for (int r = task.nbe; r < nbp; ++r) store_n<10>(B, ms, r, i, false, zero);
  }
  for (int s = 0; s < task.nshells; ++s) {
    const DevShell sh = pv.shells[pv.task_shells[task.shell_off + s]];
    const int bf = pv.task_shell_bf[task.shell_off + s];
    const double r[3] = {px - sh.x, py - sh.y, pz - sh.z};
    const double r2 = r[0] * r[0] + r[1] * r[1] + r[2] * r[2];
    double S0 = 0., S1 = 0., S2 = 0.;
    const double* __restrict__ al = pv.prim_alpha + sh.prim_off;
    const double* __restrict__ co = pv.prim_coeff + sh.prim_off;
    for (int k = 0; k < sh.nprim; ++k) {
      const double a = __ldg(al + k);
      const double e = __ldg(co + k) * exp(-a * r2);
      S0 += e;
      S1 += -2. * a * e;
      S2 += 4. * a * a * e;
    }
    double pw[3][5];
#pragma unroll
    for (int d = 0; d < 3; ++d) {
      pw[d][0] = 1.;
#pragma unroll
      for (int k = 1; k <= 4; ++k) pw[d][k] = pw[d][k - 1] * r[d];
    }
    // monomial x^n0 y^n1 z^n2 with the exponents lowered by (d0, d1, d2), times the falling factors
    auto fpow = [&](const int (&n)[3], int d0, int d1, int d2) {
      const int e0 = n[0] - d0, e1 = n[1] - d1, e2 = n[2] - d2;
      if (e0 < 0 || e1 < 0 || e2 < 0) return 0.;
      double pre = 1.;
      for (int k = 0; k < d0; ++k) pre *= double(n[0] - k);
      for (int k = 0; k < d1; ++k) pre *= double(n[1] - k);
      for (int k = 0; k < d2; ++k) pre *= double(n[2] - k);
      return pre * pw[0][e0] * pw[1][e1] * pw[2][e2];
    };
    auto mono = [&](int a, int b, int c, double (&o)[10]) {
      const int n[3] = {a, b, c};
      const double f = fpow(n, 0, 0, 0);
      const double f1[3] = {fpow(n, 1, 0, 0), fpow(n, 0, 1, 0), fpow(n, 0, 0, 1)};
      o[0] = f * S0;
#pragma unroll
      for (int d = 0; d < 3; ++d) o[1 + d] = f1[d] * S0 + f * r[d] * S1;
      o[4] = fpow(n, 2, 0, 0) * S0 + (2. * f1[0] * r[0] + f) * S1 + f * r[0] * r[0] * S2;
      o[5] = fpow(n, 1, 1, 0) * S0 + (f1[0] * r[1] + f1[1] * r[0]) * S1 + f * r[0] * r[1] * S2;
      o[6] = fpow(n, 1, 0, 1) * S0 + (f1[0] * r[2] + f1[2] * r[0]) * S1 + f * r[0] * r[2] * S2;
      o[7] = fpow(n, 0, 2, 0) * S0 + (2. * f1[1] * r[1] + f) * S1 + f * r[1] * r[1] * S2;
      o[8] = fpow(n, 0, 1, 1) * S0 + (f1[1] * r[2] + f1[2] * r[1]) * S1 + f * r[1] * r[2] * S2;
      o[9] = fpow(n, 0, 0, 2) * S0 + (2. * f1[2] * r[2] + f) * S1 + f * r[2] * r[2] * S2;
    };
    const int l = sh.l;
    if (!sh.pure) {
      int c = 0;
      for (int a = l; a >= 0; --a)
        for (int b = l - a; b >= 0; --b, ++c) {
          double o[10];
          mono(a, b, l - a - b, o);
          store_n<10>(B, ms, bf + c, i, ok, o);
        }
    } else {
      for (int m = 0; m < 2 * l + 1; ++m) {
        double acc[10] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0};
        const int nt = c_sph[l][m][0].n;
        for (int t = 0; t < nt; ++t) {
          const SphTerm tt = c_sph[l][m][t];
          double o[10];
          mono(tt.a, tt.b, tt.c, o);
#pragma unroll
          for (int q = 0; q < 10; ++q) acc[q] += tt.f * o[q];
        }
        store_n<10>(B, ms, bf + m, i, ok, acc);
      }
    }
  }
}

// ------------------------------------------------------------------------------------------------
// Gradient assembly.  Persistent CTAs of 256 threads pull tiles from *counter; thread (p, h) = point p of the tile,
// basis rows mu = h (mod 2).  Phase A: rho = sum B X, grad rho = 2 sum dB X (eval_uvvar_{lda,gga}_rks), functional,
// weights.  Phase B: per shell of the task, rows of the shell
//   g += w vrho X dB  [+ 2 w vgamma (X (H grad rho) + dB (grad rho . X_grad))]
// summed over the rows of consecutive shells on the same atom, reduced over the points by warp shuffles and added as
// -2 g to the atom (with weight derivatives: shells on the parent atom are skipped and the parent receives +2 g;
// :404-410, 575-590).  Atom sums accumulate in shared memory (3 natoms doubles per CTA) and reach HBM once per CTA.
// ------------------------------------------------------------------------------------------------
constexpr int EG_THREADS = 256;

// PHASE 2: everything (LDA).  PHASE 0 (GGA): densities, functional, per-point factors -> F, U rows.  PHASE 1 (GGA): assembly.
// Two CTAs per SM where the registers allow (the kernel streams up to 13 matrices and wants the occupancy).
template <bool GGA, bool UKS, int PHASE>
__global__ void __launch_bounds__(EG_THREADS, (UKS && GGA && PHASE == 0) ? 1 : 2)
exc_grad_kernel(PlanView pv, const DevTile* __restrict__ tiles, int ntiles, int* __restrict__ counter,
                double* __restrict__ ws, FunctionalDesc func, const int* __restrict__ shell_atom, int natoms,
                int include_wd, double* __restrict__ wf_out, double* __restrict__ grad, int smem_acc) {
  extern __shared__ double eg_dyn[];  // smem_acc: 3 natoms accumulators
  __shared__ double part[2][UKS ? 8 : 4][TP];
  __shared__ int s_tile;
  const int tid = threadIdx.x, p = tid & (TP - 1), h = tid >> 7, lane = tid & 31;
  constexpr int NB = GGA ? 10 : 4;   // basis matrices ahead of X
  constexpr int ND = UKS ? 2 : 1;    // densities: X, U, Y blocks of ND matrices each
  constexpr int SU = NB + ND, SY = NB + 2 * ND, SF = NB + 3 * ND;  // GGA slots of U, Y and the factor rows
  constexpr bool DO_A = PHASE != 1, DO_B = PHASE != 0;
  if (DO_B && smem_acc)
    for (int q = tid; q < 3 * natoms; q += EG_THREADS) eg_dyn[q] = 0.;
  __syncthreads();
  auto add_atom = [&](int atom, int c, double v) {
    if (smem_acc) atomicAdd(&eg_dyn[3 * atom + c], v);
    else atomicAdd(&grad[3 * atom + c], v);
  };

  for (;;) {
    if (tid == 0) s_tile = atomicAdd(counter, 1);
    __syncthreads();
    const int tile_idx = s_tile;
    if (tile_idx >= ntiles) break;
    const DevTile tile = tiles[tile_idx];
    const DevTask task = pv.tasks[tile.task];
    const int nbe = tile.nbe;
    const size_t ms = (size_t)pad16(nbe) * TP;
    double* __restrict__ M = ws + tile.ws_off;
    const bool ok = p < tile.npts;
    // element (mu, p) of matrix q: M[q * ms + mu * TP + (p ^ ((mu & 3) << 2))]
    auto at = [&](int q, int mu) { return M[(size_t)q * ms + (size_t)mu * TP + swz(mu, p)]; };
    double* __restrict__ F = M + (size_t)SF * ms;  // GGA: factor row q of point p at F[q * TP + p]

    // RKS: wv = w vrho, wg = w vgamma, (dx, dy, dz) = grad rho.  UKS (:424-436, 482-513): wv = 1/2 w (v+ + v-),
    // wvz = 1/2 w (v+ - v-), wg = c1 = 1/2 w (v++ + v+- + v--), c2 = 1/2 w (v++ - v--), c3 = 1/2 w (v++ - v+- + v--),
    // (dx, dy, dz) = grad n, (mx, my, mz) = grad M_z
    double wv = 0., wg = 0., wvz = 0., c2 = 0., c3 = 0.;
    double dx = 0., dy = 0., dz = 0., mx = 0., my = 0., mz = 0.;
    if (DO_A) {
      // ---- phase A
      double r0 = 0., r1 = 0., r2 = 0., r3 = 0.;
      double q0 = 0., q1 = 0., q2 = 0., q3 = 0.;  // UKS: the same sums with X of Pz
      if (ok) {
#pragma unroll 4
        for (int mu = h; mu < nbe; mu += 2) {
          const double x = at(NB, mu);
          const double b0 = at(0, mu);
          r0 = fma(b0, x, r0);
          double b1 = 0., b2 = 0., b3 = 0.;
          if (GGA) {
            b1 = at(1, mu); b2 = at(2, mu); b3 = at(3, mu);
            r1 = fma(b1, x, r1);
            r2 = fma(b2, x, r2);
            r3 = fma(b3, x, r3);
          }
          if (UKS) {
            const double xz = at(NB + 1, mu);
            q0 = fma(b0, xz, q0);
            if (GGA) {
              q1 = fma(b1, xz, q1);
              q2 = fma(b2, xz, q2);
              q3 = fma(b3, xz, q3);
            }
          }
        }
      }
      part[h][0][p] = r0;
      if (GGA) { part[h][1][p] = r1; part[h][2][p] = r2; part[h][3][p] = r3; }
      if (UKS) {
        part[h][4][p] = q0;
        if (GGA) { part[h][5][p] = q1; part[h][6][p] = q2; part[h][7][p] = q3; }
      }
      __syncthreads();
      const double rho = part[0][0][p] + part[1][0][p];
      if (GGA) {
        dx = 2. * (part[0][1][p] + part[1][1][p]);
        dy = 2. * (part[0][2][p] + part[1][2][p]);
        dz = 2. * (part[0][3][p] + part[1][3][p]);
      }
      double rho_z = 0.;
      if (UKS) {
        rho_z = part[0][4][p] + part[1][4][p];
        if (GGA) {
          mx = 2. * (part[0][5][p] + part[1][5][p]);
          my = 2. * (part[0][6][p] + part[1][6][p]);
          mz = 2. * (part[0][7][p] + part[1][7][p]);
        }
      }
      if (ok) {
        const double w = pv.w[tile.pt_off + p];
        double eps;
        if (!UKS) {
          const XcOut xc = eval_functional(func, rho, GGA ? dx * dx + dy * dy + dz * dz : 0.);
          wv = w * xc.vrho;
          wg = w * xc.vsigma;
          eps = xc.eps;
        } else if (GGA) {
          const double dn_sq = dx * dx + dy * dy + dz * dz, dm_sq = mx * mx + my * my + mz * mz,
                       dn_dm = dx * mx + dy * my + dz * mz;
          const double gpp = 0.25 * (dn_sq + dm_sq) + 0.5 * dn_dm, gpm = 0.25 * (dn_sq - dm_sq),
                       gmm = 0.25 * (dn_sq + dm_sq) - 0.5 * dn_dm;
          const XcOutPolGga xc = eval_functional_pol(func, 0.5 * (rho + rho_z), 0.5 * (rho - rho_z), gpp, gpm, gmm);
          const double vp = w * xc.va, vm = w * xc.vb;
          wv = 0.5 * (vp + vm);
          wvz = 0.5 * (vp - vm);
          const double vpp = w * xc.vaa, vpm = w * xc.vab, vmm = w * xc.vbb;
          wg = 0.5 * (vpp + vpm + vmm);
          c2 = 0.5 * (vpp - vmm);
          c3 = 0.5 * (vpp - vpm + vmm);
          eps = xc.eps;
        } else {
          const XcOutPol xc = eval_functional_pol_lda(func, 0.5 * (rho + rho_z), 0.5 * (rho - rho_z));
          const double vp = w * xc.va, vm = w * xc.vb;
          wv = 0.5 * (vp + vm);
          wvz = 0.5 * (vp - vm);
          eps = xc.eps;
        }
        if (include_wd && h == 0) wf_out[tile.pt_off + p] = eps * (rho * w);  // eps *= den * w (:396-399)
      }
      if (PHASE == 0) {
        // factor rows for the assembly pass, then U: the d11 terms of the host loop are linear in grad rho, so
        //   RKS  2 wg d11           = (2 P_sub U)_mu,          U   = 2 wg sum_a (d rho / da) dB/da
        //   UKS  wg d11nn + c2 (d11zn + d11nz) + c3 d11zz = (P_s U_s + P_z U_z)_mu,
        //        U_s = sum_a (wg dn_a + c2 dm_a) dB/da,  U_z = sum_a (c2 dn_a + c3 dm_a) dB/da
        if (h == 0 && p < tile_width(tile.npts)) {
          F[0 * TP + p] = wv; F[1 * TP + p] = wg; F[2 * TP + p] = dx; F[3 * TP + p] = dy; F[4 * TP + p] = dz;
          if (UKS) {
            F[5 * TP + p] = wvz; F[6 * TP + p] = c2; F[7 * TP + p] = c3;
            F[8 * TP + p] = mx; F[9 * TP + p] = my; F[10 * TP + p] = mz;
          }
        }
        if (p < tile_width(tile.npts)) {
          const double ux = UKS ? wg * dx + c2 * mx : 2. * wg * dx, uy = UKS ? wg * dy + c2 * my : 2. * wg * dy,
                       uz = UKS ? wg * dz + c2 * mz : 2. * wg * dz;
          const double vx = c2 * dx + c3 * mx, vy = c2 * dy + c3 * my, vz = c2 * dz + c3 * mz;
          const int nbp = pad16(nbe);
          for (int mu = h; mu < nbp; mu += 2) {
            const size_t o = (size_t)mu * TP + swz(mu, p);
            double us = 0., uzv = 0.;
            if (ok && mu < nbe) {
              const double b1 = M[ms + o], b2 = M[2 * ms + o], b3 = M[3 * ms + o];
              us = ux * b1 + uy * b2 + uz * b3;
              if (UKS) uzv = vx * b1 + vy * b2 + vz * b3;
            }
            M[(size_t)SU * ms + o] = us;
            if (UKS) M[(size_t)(SU + 1) * ms + o] = uzv;
          }
        }
      }
    } else if (ok) {
      wv = F[0 * TP + p]; wg = F[1 * TP + p]; dx = F[2 * TP + p]; dy = F[3 * TP + p]; dz = F[4 * TP + p];
      if (UKS) {
        wvz = F[5 * TP + p]; c2 = F[6 * TP + p]; c3 = F[7 * TP + p];
        mx = F[8 * TP + p]; my = F[9 * TP + p]; mz = F[10 * TP + p];
      }
    }

    if (DO_B) {
      // ---- phase B
      double acc[3] = {0., 0., 0.}, par[3] = {0., 0., 0.};
      int cur_atom = -1;
      auto flush = [&]() {
        if (cur_atom < 0) return;
        double v[3] = {acc[0], acc[1], acc[2]};
#pragma unroll
        for (int c = 0; c < 3; ++c) {
#pragma unroll
          for (int d = 16; d > 0; d >>= 1) v[c] += __shfl_xor_sync(0xffffffffu, v[c], d);
        }
        if (lane == 0) {
#pragma unroll
          for (int c = 0; c < 3; ++c) {
            add_atom(cur_atom, c, -2. * v[c]);
            par[c] += 2. * v[c];
          }
        }
        acc[0] = acc[1] = acc[2] = 0.;
      };
      for (int s = 0; s < task.nshells; ++s) {
        const int atom = __ldg(shell_atom + __ldg(pv.task_shells + task.shell_off + s));
        if (include_wd && atom == task.iParent) continue;
        if (atom != cur_atom) {
          flush();
          cur_atom = atom;
        }
        if (!ok) continue;
        const int bf0 = __ldg(pv.task_shell_bf + task.shell_off + s);
        const int bf1 = s + 1 < task.nshells ? __ldg(pv.task_shell_bf + task.shell_off + s + 1) : nbe;
        for (int mu = bf0 + ((bf0 ^ h) & 1); mu < bf1; mu += 2) {
          const double xn = at(NB, mu);
          const double xz = UKS ? at(NB + 1, mu) : 0.;
          const double dbx = at(1, mu), dby = at(2, mu), dbz = at(3, mu);
          double a = UKS ? wv * xn + wvz * xz : wv * xn;
          if (GGA) a += UKS ? at(SY, mu) + at(SY + 1, mu) : at(SY, mu);  // the d11 terms, contracted by the DMMA pass
          acc[0] = fma(a, dbx, acc[0]);
          acc[1] = fma(a, dby, acc[1]);
          acc[2] = fma(a, dbz, acc[2]);
          if (GGA) {
            const double xx = at(4, mu), xy = at(5, mu), xzz = at(6, mu), yy = at(7, mu), yz = at(8, mu), zz = at(9, mu);
            const double d2x = xx * dx + xy * dy + xzz * dz;  // H grad rho (UKS: H grad n)
            const double d2y = xy * dx + yy * dy + yz * dz;
            const double d2z = xzz * dx + yz * dy + zz * dz;
            if (!UKS) {
              const double g2x = 2. * wg * xn;
              acc[0] = fma(g2x, d2x, acc[0]);
              acc[1] = fma(g2x, d2y, acc[1]);
              acc[2] = fma(g2x, d2z, acc[2]);
            } else {
              const double e2x = xx * mx + xy * my + xzz * mz;  // H grad M_z
              const double e2y = xy * mx + yy * my + yz * mz;
              const double e2z = xzz * mx + yz * my + zz * mz;
              // c1 d2n xN + c2 d2z xN + c2 d2n xZ + c3 d2z xZ
              const double sx = wg * xn + c2 * xz, tx = c2 * xn + c3 * xz;  // multiply d2n resp. d2z
              acc[0] += sx * d2x + tx * e2x;
              acc[1] += sx * d2y + tx * e2y;
              acc[2] += sx * d2z + tx * e2z;
            }
          }
        }
      }
      flush();
      if (include_wd && lane == 0) {
#pragma unroll
        for (int c = 0; c < 3; ++c)
          if (par[c] != 0.) add_atom(task.iParent, c, par[c]);
      }
    }
    __syncthreads();  // part[] and s_tile are reused by the next tile
  }
  if (DO_B && smem_acc) {
    __syncthreads();
    for (int q = tid; q < 3 * natoms; q += EG_THREADS) {
      const double v = eg_dyn[q];
      if (v != 0.) atomicAdd(&grad[q], v);
    }
  }
}

// ------------------------------------------------------------------------------------------------
// SSF weight derivatives contracted with wf = w eps rho.  One warp per grid point, lanes over atoms.  With
// kappa = (1 + 0.64)/(1 - 0.64) and r_min the distance to the nearest atom (the bounds of ssf_weights.cu):
//   * P_A = 0 exactly unless r_A < kappa r_min                                   (list0)
//   * atom B changes P_A (s(mu_AB) != 1) only if r_B < kappa r_A < kappa^2 r_min  (list1)
//   * the derivative terms need |mu_BC| < 0.64 - 1e-4 with P_B > 1e-13, or |mu_parent,B| < 0.64 - 1e-4 with
//     P_parent > 0 (wf = 0 otherwise): all inside list1.
// So every loop of the host function runs over list1 (compacted into shared memory with coordinates and distances)
// instead of all atoms, with identical terms.  The host's pair loop skips pairs whose two partial products are both
// <= 1e-13; here P_A is the full product (differences below 1e-13 of the sum).  1 / R_AB comes from the plan's table
// (the SSF weights kernel's), so the pair loops hold no square root or division by R.  Atom sums accumulate in warp-private shared memory (no atomics) and reach HBM once per warp.
// ------------------------------------------------------------------------------------------------
constexpr double magic_ssf = 0.64;

__device__ __forceinline__ double g_frisch_d(double x) {
  const double s = x / magic_ssf;
  const double s2 = s * s, s3 = s * s2, s5 = s3 * s2, s7 = s5 * s2;
  return (35. * (s - s3) + 21. * s5 - 5. * s7) / 16.;
}
__device__ __forceinline__ double t_frisch_d(double x) {
  const double s = x / magic_ssf;
  const double s2 = s * s, s3 = s * s2;
  const double num = 35. * (s3 + 3. * s2 + 3. * s + 1.);
  const double den = (x - magic_ssf) * (5. * s3 + 20. * s2 + 29. * s + 16.);
  return num / den;
}
__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int d = 16; d > 0; d >>= 1) v += __shfl_xor_sync(0xffffffffu, v, d);
  return v;
}

__global__ void ssf_weight_grad_kernel(PlanView pv, const DevTile* __restrict__ tiles, int ntiles,
                                       int* __restrict__ counter, const double* __restrict__ atoms,
                                       const double* __restrict__ rab_inv,
                                       const double* __restrict__ dist_nearest, int natoms,
                                       const double* __restrict__ wf, double* __restrict__ grad) {
  // per warp: ld, lp (list distances / partition products), gw (3 natoms PRIVATE atom sums: within a warp every list entry
  // -- hence every atom -- is owned by one lane, so the sums need no atomics; shared-memory FP64 atomics are CAS loops) and
  // li (list -> atom); coordinates are read through li from the (L1-resident) atom array
  extern __shared__ double wg_dyn[];
  __shared__ int s_tile;
  const int nwarps = blockDim.x >> 5, warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  double* base = wg_dyn + (size_t)warp * 5 * natoms;
  double *ld = base, *lp = base + natoms, *gw = base + 2 * natoms;
  int* li = reinterpret_cast<int*>(wg_dyn + (size_t)nwarps * 5 * natoms) + (size_t)warp * natoms;
  for (int q = lane; q < 3 * natoms; q += 32) gw[q] = 0.;
  __syncthreads();
  const double kappa = (1. + magic_ssf) / (1. - magic_ssf) * (1. + 1e-9);
  const double bound = magic_ssf - 1.e-4, weight_tol = 1e-13, wf_thresh = 1.e-12;
  auto X = [&](int a, int c) { return atoms[3 * a + c]; };

  for (;;) {
    if (threadIdx.x == 0) s_tile = atomicAdd(counter, 1);
    __syncthreads();
    const int tile_idx = s_tile;
    __syncthreads();
    if (tile_idx >= ntiles) break;
    const DevTile tile = tiles[tile_idx];
    const int par = pv.tasks[tile.task].iParent;
    const double dist_cutoff = 0.5 * (1. - magic_ssf) * dist_nearest[par];
    const double ax = atoms[3 * par], ay = atoms[3 * par + 1], az = atoms[3 * par + 2];
    for (int i = warp; i < tile.npts; i += nwarps) {
      const int ip = tile.pt_off + i;
      const double wfi = wf[ip];
      if (fabs(wfi) < wf_thresh) continue;
      const double px = pv.px[ip], py = pv.py[ip], pz = pv.pz[ip];
      const double d_par = sqrt((px - ax) * (px - ax) + (py - ay) * (py - ay) + (pz - az) * (pz - az));
      if (d_par < dist_cutoff) continue;
      // nearest atom
      double dmin = d_par;
      for (int a = lane; a < natoms; a += 32) {
        const double dx = px - atoms[3 * a], dy = py - atoms[3 * a + 1], dz = pz - atoms[3 * a + 2];
        dmin = fmin(dmin, sqrt(dx * dx + dy * dy + dz * dz));
      }
#pragma unroll
      for (int d = 16; d > 0; d >>= 1) dmin = fmin(dmin, __shfl_xor_sync(0xffffffffu, dmin, d));
      // list1: atoms with r < kappa^2 r_min, compacted in index order
      const double r1 = kappa * kappa * dmin, r0 = kappa * dmin;
      int n1 = 0, kpar = -1;
      __syncwarp();
      for (int a0 = 0; a0 < natoms; a0 += 32) {
        const int a = a0 + lane;
        double d = 1e300;
        if (a < natoms) {
          const double x = atoms[3 * a], y = atoms[3 * a + 1], z = atoms[3 * a + 2];
          d = a == par ? d_par : sqrt((px - x) * (px - x) + (py - y) * (py - y) + (pz - z) * (pz - z));
        }
        const bool keep = a < natoms && (d < r1 || a == par);
        const unsigned m = __ballot_sync(0xffffffffu, keep);
        const int pos = n1 + __popc(m & ((1u << lane) - 1u));
        if (keep) { li[pos] = a; ld[pos] = d; }
        const unsigned mp = __ballot_sync(0xffffffffu, keep && a == par);
        if (mp) kpar = n1 + __popc(m & ((1u << (__ffs(mp) - 1)) - 1u));
        n1 += __popc(m);
      }
      __syncwarp();
      // partition products over list1
      double psum = 0.;
      for (int k = lane; k < n1; k += 32) {
        double P = 0.;
        const double dk = ld[k];
        if (dk < r0) {
          P = 1.;
          const double* __restrict__ rinv_k = rab_inv + (size_t)li[k] * natoms;  // 1 / R_AB from the plan's table
          for (int j = 0; j < n1; ++j) {
            if (j == k) continue;
            const double mu = (dk - ld[j]) * rinv_k[li[j]];
            if (mu >= magic_ssf) { P = 0.; break; }
            if (mu > -magic_ssf) P *= 0.5 * (1. - g_frisch_d(mu));
          }
        }
        lp[k] = P;
        psum += P;
      }
      const double sum = warp_sum(psum);
      __syncwarp();
      const double P_par = lp[kpar];
      double pa[3] = {0., 0., 0.};  // lane-local share of what the parent receives (translational invariance)
      // first term: - coef1 nabla_B mu_BA, B != parent; lane = owner of its list entries
      for (int k = lane; k < n1; k += 32) {
        if (k == kpar) continue;
        const int ab = li[k];
        const double bx = X(ab, 0), by = X(ab, 1), bz = X(ab, 2);
        const double ux = bx - ax, uy = by - ay, uz = bz - az;
        const double rAB_inv = rab_inv[(size_t)par * natoms + ab];
        const double dB = ld[k];
        const double mu_AB = (d_par - dB) * rAB_inv;
        if (fabs(mu_AB) < bound) {
          const double coef1 = t_frisch_d(mu_AB) * rAB_inv * (P_par - sum) / sum * wfi / dB;
          const double gx = coef1 * ((bx - px) + mu_AB * ux * rAB_inv * dB);
          const double gy = coef1 * ((by - py) + mu_AB * uy * rAB_inv * dB);
          const double gz = coef1 * ((bz - pz) + mu_AB * uz * rAB_inv * dB);
          gw[3 * ab] += gx; gw[3 * ab + 1] += gy; gw[3 * ab + 2] += gz;
          pa[0] -= gx; pa[1] -= gy; pa[2] -= gz;
        }
      }
      // second term: B with P_B > tol, C over list1 (lane = owner of entry kc)
      for (int kb = 0; kb < n1; ++kb) {
        const double PB = lp[kb];
        if (!(PB > weight_tol) || kb == kpar) continue;
        const int ab = li[kb];
        const double xb = X(ab, 0), yb = X(ab, 1), zb = X(ab, 2), dB = ld[kb];
        const double dB_inv = 1. / dB, ubx = (xb - px) * dB_inv, uby = (yb - py) * dB_inv, ubz = (zb - pz) * dB_inv;
        const double* __restrict__ rinv_b = rab_inv + (size_t)ab * natoms;
        const double pref = PB / sum * wfi;
        double gb[3] = {0., 0., 0.};
        for (int kc = lane; kc < n1; kc += 32) {
          if (kc == kb) continue;
          const int ac = li[kc];
          const double dC = ld[kc];
          const double Rinv = rinv_b[ac];
          const double mu_BC = (dB - dC) * Rinv;
          if (fabs(mu_BC) < bound) {
            const double xc = X(ac, 0), yc = X(ac, 1), zc = X(ac, 2);
            const double coef = pref * t_frisch_d(mu_BC) * Rinv;
            const double mr = mu_BC * Rinv;
            const double ex = (xb - xc) * mr, ey = (yb - yc) * mr, ez = (zb - zc) * mr;  // mu_BC (r_B - r_C) / R_BC
            gb[0] -= coef * (ubx - ex);
            gb[1] -= coef * (uby - ey);
            gb[2] -= coef * (ubz - ez);
            if (kc != kpar) {
              const double dC_inv = 1. / dC;
              const double cx = coef * ((xc - px) * dC_inv - ex);
              const double cy = coef * ((yc - py) * dC_inv - ey);
              const double cz = coef * ((zc - pz) * dC_inv - ez);
              gw[3 * ac] += cx; gw[3 * ac + 1] += cy; gw[3 * ac + 2] += cz;
              pa[0] -= cx; pa[1] -= cy; pa[2] -= cz;
            }
          }
        }
        const double vx = warp_sum(gb[0]), vy = warp_sum(gb[1]), vz = warp_sum(gb[2]);
        if (lane == (kb & 31)) {  // the owner lane of entry kb: per-lane program order keeps gw consistent
          gw[3 * ab] += vx; gw[3 * ab + 1] += vy; gw[3 * ab + 2] += vz;
          pa[0] -= vx; pa[1] -= vy; pa[2] -= vz;
        }
      }
      {
        const double vx = warp_sum(pa[0]), vy = warp_sum(pa[1]), vz = warp_sum(pa[2]);
        if (lane == (kpar & 31)) { gw[3 * par] += vx; gw[3 * par + 1] += vy; gw[3 * par + 2] += vz; }
      }
      __syncwarp();
    }
  }
  __syncwarp();
  for (int q = lane; q < 3 * natoms; q += 32) {
    const double v = gw[q];
    if (v != 0.) atomicAdd(&grad[q], v);
  }
}

bool g_sph_ready[64] = {};

void upload_sph_table() {
  int dev = 0;
  cudaGetDevice(&dev);
  if (dev >= 0 && dev < 64 && g_sph_ready[dev]) return;
  static SphTerm T[5][9][6];
  auto row = [&](int l, int m, std::initializer_list<SphTerm> ts) {
    int k = 0;
    for (auto t : ts) T[l][m][k++] = t;
    T[l][m][0].n = (signed char)ts.size();
  };
  for (auto& a : T) for (auto& b : a) for (auto& c : b) c = SphTerm{0, 0, 0, 0, 0.};
  const double s3 = 1.7320508075688772935;
  row(0, 0, {{0, 0, 0, 0, 1.}});
  row(1, 0, {{0, 1, 0, 0, 1.}}); row(1, 1, {{0, 0, 1, 0, 1.}}); row(1, 2, {{1, 0, 0, 0, 1.}});
  row(2, 0, {{1, 1, 0, 0, s3}});
  row(2, 1, {{0, 1, 1, 0, s3}});
  row(2, 2, {{0, 0, 2, 0, 1.}, {2, 0, 0, 0, -0.5}, {0, 2, 0, 0, -0.5}});
  row(2, 3, {{1, 0, 1, 0, s3}});
  row(2, 4, {{2, 0, 0, 0, 0.5 * s3}, {0, 2, 0, 0, -0.5 * s3}});
  const double s10 = 0.79056941504209483300, s15 = 3.8729833462074168852, s6 = 0.61237243569579452455,
               s15h = 1.9364916731037084426;
  row(3, 0, {{2, 1, 0, 0, 3 * s10}, {0, 3, 0, 0, -s10}});
  row(3, 1, {{1, 1, 1, 0, s15}});
  row(3, 2, {{0, 1, 2, 0, 4 * s6}, {2, 1, 0, 0, -s6}, {0, 3, 0, 0, -s6}});
  row(3, 3, {{0, 0, 3, 0, 1.0}, {2, 0, 1, 0, -1.5}, {0, 2, 1, 0, -1.5}});
  row(3, 4, {{1, 0, 2, 0, 4 * s6}, {3, 0, 0, 0, -s6}, {1, 2, 0, 0, -s6}});
  row(3, 5, {{2, 0, 1, 0, s15h}, {0, 2, 1, 0, -s15h}});
  row(3, 6, {{3, 0, 0, 0, s10}, {1, 2, 0, 0, -3 * s10}});
  const double s35h = 2.9580398915498080213, s70q = 2.0916500663351888699, s5h = 1.1180339887498948482,
               s10q = 0.79056941504209483300, s5q = 0.55901699437494742410, s35e = 0.73950997288745200532;
  row(4, 0, {{3, 1, 0, 0, s35h}, {1, 3, 0, 0, -s35h}});
  row(4, 1, {{2, 1, 1, 0, 3 * s70q}, {0, 3, 1, 0, -s70q}});
  row(4, 2, {{1, 1, 2, 0, 6 * s5h}, {3, 1, 0, 0, -s5h}, {1, 3, 0, 0, -s5h}});
  row(4, 3, {{0, 1, 3, 0, 4 * s10q}, {2, 1, 1, 0, -3 * s10q}, {0, 3, 1, 0, -3 * s10q}});
  row(4, 4, {{0, 0, 4, 0, 1.0}, {2, 0, 2, 0, -3.0}, {0, 2, 2, 0, -3.0}, {4, 0, 0, 0, 0.375}, {2, 2, 0, 0, 0.75},
             {0, 4, 0, 0, 0.375}});
  row(4, 5, {{1, 0, 3, 0, 4 * s10q}, {3, 0, 1, 0, -3 * s10q}, {1, 2, 1, 0, -3 * s10q}});
  row(4, 6, {{2, 0, 2, 0, 6 * s5q}, {0, 2, 2, 0, -6 * s5q}, {4, 0, 0, 0, -s5q}, {0, 4, 0, 0, s5q}});
  row(4, 7, {{3, 0, 1, 0, s70q}, {1, 2, 1, 0, -3 * s70q}});
  row(4, 8, {{4, 0, 0, 0, s35e}, {2, 2, 0, 0, -6 * s35e}, {0, 4, 0, 0, s35e}});
  cudaMemcpyToSymbol(c_sph, T, sizeof(T));
  if (dev >= 0 && dev < 64) g_sph_ready[dev] = true;
}

}  // namespace

void launch_collocation_hessian(const PlanView& pv, const DevTile* tiles, int ntiles, double* ws, cudaStream_t s) {
  if (ntiles <= 0) return;
  upload_sph_table();
  collocation_hessian_kernel<<<ntiles, TP, 0, s>>>(pv, tiles, ws);
}

cudaError_t launch_exc_grad(const PlanView& pv, const DevTile* tiles, int ntiles, int* counter, int nsm, double* ws,
                            FunctionalDesc func, bool gga, bool uks, int phase, const int* shell_atom, int natoms,
                            bool include_wd, double* wf_out, double* grad, cudaStream_t s) {
  if (ntiles <= 0) return cudaSuccess;
  const bool accumulates = phase != 0;
  const size_t dyn = accumulates ? (size_t)3 * natoms * sizeof(double) : 0;
  const bool smem_acc = accumulates && dyn <= 96 * 1024;
  const int ncta = std::min(ntiles, std::max(1, nsm) * 2);
  auto launch = [&](auto kern) {
    if (smem_acc) {  // static (up to 16 KB) + dynamic may pass the 48 KB default well before dyn does: always opt in
      cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)dyn);
      if (e != cudaSuccess) return e;
    }
    kern<<<ncta, EG_THREADS, smem_acc ? dyn : 0, s>>>(pv, tiles, ntiles, counter, ws, func, shell_atom, natoms,
                                                      include_wd ? 1 : 0, wf_out, grad, smem_acc ? 1 : 0);
    return cudaGetLastError();
  };
  if (!gga) return uks ? launch(exc_grad_kernel<false, true, 2>) : launch(exc_grad_kernel<false, false, 2>);
  if (phase == 0) return uks ? launch(exc_grad_kernel<true, true, 0>) : launch(exc_grad_kernel<true, false, 0>);
  return uks ? launch(exc_grad_kernel<true, true, 1>) : launch(exc_grad_kernel<true, false, 1>);
}

cudaError_t launch_ssf_weight_grad(const PlanView& pv, const DevTile* tiles, int ntiles, int* counter, int nsm,
                                   const double* atoms, const double* rab_inv, const double* dist_nearest, int natoms,
                                   const double* wf, double* grad, cudaStream_t s) {
  if (ntiles <= 0) return cudaSuccess;
  // shared memory per warp: 5 natoms doubles (ld, lp, 3 natoms private sums) + natoms ints
  auto bytes = [&](int nw) { return (size_t)natoms * (size_t)nw * (5 * 8 + 4) + 16; };
  int nw = 8;
  while (nw > 1 && bytes(nw) > 100 * 1024) nw >>= 1;
  const size_t dyn = bytes(nw);
  if (dyn > 227 * 1024) return cudaErrorInvalidConfiguration;  // > ~5 000 atoms: the caller reports NYI
  cudaError_t e = cudaFuncSetAttribute(ssf_weight_grad_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)dyn);
  if (e != cudaSuccess) return e;
  const int per_sm = (int)std::max<size_t>(1, std::min<size_t>(8, (200 * 1024) / dyn));
  const int ncta = std::min(ntiles, std::max(1, nsm) * per_sm);
  ssf_weight_grad_kernel<<<ncta, nw * 32, dyn, s>>>(pv, tiles, ntiles, counter, atoms, rab_inv, dist_nearest, natoms, wf,
                                                    grad);
  return cudaGetLastError();
}

}  // namespace gxb
