// Unpolarised LDA / GGA exchange-correlation kernels evaluated per grid point.
//
// The reference delegates this to ExchCXX (un-vendored, pinned at 67be5c6e in
// cmake/gauxc-dep-versions.cmake:10-11): host call
// reference_replicated_xc_host_integrator_exc_vxc.hpp:446-450, device call
// local_work_driver/device/cuda/xc_functional_eval_wrapper.cxx:17-31.  ExchCXX's builtin
// kernels transcribe libxc (lda_x, lda_c_vwn, lda_c_pw[mod], gga_x_pbe, gga_c_pbe); the
// closed forms below are those published functionals, validated against the reference's
// golden benzene SVWN5 / PBE0 EXC+VXC (tests/golden, tests/test_oracle_golden.py).
//
// Convention (libxc/ExchCXX): eps = energy per particle, vrho = d(rho eps)/d rho,
// vsigma = d(rho eps)/d sigma with sigma = |grad rho|^2, rho = total density.
#pragma once
#include "device_plan.hpp"
#include <cmath>

#ifdef __CUDACC__
#define GXB_HD __host__ __device__ __forceinline__
#else
#define GXB_HD inline
#endif

namespace gxb {

struct XcOut {
  double eps, vrho, vsigma;
};

// --- Slater exchange ------------------------------------------------------------
GXB_HD XcOut slater_x(double rho) {
  XcOut o{0., 0., 0.};
  if (rho <= 1e-24) return o;
  const double cx = -0.73855876638202240588;  // -(3/4)(3/pi)^(1/3)
  const double e = cx * cbrt(rho);
  o.eps = e;
  o.vrho = (4. / 3.) * e;
  return o;
}

// --- VWN5 correlation (libxc lda_c_vwn, paramagnetic set) -------------------------
GXB_HD XcOut vwn5_c(double rho) {
  XcOut o{0., 0., 0.};
  if (rho <= 1e-24) return o;
  // ExchCXX maps Kernel::VWN5 onto libxc's XC_LDA_C_VWN_RPA parameter set (pinned by the
  // golden benzene SVWN5 EXC/VXC): paramagnetic RPA fit
  const double A = 0.0310907, b = 13.0720, c = 42.7198, x0 = -0.409286;
  const double Q = sqrt(4. * c - b * b);
  const double X0 = x0 * x0 + b * x0 + c;
  const double rs = cbrt(0.75 / (M_PI * rho));
  const double x = sqrt(rs);
  const double X = rs + b * x + c;
  const double tx = 2. * x + b;
  const double at = atan(Q / tx);
  const double xm = x - x0;
  const double k0 = b * x0 / X0;
  const double e = A * (log(rs / X) + (2. * b / Q) * at -
                        k0 * (log(xm * xm / X) + (2. * (b + 2. * x0) / Q) * at));
  const double den = tx * tx + Q * Q;
  const double de_dx = A * (2. / x - tx / X - 4. * b / den -
                            k0 * (2. / xm - tx / X - 4. * (b + 2. * x0) / den));
  o.eps = e;
  o.vrho = e - (x / 6.) * de_dx;
  return o;
}

// --- PW92 (modified constants, as used inside PBE correlation) --------------------
GXB_HD void pw92_mod_unpol(double rs, double& e, double& de_drs) {
  const double A = 0.0310907, a1 = 0.21370, b1 = 7.5957, b2 = 3.5876, b3 = 1.6382, b4 = 0.49294;
  const double srs = sqrt(rs);
  const double G = 2. * A * (b1 * srs + b2 * rs + b3 * rs * srs + b4 * rs * rs);
  const double Gp = 2. * A * (0.5 * b1 / srs + b2 + 1.5 * b3 * srs + 2. * b4 * rs);
  const double L = log1p(1. / G);
  const double pre = -2. * A * (1. + a1 * rs);
  e = pre * L;
  de_drs = -2. * A * a1 * L - pre * Gp / (G * (G + 1.));
}

GXB_HD XcOut pw92_c(double rho) {
  XcOut o{0., 0., 0.};
  if (rho <= 1e-24) return o;
  const double rs = cbrt(0.75 / (M_PI * rho));
  double e, de;
  pw92_mod_unpol(rs, e, de);
  o.eps = e;
  o.vrho = e - (rs / 3.) * de;
  return o;
}

// --- PBE exchange -----------------------------------------------------------------
GXB_HD XcOut pbe_x(double rho, double sigma, double kappa, double mu) {
  XcOut o{0., 0., 0.};
  if (rho <= 1e-32) return o;
  sigma = fmax(sigma, 1e-40);
  const double cx = -0.73855876638202240588;
  const double r13 = cbrt(rho);
  const double r43 = rho * r13;
  // s^2 = sigma / (4 (3 pi^2)^(2/3) rho^(8/3))
  const double c2 = 4. * 9.5707800006273050;  // 4 (3 pi^2)^(2/3)
  const double s2 = sigma / (c2 * r43 * r43);
  const double d = 1. + mu * s2 / kappa;
  const double F = 1. + kappa - kappa / d;
  const double Fp = mu / (d * d);  // dF/d(s^2)
  o.eps = cx * r13 * F;
  o.vrho = cx * r13 * ((4. / 3.) * F - (8. / 3.) * s2 * Fp);
  o.vsigma = cx * r43 * Fp / (c2 * r43 * r43);
  return o;
}

// --- PBE correlation --------------------------------------------------------------
GXB_HD XcOut pbe_c(double rho, double sigma) {
  XcOut o{0., 0., 0.};
  if (rho <= 1e-12) return o;
  sigma = fmax(sigma, 1e-32);  // (dens_tol^(4/3))^2
  const double beta = 0.06672455060314922;
  const double gamma = 0.031090690869654895034;  // (1 - ln 2)/pi^2
  const double bg = beta / gamma;
  const double rs = cbrt(0.75 / (M_PI * rho));
  double ec, dec_drs;
  pw92_mod_unpol(rs, ec, dec_drs);
  const double dec_drho = -dec_drs * rs / (3. * rho);

  // t^2 = sigma pi / (16 (3 pi^2)^(1/3) rho^(7/3))
  const double r13 = cbrt(rho);
  const double r73 = rho * rho * r13;
  const double ct = M_PI / (16. * 3.0936677262801359310);  // (3 pi^2)^(1/3)
  const double y = ct * sigma / r73;
  const double dy_drho = -(7. / 3.) * y / rho;
  const double dy_dsigma = ct / r73;

  const double E = exp(-ec / gamma);
  const double Em1 = expm1(-ec / gamma);
  const double Aa = bg / Em1;
  const double dA_dec = bg * E / (gamma * Em1 * Em1);

  const double Ay = Aa * y;
  const double N = y + Ay * y;
  const double D = 1. + Ay + Ay * Ay;
  const double Qv = N / D;
  const double Q_y = ((1. + 2. * Ay) * D - N * (Aa + 2. * Aa * Ay)) / (D * D);
  const double Q_A = (y * y * D - N * (y + 2. * Ay * y)) / (D * D);
  const double arg = 1. + bg * Qv;
  const double H = gamma * log(arg);
  const double H_y = beta * Q_y / arg;
  const double H_A = beta * Q_A / arg;

  const double dH_drho = H_y * dy_drho + H_A * dA_dec * dec_drho;
  o.eps = ec + H;
  o.vrho = ec + H + rho * (dec_drho + dH_drho);
  o.vsigma = rho * H_y * dy_dsigma;
  return o;
}

// out of line on the device: the dual-number code is large, and inlined into the persistent kernel it costs
// the MMA loops instruction-cache hits (taxol fused 123 -> 130 ms); it runs once per grid point
// B88 exchange / LYP correlation at zeta = 0 go through the spin-resolved dual-number kernels
// (xc_functionals_pol_gga.cuh).  That code is large: a persistent kernel that carries it loses
// instruction-cache hits in its MMA loops (taxol PBE fused: 123 -> 130 ms inlined, 145 ms as an out-of-line
// call), so the kernels are instantiated twice (DUAL = false: closed forms only) and dispatched on the
// functional's kernel list.
#ifndef GXB_VIA_POL_INLINE
#define GXB_VIA_POL_INLINE __forceinline__
#endif
#ifdef __CUDACC__
__host__ __device__ GXB_VIA_POL_INLINE
#endif
XcOut eval_kernel_via_pol(int id, double rho, double sigma);

GXB_HD XcOut eval_kernel_closed(int id, double rho, double sigma) {
  switch (id) {
    case K_SLATER_X: return slater_x(rho);
    case K_VWN5_C: return vwn5_c(rho);
    case K_PW92_C: return pw92_c(rho);
    case K_PBE_X: return pbe_x(rho, sigma, 0.8040, 0.2195149727645171);
    case K_REVPBE_X: return pbe_x(rho, sigma, 1.245, 0.2195149727645171);  // libxc gga_x_pbe_r
    case K_PBE_C: return pbe_c(rho, sigma);
    default: return XcOut{0., 0., 0.};
  }
}

GXB_HD bool kernel_needs_dual(int id) { return id == K_B88_X || id == K_LYP_C; }
inline bool functional_needs_dual(const FunctionalDesc& f) {
  for (int k = 0; k < f.nkern; ++k)
    if (kernel_needs_dual(f.kern[k])) return true;
  return false;
}

template <bool DUAL>
GXB_HD XcOut eval_functional_t(const FunctionalDesc& f, double rho, double sigma) {
  XcOut t{0., 0., 0.};
  for (int k = 0; k < f.nkern; ++k) {
    XcOut o;
    if (DUAL && kernel_needs_dual(f.kern[k])) o = eval_kernel_via_pol(f.kern[k], rho, sigma);
    else o = eval_kernel_closed(f.kern[k], rho, sigma);
    t.eps += f.coeff[k] * o.eps;
    t.vrho += f.coeff[k] * o.vrho;
    t.vsigma += f.coeff[k] * o.vsigma;
  }
  return t;
}
GXB_HD XcOut eval_functional(const FunctionalDesc& f, double rho, double sigma) {
  return eval_functional_t<true>(f, rho, sigma);
}

// ---- spin-polarised LDA (UKS) -----------------------------------------------------------------
// Inputs rho_+ / rho_- (eval_uvvar_lda_uks, reference_local_host_work_driver.cxx:166-188); eps is the
// energy per particle of the total density, va / vb = d(rho eps)/d rho_+-.  Same published forms as the
// oracle (exact spin scaling of Slater exchange; libxc lda_c_vwn_rpa: para-/ferromagnetic RPA fits
// interpolated with f(zeta)), pinned by the reference's cytosine SVWN5 UKS fixture.
struct XcOutPol {
  double eps, va, vb;
};

GXB_HD XcOutPol slater_x_pol(double ra, double rb) {
  XcOutPol o{0., 0., 0.};
  const double rho = ra + rb;
  if (rho <= 1e-24) return o;
  const double cx = -0.73855876638202240588;  // -(3/4)(3/pi)^(1/3)
  double ea = 0., eb = 0.;
  if (ra > 0.) { const double t = cbrt(2. * ra); ea = cx * t * ra; o.va = (4. / 3.) * cx * t; }
  if (rb > 0.) { const double t = cbrt(2. * rb); eb = cx * t * rb; o.vb = (4. / 3.) * cx * t; }
  o.eps = (ea + eb) / rho;
  return o;
}

GXB_HD void vwn_fit(double A, double b, double c, double x0, double rs, double& e, double& de_drs) {
  const double Q = sqrt(4. * c - b * b);
  const double X0 = x0 * x0 + b * x0 + c;
  const double x = sqrt(rs);
  const double X = rs + b * x + c;
  const double tx = 2. * x + b;
  const double at = atan(Q / tx);
  const double xm = x - x0;
  const double k0 = b * x0 / X0;
  e = A * (log(rs / X) + (2. * b / Q) * at - k0 * (log(xm * xm / X) + (2. * (b + 2. * x0) / Q) * at));
  const double den = tx * tx + Q * Q;
  const double de_dx = A * (2. / x - tx / X - 4. * b / den - k0 * (2. / xm - tx / X - 4. * (b + 2. * x0) / den));
  de_drs = de_dx / (2. * x);
}

GXB_HD XcOutPol vwn5_c_pol(double ra, double rb) {
  XcOutPol o{0., 0., 0.};
  const double rho = ra + rb;
  if (rho <= 1e-24) return o;
  const double rs = cbrt(0.75 / (M_PI * rho));
  double z = (ra - rb) / rho;
  z = fmin(1., fmax(-1., z));
  double eP, dP, eF, dF;
  vwn_fit(0.0310907, 13.0720, 42.7198, -0.409286, rs, eP, dP);
  vwn_fit(0.01554535, 20.1231, 101.578, -0.743294, rs, eF, dF);
  const double den = 2. * cbrt(2.) - 2.;
  const double op = cbrt(1. + z), om = cbrt(1. - z);
  const double fz = (op * (1. + z) + om * (1. - z) - 2.) / den;
  const double dfz = (4. / 3.) * (op - om) / den;
  const double e = eP + (eF - eP) * fz;
  const double common = e - (rs / 3.) * (dP + (dF - dP) * fz);
  const double de_dz = (eF - eP) * dfz;
  o.eps = e;
  o.va = common + (1. - z) * de_dz;
  o.vb = common - (1. + z) * de_dz;
  return o;
}

// LDA kernels only (the reference's UKS GGA fixtures are BLYP, not restated here)
GXB_HD XcOutPol eval_functional_pol_lda(const FunctionalDesc& f, double ra, double rb) {
  XcOutPol t{0., 0., 0.};
  for (int k = 0; k < f.nkern; ++k) {
    XcOutPol o{0., 0., 0.};
    if (f.kern[k] == K_SLATER_X) o = slater_x_pol(ra, rb);
    else if (f.kern[k] == K_VWN5_C) o = vwn5_c_pol(ra, rb);
    t.eps += f.coeff[k] * o.eps;
    t.va += f.coeff[k] * o.va;
    t.vb += f.coeff[k] * o.vb;
  }
  return t;
}

}  // namespace gxb
