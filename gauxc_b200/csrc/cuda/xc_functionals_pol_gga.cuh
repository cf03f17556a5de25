// Spin-polarised GGA kernels for the UKS GGA path (B88 exchange, LYP correlation, PBE exchange and
// correlation to follow), evaluated with 5-partial forward-mode dual numbers: the energy density
// E(rho_a, rho_b, sigma_aa, sigma_ab, sigma_bb) of the published functional is written once and
// vrho[2] / vsigma[3] (the libxc / ExchCXX polarised GGA outputs the reference consumes in
// eval_zmat_gga_vxc_uks, reference_local_host_work_driver.cxx:715-773) fall out.  The functional is
// evaluated once per grid point (~10^3 flop against ~10^5 of the contractions), so the 6x arithmetic
// of the dual numbers is invisible next to the DMMA work.
//
// Used by the fused kernel for the UKS GGA path and, at zeta = 0, for the unpolarised BLYP / B3LYP kernels.
#pragma once
#include "xc_functionals.cuh"

namespace gxb {

struct Dual5 {
  double v;
  double d[5];
  GXB_HD Dual5(double x = 0.) : v(x) {
    for (int k = 0; k < 5; ++k) d[k] = 0.;
  }
};
GXB_HD Dual5 dual_var(double x, int k) {
  Dual5 r(x);
  r.d[k] = 1.;
  return r;
}
GXB_HD Dual5 dual_lift(const Dual5& a, double f, double fp) {
  Dual5 r(f);
  for (int k = 0; k < 5; ++k) r.d[k] = fp * a.d[k];
  return r;
}
GXB_HD Dual5 operator+(const Dual5& a, const Dual5& b) {
  Dual5 r(a.v + b.v);
  for (int k = 0; k < 5; ++k) r.d[k] = a.d[k] + b.d[k];
  return r;
}
GXB_HD Dual5 operator-(const Dual5& a, const Dual5& b) {
  Dual5 r(a.v - b.v);
  for (int k = 0; k < 5; ++k) r.d[k] = a.d[k] - b.d[k];
  return r;
}
GXB_HD Dual5 operator-(const Dual5& a) {
  Dual5 r(-a.v);
  for (int k = 0; k < 5; ++k) r.d[k] = -a.d[k];
  return r;
}
GXB_HD Dual5 operator*(const Dual5& a, const Dual5& b) {
  Dual5 r(a.v * b.v);
  for (int k = 0; k < 5; ++k) r.d[k] = a.d[k] * b.v + a.v * b.d[k];
  return r;
}
GXB_HD Dual5 operator/(const Dual5& a, const Dual5& b) {
  Dual5 r(a.v / b.v);
  for (int k = 0; k < 5; ++k) r.d[k] = (a.d[k] - r.v * b.d[k]) / b.v;
  return r;
}
GXB_HD Dual5 dual_pow(const Dual5& a, double p) { return dual_lift(a, pow(a.v, p), p * pow(a.v, p - 1.)); }
GXB_HD Dual5 dual_exp(const Dual5& a) {
  const double e = exp(a.v);
  return dual_lift(a, e, e);
}
GXB_HD Dual5 dual_sqrt(const Dual5& a) {
  const double q = sqrt(a.v);
  return dual_lift(a, q, 0.5 / q);
}
GXB_HD Dual5 dual_asinh(const Dual5& a) { return dual_lift(a, asinh(a.v), 1. / sqrt(1. + a.v * a.v)); }
GXB_HD Dual5 dual_log(const Dual5& a) { return dual_lift(a, log(a.v), 1. / a.v); }
GXB_HD Dual5 dual_log1p(const Dual5& a) { return dual_lift(a, log1p(a.v), 1. / (1. + a.v)); }
GXB_HD Dual5 dual_expm1(const Dual5& a) { return dual_lift(a, expm1(a.v), exp(a.v)); }

// Becke 1988: E_x = sum_s -rho_s^{4/3} [ C + beta x^2 / (1 + 6 beta x asinh x) ],  x = |grad rho_s| / rho_s^{4/3}
GXB_HD Dual5 b88_x_spin(const Dual5& r, const Dual5& s) {
  const double beta = 0.0042, C = 0.93052573634910002500;  // (3/2) (3 / (4 pi))^(1/3)
  if (r.v <= 1e-20) return Dual5(0.);
  const Dual5 r43 = dual_pow(r, 4. / 3.);
  if (s.v <= 1e-40) return -(r43 * Dual5(C));
  const Dual5 x = dual_sqrt(s) / r43;
  return -(r43 * (Dual5(C) + Dual5(beta) * x * x / (Dual5(1.) + Dual5(6. * beta) * x * dual_asinh(x))));
}

// Lee-Yang-Parr in the closed form of Miehlich, Savin, Stoll, Preuss, Chem. Phys. Lett. 157, 200 (1989)
GXB_HD Dual5 lyp_c_energy(const Dual5& ra, const Dual5& rb, const Dual5& saa, const Dual5& sab, const Dual5& sbb) {
  const double a = 0.04918, b = 0.132, c = 0.2533, d = 0.349;
  const double CF = 2.87123400018819181594;  // (3/10) (3 pi^2)^(2/3)
  const Dual5 rho = ra + rb;
  if (rho.v <= 1e-20) return Dual5(0.);
  const Dual5 rm13 = dual_pow(rho, -1. / 3.);
  const Dual5 den = Dual5(1.) + Dual5(d) * rm13;
  const Dual5 omega = dual_exp(-(Dual5(c) * rm13)) / den * dual_pow(rho, -11. / 3.);
  const Dual5 delta = Dual5(c) * rm13 + Dual5(d) * rm13 / den;
  const Dual5 sig = saa + Dual5(2.) * sab + sbb;
  const Dual5 rab = ra * rb;
  const Dual5 t1 = Dual5(12.69920841574560865 * CF) * (dual_pow(ra, 8. / 3.) + dual_pow(rb, 8. / 3.));  // 2^(11/3)
  const Dual5 t2 = (Dual5(47. / 18.) - Dual5(7. / 18.) * delta) * sig;
  const Dual5 t3 = (Dual5(2.5) - delta / Dual5(18.)) * (saa + sbb);
  const Dual5 t4 = (delta - Dual5(11.)) / Dual5(9.) * (ra / rho * saa + rb / rho * sbb);
  const Dual5 r2 = rho * rho;
  const Dual5 brace = rab * (t1 + t2 - t3 - t4) - Dual5(2. / 3.) * r2 * sig + (Dual5(2. / 3.) * r2 - ra * ra) * sbb +
                      (Dual5(2. / 3.) * r2 - rb * rb) * saa;
  return -(Dual5(a) * Dual5(4.) / den * rab / rho) - Dual5(a * b) * omega * brace;
}

enum PolGgaKernelId : int { PK_B88_X = 0, PK_LYP_C = 1 };

// unpolarised limit: E(rho, sigma) = E_pol(rho/2, rho/2, sigma/4, sigma/4, sigma/4), so
// vrho = dE/d rho_a and vsigma = (vaa + vab + vbb) / 4
#ifdef __CUDACC__
__host__ __device__ GXB_VIA_POL_INLINE
#else
inline
#endif
XcOut eval_kernel_via_pol(int id, double rho, double sigma) {
  XcOut o{0., 0., 0.};
  if (rho <= 1e-24) return o;
  const Dual5 ra = dual_var(fmax(0.5 * rho, 1e-30), 0), rb = dual_var(fmax(0.5 * rho, 1e-30), 1);
  const double q = 0.25 * fmax(sigma, 0.);
  const Dual5 saa = dual_var(q, 2), sab = dual_var(q, 3), sbb = dual_var(q, 4);
  Dual5 E(0.);
  if (id == K_B88_X) E = b88_x_spin(ra, saa) + b88_x_spin(rb, sbb);
  else if (id == K_LYP_C) E = lyp_c_energy(ra, rb, saa, sab, sbb);
  o.eps = E.v / rho;
  o.vrho = E.d[0];
  o.vsigma = 0.25 * (E.d[2] + E.d[3] + E.d[4]);
  return o;
}

struct XcOutPolGga {
  double eps, va, vb, vaa, vab, vbb;
};

// Perdew-Wang 92 correlation, spin-polarised, with the "modified" constants PBE correlation uses (libxc
// lda_c_pw_mod): G(rs) = -2A(1 + a1 rs) ln(1 + 1 / (2A(b1 rs^1/2 + b2 rs + b3 rs^3/2 + b4 rs^2)))
GXB_HD Dual5 pw92_G(const Dual5& rs, const Dual5& srs, double A, double a1, double b1, double b2, double b3,
                    double b4) {
  const Dual5 den = Dual5(2. * A) * (Dual5(b1) * srs + Dual5(b2) * rs + Dual5(b3) * rs * srs + Dual5(b4) * rs * rs);
  return -(Dual5(2. * A) * (Dual5(1.) + Dual5(a1) * rs) * dual_log1p(Dual5(1.) / den));
}

// PBE correlation, spin-polarised (Perdew, Burke, Ernzerhof 1996, eqs. 3, 7, 8): energy per volume
//   E = rho [ eps_c^PW92(rs, zeta) + H ],  H = gamma phi^3 ln(1 + beta/gamma t^2 (1 + A t^2) / (1 + A t^2 + A^2 t^4))
GXB_HD Dual5 pbe_c_pol_energy(const Dual5& ra, const Dual5& rb, const Dual5& saa, const Dual5& sab, const Dual5& sbb) {
  const double beta = 0.06672455060314922, gamma = 0.031090690869654895034;  // (1 - ln 2) / pi^2
  const double fz20 = 1.709920934161365617563962776245, cf = 1.92366105093153631981;  // 1 / (2^(4/3) - 2)
  const Dual5 rho = ra + rb;
  if (rho.v <= 1e-12) return Dual5(0.);
  const Dual5 rs = dual_pow(Dual5(0.75 / M_PI) / rho, 1. / 3.);
  const Dual5 srs = dual_sqrt(rs);
  Dual5 z = (ra - rb) / rho;
  // (1 +- zeta)^(2/3) has an unbounded slope at full polarisation: keep a hair away from it
  const double zlim = 1. - 1e-12;
  if (z.v > zlim) z = Dual5(zlim);
  if (z.v < -zlim) z = Dual5(-zlim);
  const Dual5 opz = Dual5(1.) + z, omz = Dual5(1.) - z;
  const Dual5 fz = (dual_pow(opz, 4. / 3.) + dual_pow(omz, 4. / 3.) - Dual5(2.)) * Dual5(cf);
  const Dual5 g0 = pw92_G(rs, srs, 0.0310907, 0.21370, 7.5957, 3.5876, 1.6382, 0.49294);
  const Dual5 g1 = pw92_G(rs, srs, 0.01554535, 0.20548, 14.1189, 6.1977, 3.3662, 0.62517);
  const Dual5 g2 = pw92_G(rs, srs, 0.0168869, 0.11125, 10.357, 3.6231, 0.88026, 0.49671);  // = -alpha_c
  const Dual5 z2 = z * z;
  const Dual5 z4 = z2 * z2;
  const Dual5 ec = g0 + z4 * fz * (g1 - g0 + g2 / Dual5(fz20)) - fz * g2 / Dual5(fz20);
  const Dual5 phi = Dual5(0.5) * (dual_pow(opz, 2. / 3.) + dual_pow(omz, 2. / 3.));
  const Dual5 phi3 = phi * phi * phi;
  Dual5 sig = saa + Dual5(2.) * sab + sbb;
  if (sig.v < 1e-32) sig = Dual5(1e-32);
  const double ct = M_PI / (16. * 3.0936677262801359310);  // pi / (16 (3 pi^2)^(1/3))
  const Dual5 t2 = Dual5(ct) * sig / (phi * phi * dual_pow(rho, 7. / 3.));
  const Dual5 Aa = Dual5(beta / gamma) / dual_expm1(-(ec / (Dual5(gamma) * phi3)));
  const Dual5 At2 = Aa * t2;
  const Dual5 H = Dual5(gamma) * phi3 *
                  dual_log1p(Dual5(beta / gamma) * t2 * (Dual5(1.) + At2) / (Dual5(1.) + At2 + At2 * At2));
  return rho * (ec + H);
}

// One kernel of the functional at one point, spin-resolved.  Slater / VWN use their closed forms, PBE exchange
// the exact spin scaling E_x[ra, rb] = (E_x[2 ra] + E_x[2 rb]) / 2 of the unpolarised closed form, B88 / LYP /
// PBE correlation the dual numbers.
GXB_HD XcOutPolGga eval_kernel_pol(int id, double rho_a, double rho_b, double s_aa, double s_ab, double s_bb) {
  XcOutPolGga o{0., 0., 0., 0., 0., 0.};
  const double rho = rho_a + rho_b;
  if (rho <= 1e-24) return o;
  if (id == K_SLATER_X || id == K_VWN5_C) {
    const XcOutPol l = (id == K_SLATER_X) ? slater_x_pol(rho_a, rho_b) : vwn5_c_pol(rho_a, rho_b);
    o.eps = l.eps; o.va = l.va; o.vb = l.vb;
    return o;
  }
  if (id == K_PBE_X || id == K_REVPBE_X) {
    const double kappa = (id == K_PBE_X) ? 0.8040 : 1.245, mu = 0.2195149727645171;
    const XcOut xa = pbe_x(2. * rho_a, 4. * s_aa, kappa, mu), xb = pbe_x(2. * rho_b, 4. * s_bb, kappa, mu);
    o.eps = (rho_a * xa.eps + rho_b * xb.eps) / rho;
    o.va = xa.vrho; o.vb = xb.vrho;
    o.vaa = 2. * xa.vsigma; o.vbb = 2. * xb.vsigma;
    return o;
  }
  const Dual5 ra = dual_var(fmax(rho_a, 1e-30), 0), rb = dual_var(fmax(rho_b, 1e-30), 1);
  const Dual5 saa = dual_var(fmax(s_aa, 0.), 2), sab = dual_var(s_ab, 3), sbb = dual_var(fmax(s_bb, 0.), 4);
  Dual5 E(0.);
  if (id == K_B88_X) E = b88_x_spin(ra, saa) + b88_x_spin(rb, sbb);
  else if (id == K_LYP_C) E = lyp_c_energy(ra, rb, saa, sab, sbb);
  else if (id == K_PBE_C) E = pbe_c_pol_energy(ra, rb, saa, sab, sbb);
  o.eps = E.v / rho;
  o.va = E.d[0]; o.vb = E.d[1];
  o.vaa = E.d[2]; o.vab = E.d[3]; o.vbb = E.d[4];
  return o;
}

GXB_HD XcOutPolGga eval_functional_pol(const FunctionalDesc& f, double rho_a, double rho_b, double s_aa,
                                       double s_ab, double s_bb) {
  XcOutPolGga t{0., 0., 0., 0., 0., 0.};
  for (int k = 0; k < f.nkern; ++k) {
    const XcOutPolGga o = eval_kernel_pol(f.kern[k], rho_a, rho_b, s_aa, s_ab, s_bb);
    const double c = f.coeff[k];
    t.eps += c * o.eps; t.va += c * o.va; t.vb += c * o.vb;
    t.vaa += c * o.vaa; t.vab += c * o.vab; t.vbb += c * o.vbb;
  }
  return t;
}

// sum_k coeff_k kernel_k at one point; gamma = (sigma_aa, sigma_ab, sigma_bb)
GXB_HD XcOutPolGga eval_pol_gga(int nkern, const int* kern, const double* coeff, double rho_a, double rho_b,
                                double s_aa, double s_ab, double s_bb) {
  XcOutPolGga o{0., 0., 0., 0., 0., 0.};
  const double rho = rho_a + rho_b;
  if (rho <= 1e-24) return o;
  const Dual5 ra = dual_var(fmax(rho_a, 1e-30), 0), rb = dual_var(fmax(rho_b, 1e-30), 1);
  const Dual5 saa = dual_var(fmax(s_aa, 0.), 2), sab = dual_var(s_ab, 3), sbb = dual_var(fmax(s_bb, 0.), 4);
  Dual5 E(0.);
  for (int k = 0; k < nkern; ++k) {
    Dual5 e(0.);
    if (kern[k] == PK_B88_X) e = b88_x_spin(ra, saa) + b88_x_spin(rb, sbb);
    else if (kern[k] == PK_LYP_C) e = lyp_c_energy(ra, rb, saa, sab, sbb);
    E = E + Dual5(coeff[k]) * e;
  }
  o.eps = E.v / rho;
  o.va = E.d[0]; o.vb = E.d[1];
  o.vaa = E.d[2]; o.vab = E.d[3]; o.vbb = E.d[4];
  return o;
}

}  // namespace gxb
