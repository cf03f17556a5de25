// ORACLE -- TEST INFRASTRUCTURE ONLY.  Nothing under gauxc_b200/ may include, link or call
// this file; only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl
// reference legs use it, and only as the checker / CPU baseline.
//
// CPU restatement (C++/OpenMP, plain C interface) of the reference's HOST execution space
// for the EXC/VXC hot path.  Each function cites the reference file:line it follows:
//   driver         src/xc_integrator/replicated/host/reference_replicated_xc_host_integrator_exc_vxc.hpp:107-601
//   primitives     src/xc_integrator/local_work_driver/host/reference_local_host_work_driver.cxx
//                  eval_xmat :123-146, eval_uvvar_lda_rks :150-163, eval_uvvar_gga_rks :242-268,
//                  eval_zmat_lda_vxc_rks :586-604, eval_zmat_gga_vxc_rks :678-713, inc_vxc :1678-1692
//   gather/scatter src/xc_integrator/local_work_driver/host/util.hpp:21-168
//   cut map        src/xc_integrator/integrator_util/integrator_common.cxx:22-146
//   collocation    src/xc_integrator/local_work_driver/host/reference/gau2grid_collocation.cxx:25-116
//                  over gau2grid (external/gau2grid, CCA orderings); restated here and checked
//                  against gau2grid itself compiled into oracle/_ref (tests/test_oracle_golden.py)
//   SSF weights    src/xc_integrator/local_work_driver/host/reference/weights.cxx:117-237
//   functionals    ExchCXX (un-vendored, pinned 67be5c6e, cmake/gauxc-dep-versions.cmake:10-11):
//                  published Slater / VWN5 / PW92(mod) / PBE closed forms, written independently of
//                  the product's xc_functionals.cuh
// Parity is PINNED: golden vectors of the reference's own tests (tests/golden/*.npz converted
// from tests/ref_data/*.hdf5) are reproduced by tests/test_oracle_golden.py.
#include <algorithm>
#include <array>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <dlfcn.h>
#include <string>
#include <vector>
#ifdef _OPENMP
#include <omp.h>
#endif

namespace {

// ----------------------------------------------------------------------------------
// BLAS: OpenBLAS found at run time (the reference Host path calls a vendor BLAS), else
// straightforward loops.
// ----------------------------------------------------------------------------------
typedef void (*dgemm_t)(const char*, const char*, const int*, const int*, const int*, const double*,
                        const double*, const int*, const double*, const int*, const double*, double*,
                        const int*);
typedef void (*dsyr2k_t)(const char*, const char*, const int*, const int*, const double*, const double*,
                         const int*, const double*, const int*, const double*, double*, const int*);
dgemm_t f_dgemm = nullptr;
dsyr2k_t f_dsyr2k = nullptr;
char blas_name[512] = "builtin-loops";

void naive_gemm_nn(int m, int n, int k, double alpha, const double* A, int lda, const double* B, int ldb,
                   double* C, int ldc) {
  // C(m x n) = alpha * A(m x k) B(k x n), column major
  for (int j = 0; j < n; ++j) {
    double* c = C + (size_t)j * ldc;
    for (int i = 0; i < m; ++i) c[i] = 0.;
    for (int p = 0; p < k; ++p) {
      const double b = alpha * B[p + (size_t)j * ldb];
      const double* a = A + (size_t)p * lda;
      for (int i = 0; i < m; ++i) c[i] += a[i] * b;
    }
  }
}
void naive_syr2k_ln(int n, int k, const double* A, int lda, const double* B, int ldb, double* C, int ldc) {
  // lower(C) = A B^T + B A^T, A,B n x k
  for (int j = 0; j < n; ++j)
    for (int i = j; i < n; ++i) C[i + (size_t)j * ldc] = 0.;
  for (int p = 0; p < k; ++p) {
    const double* a = A + (size_t)p * lda;
    const double* b = B + (size_t)p * ldb;
    for (int j = 0; j < n; ++j) {
      const double aj = a[j], bj = b[j];
      double* c = C + (size_t)j * ldc;
      for (int i = j; i < n; ++i) c[i] += a[i] * bj + b[i] * aj;
    }
  }
}

void gemm_nn(int m, int n, int k, double alpha, const double* A, int lda, const double* B, int ldb,
             double* C, int ldc) {
  if (f_dgemm) {
    const double beta = 0.;
    f_dgemm("N", "N", &m, &n, &k, &alpha, A, &lda, B, &ldb, &beta, C, &ldc);
  } else naive_gemm_nn(m, n, k, alpha, A, lda, B, ldb, C, ldc);
}
void syr2k_ln(int n, int k, const double* A, int lda, const double* B, int ldb, double* C, int ldc) {
  if (f_dsyr2k) {
    const double one = 1., zero = 0.;
    f_dsyr2k("L", "N", &n, &k, &one, A, &lda, B, &ldb, &zero, C, &ldc);
  } else naive_syr2k_ln(n, k, A, lda, B, ldb, C, ldc);
}

// ----------------------------------------------------------------------------------
// functionals (independent derivation; unpolarised)
// ----------------------------------------------------------------------------------
const double PI = 3.14159265358979323846;

void f_slater(double rho, double& e, double& v) {
  if (rho <= 1e-24) { e = v = 0; return; }
  // eps_x = -3/4 (3/pi)^(1/3) rho^(1/3)
  e = -0.75 * std::cbrt(3. / PI) * std::cbrt(rho);
  v = 4. * e / 3.;
}

// VWN5: eps_c(x), x = sqrt(rs); v = eps - rs/3 d eps/d rs
void f_vwn5(double rho, double& e, double& v) {
  if (rho <= 1e-24) { e = v = 0; return; }
  // ExchCXX maps Kernel::VWN5 onto libxc's XC_LDA_C_VWN_RPA parameter set (pinned by the
  // golden benzene SVWN5 EXC/VXC): paramagnetic RPA fit
  const double A = 0.0310907, b = 13.0720, c = 42.7198, x0 = -0.409286;
  const double rs = std::cbrt(3. / (4. * PI * rho));
  const double x = std::sqrt(rs);
  auto Xf = [&](double y) { return y * y + b * y + c; };
  const double Q = std::sqrt(4. * c - b * b);
  auto eps = [&](double xx) {
    const double X = Xf(xx);
    const double at = std::atan(Q / (2. * xx + b));
    return A * (std::log(xx * xx / X) + 2. * b / Q * at -
                b * x0 / Xf(x0) * (std::log((xx - x0) * (xx - x0) / X) + 2. * (b + 2. * x0) / Q * at));
  };
  e = eps(x);
  // analytic derivative: d eps/dx = A [ c (x - x0) - b x x0 ] / [ x X(x) (x - x0) ] * ... (closed form below)
  // d/dx of each term
  const double X = Xf(x), Xp = 2. * x + b;
  const double dat = -2. * Q / (Xp * Xp + Q * Q);
  const double de = A * ((2. / x - Xp / X) + 2. * b / Q * dat -
                         b * x0 / Xf(x0) * ((2. / (x - x0) - Xp / X) + 2. * (b + 2. * x0) / Q * dat));
  v = e - rs / 3. * de / (2. * x);
}

void pw92mod(double rs, double& e, double& de) {
  const double A = 0.0310907, a1 = 0.21370, b1 = 7.5957, b2 = 3.5876, b3 = 1.6382, b4 = 0.49294;
  const double q0 = -2. * A * (1. + a1 * rs);
  const double rs12 = std::sqrt(rs), rs32 = rs * rs12;
  const double q1 = 2. * A * (b1 * rs12 + b2 * rs + b3 * rs32 + b4 * rs * rs);
  const double q1p = A * (b1 / rs12 + 2. * b2 + 3. * b3 * rs12 + 4. * b4 * rs);
  const double lg = std::log(1. + 1. / q1);
  e = q0 * lg;
  de = -2. * A * a1 * lg - q0 * q1p / (q1 * q1 + q1);
}
void f_pw92(double rho, double& e, double& v) {
  if (rho <= 1e-24) { e = v = 0; return; }
  const double rs = std::cbrt(3. / (4. * PI * rho));
  double de;
  pw92mod(rs, e, de);
  v = e - rs / 3. * de;
}

// kappa = 0.8040: PBE (libxc gga_x_pbe); kappa = 1.245: revPBE (gga_x_pbe_r, Zhang & Yang 1998)
void f_pbe_x(double rho, double sigma, double& e, double& v, double& vs, double kappa = 0.8040) {
  if (rho <= 1e-32) { e = v = vs = 0; return; }
  sigma = std::max(sigma, 1e-40);
  const double mu = 0.2195149727645171;
  const double kf = std::cbrt(3. * PI * PI * rho);
  const double s = std::sqrt(sigma) / (2. * kf * rho);
  const double s2 = s * s;
  const double Fx = 1. + kappa - kappa / (1. + mu * s2 / kappa);
  const double dF_ds2 = mu / ((1. + mu * s2 / kappa) * (1. + mu * s2 / kappa));
  const double ex_lda = -0.75 * std::cbrt(3. / PI) * std::cbrt(rho);
  e = ex_lda * Fx;
  // E = rho ex_lda Fx ; s2 ~ sigma rho^(-8/3)
  v = (4. / 3.) * ex_lda * Fx + rho * ex_lda * dF_ds2 * (-8. / 3.) * s2 / rho;
  vs = rho * ex_lda * dF_ds2 * s2 / sigma;
}

void f_pbe_c(double rho, double sigma, double& e, double& v, double& vs) {
  if (rho <= 1e-12) { e = v = vs = 0; return; }
  sigma = std::max(sigma, 1e-32);
  const double beta = 0.06672455060314922, gamma = (1. - std::log(2.)) / (PI * PI);
  const double rs = std::cbrt(3. / (4. * PI * rho));
  double ec, dec;
  pw92mod(rs, ec, dec);
  const double kf = std::cbrt(3. * PI * PI * rho);
  const double ks = std::sqrt(4. * kf / PI);
  const double t = std::sqrt(sigma) / (2. * ks * rho);
  const double t2 = t * t, t4 = t2 * t2;
  const double ex = std::exp(-ec / gamma);
  const double Ac = beta / gamma / (ex - 1.);
  const double num = 1. + Ac * t2;
  const double den = 1. + Ac * t2 + Ac * Ac * t4;
  const double arg = 1. + beta / gamma * t2 * num / den;
  const double H = gamma * std::log(arg);
  e = ec + H;
  // partial derivatives of H wrt t2 and Ac
  const double dq_dt2 = (num / den) + t2 * (Ac * den - num * (Ac + 2. * Ac * Ac * t2)) / (den * den);
  const double dq_dA = t2 * (t2 * den - num * (t2 + 2. * Ac * t4)) / (den * den);
  const double dH_dt2 = beta * dq_dt2 / arg;
  const double dH_dA = beta * dq_dA / arg;
  const double dA_dec = beta / gamma * ex / (gamma * (ex - 1.) * (ex - 1.));
  const double drs_drho = -rs / (3. * rho);
  const double dec_drho = dec * drs_drho;
  const double dt2_drho = -7. / 3. * t2 / rho;
  const double dH_drho = dH_dt2 * dt2_drho + dH_dA * dA_dec * dec_drho;
  v = e + rho * (dec_drho + dH_drho);
  vs = rho * dH_dt2 * t2 / sigma;
}

enum { K_SLATER_X = 0, K_VWN5_C = 1, K_PBE_X = 2, K_PBE_C = 3, K_VWN3_C = 4, K_PW92_C = 5, K_B88_X = 6, K_LYP_C = 7,
       K_REVPBE_X = 8 };

struct Func {
  int nkern, is_gga;
  int kern[4];
  double coeff[4];
};

void eval_func_pol_gga(const Func& f, int npts, const double* rho2, const double* gamma3, double* eps,
                       double* vrho2, double* vgamma3);

// B88 / LYP for a closed shell: the spin-resolved functional (pinned by the reference's BLYP UKS fixture) at
// rho_a = rho_b = rho/2, sigma_aa = sigma_ab = sigma_bb = sigma/4; d/d rho = d/d rho_a, d/d sigma = the mean
// of the three sigma derivatives... times 1 (chain rule: each of the three carries 1/4)
void f_via_pol(int kern, double rho, double sigma, double& e, double& v, double& vs) {
  Func one{1, 1, {kern, 0, 0, 0}, {1., 0., 0., 0.}};
  const double r2[2] = {0.5 * rho, 0.5 * rho};
  const double q = 0.25 * std::max(sigma, 0.);
  const double g3[3] = {q, q, q};
  double ee, v2[2], v3[3];
  eval_func_pol_gga(one, 1, r2, g3, &ee, v2, v3);
  e = ee;
  v = v2[0];
  vs = 0.25 * (v3[0] + v3[1] + v3[2]);
}

void eval_func(const Func& f, int npts, const double* rho, const double* sigma, double* eps, double* vrho,
               double* vsigma) {
  for (int i = 0; i < npts; ++i) {
    double E = 0, V = 0, S = 0;
    for (int k = 0; k < f.nkern; ++k) {
      double e = 0, v = 0, s = 0;
      switch (f.kern[k]) {
        case K_REVPBE_X: f_pbe_x(rho[i], sigma[i], e, v, s, 1.245); break;
        case K_B88_X:
        case K_LYP_C: f_via_pol(f.kern[k], rho[i], sigma ? sigma[i] : 0., e, v, s); break;
        case K_SLATER_X: f_slater(rho[i], e, v); break;
        case K_VWN5_C: f_vwn5(rho[i], e, v); break;
        case K_PW92_C: f_pw92(rho[i], e, v); break;
        case K_PBE_X: f_pbe_x(rho[i], sigma[i], e, v, s); break;
        case K_PBE_C: f_pbe_c(rho[i], sigma[i], e, v, s); break;
      }
      E += f.coeff[k] * e; V += f.coeff[k] * v; S += f.coeff[k] * s;
    }
    eps[i] = E; vrho[i] = V;
    if (vsigma) vsigma[i] = S;
  }
}

// ---- spin-polarised LDA (UKS) -------------------------------------------
// Inputs rho_+ / rho_- (eval_uvvar_lda_uks, reference_local_host_work_driver.cxx:166-188), outputs
// eps (energy per particle of the TOTAL density) and d(rho eps)/d rho_sigma, the interleaved layout
// ExchCXX / libxc use for polarised LDA.  Pinned by cytosine_svwn5_cc-pvdz_ufg_ssf_robust_uks.
void f_slater_pol(double ra, double rb, double& e, double& va, double& vb) {
  const double rho = ra + rb;
  if (rho <= 1e-24) { e = va = vb = 0; return; }
  // exact spin scaling E_x[ra, rb] = (E_x[2 ra] + E_x[2 rb]) / 2
  const double cx = -0.75 * std::cbrt(3. / PI);
  auto per_spin = [&](double r, double& en, double& v) {
    if (r <= 0.) { en = v = 0; return; }
    const double t = std::cbrt(2. * r);
    en = cx * t * r;        // (1/2) cx (2 r)^(4/3)
    v = 4. / 3. * cx * t;   // d en / d r
  };
  double ea, eb;
  per_spin(ra, ea, va);
  per_spin(rb, eb, vb);
  e = (ea + eb) / rho;
}

// one VWN fit eps(x = sqrt(rs)) and d eps/d rs for a parameter set (A, b, c, x0)
void vwn_fit(double A, double b, double c, double x0, double rs, double& e, double& de_drs) {
  const double x = std::sqrt(rs);
  auto Xf = [&](double y) { return y * y + b * y + c; };
  const double Q = std::sqrt(4. * c - b * b);
  const double X = Xf(x), Xp = 2. * x + b;
  const double at = std::atan(Q / Xp);
  e = A * (std::log(x * x / X) + 2. * b / Q * at -
           b * x0 / Xf(x0) * (std::log((x - x0) * (x - x0) / X) + 2. * (b + 2. * x0) / Q * at));
  const double dat = -2. * Q / (Xp * Xp + Q * Q);
  const double de_dx = A * ((2. / x - Xp / X) + 2. * b / Q * dat -
                            b * x0 / Xf(x0) * ((2. / (x - x0) - Xp / X) + 2. * (b + 2. * x0) / Q * dat));
  de_drs = de_dx / (2. * x);
}
// libxc XC_LDA_C_VWN_RPA (what ExchCXX's Kernel::VWN5 evaluates, see f_vwn5): para- and ferromagnetic
// RPA fits interpolated with f(zeta) = ((1+z)^(4/3) + (1-z)^(4/3) - 2) / (2^(4/3) - 2)
void f_vwn5_pol(double ra, double rb, double& e, double& va, double& vb) {
  const double rho = ra + rb;
  if (rho <= 1e-24) { e = va = vb = 0; return; }
  const double rs = std::cbrt(3. / (4. * PI * rho));
  double z = (ra - rb) / rho;
  z = std::min(1., std::max(-1., z));
  double eP, dP, eF, dF;
  vwn_fit(0.0310907, 13.0720, 42.7198, -0.409286, rs, eP, dP);
  vwn_fit(0.01554535, 20.1231, 101.578, -0.743294, rs, eF, dF);
  const double den = std::cbrt(2.) * 2. - 2.;
  const double op = std::cbrt(1. + z), om = std::cbrt(1. - z);
  const double fz = (op * (1. + z) + om * (1. - z) - 2.) / den;
  const double dfz = 4. / 3. * (op - om) / den;
  e = eP + (eF - eP) * fz;
  const double de_drs = dP + (dF - dP) * fz;
  const double de_dz = (eF - eP) * dfz;
  // d(rho e)/d rho_sigma = e - rs/3 de/drs + (+-1 - z) de/dz
  const double common = e - rs / 3. * de_drs;
  va = common + (1. - z) * de_dz;
  vb = common - (1. + z) * de_dz;
}

void eval_func_pol_lda(const Func& f, int npts, const double* rho2, double* eps, double* vrho2) {
  for (int i = 0; i < npts; ++i) {
    double E = 0, VA = 0, VB = 0;
    for (int k = 0; k < f.nkern; ++k) {
      double e = 0, va = 0, vb = 0;
      switch (f.kern[k]) {
        case K_SLATER_X: f_slater_pol(rho2[2 * i], rho2[2 * i + 1], e, va, vb); break;
        case K_VWN5_C: f_vwn5_pol(rho2[2 * i], rho2[2 * i + 1], e, va, vb); break;
        default: e = va = vb = std::nan(""); break;  // not restated for UKS yet
      }
      E += f.coeff[k] * e; VA += f.coeff[k] * va; VB += f.coeff[k] * vb;
    }
    eps[i] = E; vrho2[2 * i] = VA; vrho2[2 * i + 1] = VB;
  }
}

// ---- spin-polarised GGA by forward-mode differentiation (UKS GGA) ---------------
// The oracle only needs to be right, not fast: the energy density E(rho_a, rho_b, sigma_aa, sigma_ab,
// sigma_bb) of the published functional is written once on a 5-partial dual number and the
// derivatives libxc / ExchCXX return (vrho[2], vsigma[3]) fall out.  B88 exchange (Becke 1988) and
// LYP correlation (Miehlich et al. 1989 form) = the reference's UKS GGA fixture
// cytosine_blyp_cc-pvdz_ufg_ssf_robust_uks (tests/xc_integrator.cxx:468-472).
struct D5 {
  double v;
  double d[5];
  D5(double x = 0.) : v(x) { for (double& q : d) q = 0.; }
  static D5 var(double x, int k) { D5 r(x); r.d[k] = 1.; return r; }
};
inline D5 lift(const D5& a, double f, double fp) {  // f(a), f'(a)
  D5 r(f);
  for (int k = 0; k < 5; ++k) r.d[k] = fp * a.d[k];
  return r;
}
inline D5 operator+(const D5& a, const D5& b) { D5 r(a.v + b.v); for (int k = 0; k < 5; ++k) r.d[k] = a.d[k] + b.d[k]; return r; }
inline D5 operator-(const D5& a, const D5& b) { D5 r(a.v - b.v); for (int k = 0; k < 5; ++k) r.d[k] = a.d[k] - b.d[k]; return r; }
inline D5 operator-(const D5& a) { D5 r(-a.v); for (int k = 0; k < 5; ++k) r.d[k] = -a.d[k]; return r; }
inline D5 operator*(const D5& a, const D5& b) { D5 r(a.v * b.v); for (int k = 0; k < 5; ++k) r.d[k] = a.d[k] * b.v + a.v * b.d[k]; return r; }
inline D5 operator/(const D5& a, const D5& b) { D5 r(a.v / b.v); for (int k = 0; k < 5; ++k) r.d[k] = (a.d[k] - r.v * b.d[k]) / b.v; return r; }
inline D5 dpow(const D5& a, double p) { return lift(a, std::pow(a.v, p), p * std::pow(a.v, p - 1.)); }
inline D5 dexp(const D5& a) { const double e = std::exp(a.v); return lift(a, e, e); }
inline D5 dsqrt(const D5& a) { const double q = std::sqrt(a.v); return lift(a, q, 0.5 / q); }
inline D5 dasinh(const D5& a) { return lift(a, std::asinh(a.v), 1. / std::sqrt(1. + a.v * a.v)); }

// E_x^{B88} = sum_sigma -rho_s^{4/3} [ C + beta x^2 / (1 + 6 beta x asinh x) ],  x = |grad rho_s| / rho_s^{4/3}
D5 b88_energy(const D5& ra, const D5& rb, const D5& saa, const D5& sbb) {
  const double beta = 0.0042, C = 1.5 * std::cbrt(3. / (4. * PI));
  auto spin = [&](const D5& r, const D5& s) {
    if (r.v <= 1e-20) return D5(0.);
    const D5 r43 = dpow(r, 4. / 3.);
    if (s.v <= 1e-40) return -(r43 * D5(C));
    const D5 x = dsqrt(s) / r43;
    return -(r43 * (D5(C) + D5(beta) * x * x / (D5(1.) + D5(6. * beta) * x * dasinh(x))));
  };
  return spin(ra, saa) + spin(rb, sbb);
}

// LYP, Miehlich-Savin-Stoll-Preuss closed form (Chem. Phys. Lett. 157, 200 (1989), eq. 2)
D5 lyp_energy(const D5& ra, const D5& rb, const D5& saa, const D5& sab, const D5& sbb) {
  const double a = 0.04918, b = 0.132, c = 0.2533, d = 0.349;
  const double CF = 0.3 * std::pow(3. * PI * PI, 2. / 3.);
  const D5 rho = ra + rb;
  if (rho.v <= 1e-20) return D5(0.);
  const D5 rm13 = dpow(rho, -1. / 3.);
  const D5 den = D5(1.) + D5(d) * rm13;
  const D5 omega = dexp(-(D5(c) * rm13)) / den * dpow(rho, -11. / 3.);
  const D5 delta = D5(c) * rm13 + D5(d) * rm13 / den;
  const D5 sig = saa + D5(2.) * sab + sbb;
  const D5 rab = ra * rb;
  const D5 t1 = D5(std::pow(2., 11. / 3.) * CF) * (dpow(ra, 8. / 3.) + dpow(rb, 8. / 3.));
  const D5 t2 = (D5(47. / 18.) - D5(7. / 18.) * delta) * sig;
  const D5 t3 = (D5(2.5) - delta / D5(18.)) * (saa + sbb);
  const D5 t4 = (delta - D5(11.)) / D5(9.) * (ra / rho * saa + rb / rho * sbb);
  const D5 r2 = rho * rho;
  const D5 brace = rab * (t1 + t2 - t3 - t4) - D5(2. / 3.) * r2 * sig + (D5(2. / 3.) * r2 - ra * ra) * sbb +
                   (D5(2. / 3.) * r2 - rb * rb) * saa;
  return -(D5(a) * D5(4.) / den * rab / rho) - D5(a * b) * omega * brace;
}

inline D5 dlog(const D5& a) { return lift(a, std::log(a.v), 1. / a.v); }

// PBE correlation for a spin-polarised density, written as in the paper (Perdew, Burke, Ernzerhof, PRL 77,
// 3865 (1996), eqs. 3, 7, 8) on top of PW92 (Perdew & Wang, PRB 45, 13244 (1992), eqs. 8-10, the "modified"
// higher-precision constants libxc / ExchCXX use inside PBE):
//   eps_c(rs, zeta) = eps0 + alpha_c f(zeta)/f''(0) (1 - zeta^4) + (eps1 - eps0) f(zeta) zeta^4,
//   H = gamma phi^3 ln{1 + (beta/gamma) t^2 [1 + A t^2] / [1 + A t^2 + A^2 t^4]},
//   t = |grad n| / (2 phi k_s n),  k_s = sqrt(4 k_F / pi),  A = (beta/gamma) / (exp(-eps_c / (gamma phi^3)) - 1)
D5 pbe_c_pol_energy(const D5& na, const D5& nb, const D5& saa, const D5& sab, const D5& sbb) {
  const D5 n = na + nb;
  if (n.v <= 1e-12) return D5(0.);
  const double beta = 0.06672455060314922, gamma = (1. - std::log(2.)) / (PI * PI);
  const D5 rs = dpow(D5(3. / (4. * PI)) / n, 1. / 3.);
  auto Gpw = [&](double A, double a1, double b1, double b2, double b3, double b4) {
    const D5 q1 = D5(2. * A) * (D5(b1) * dsqrt(rs) + D5(b2) * rs + D5(b3) * dpow(rs, 1.5) + D5(b4) * rs * rs);
    return D5(-2. * A) * (D5(1.) + D5(a1) * rs) * dlog(D5(1.) + D5(1.) / q1);
  };
  const D5 e0 = Gpw(0.0310907, 0.21370, 7.5957, 3.5876, 1.6382, 0.49294);
  const D5 e1 = Gpw(0.01554535, 0.20548, 14.1189, 6.1977, 3.3662, 0.62517);
  const D5 mac = Gpw(0.0168869, 0.11125, 10.357, 3.6231, 0.88026, 0.49671);  // -alpha_c
  D5 zeta = (na - nb) / n;
  const double zmax = 1. - 1e-12;
  if (zeta.v > zmax) zeta = D5(zmax);
  if (zeta.v < -zmax) zeta = D5(-zmax);
  const D5 up = D5(1.) + zeta, dn = D5(1.) - zeta;
  const double fpp0 = 1.709920934161365617563962776245;  // f''(0) = 8 / (9 (2^{4/3} - 2))
  const D5 fz = (dpow(up, 4. / 3.) + dpow(dn, 4. / 3.) - D5(2.)) / D5(std::pow(2., 4. / 3.) - 2.);
  const D5 z4 = zeta * zeta * zeta * zeta;
  const D5 ec = e0 - mac * fz / D5(fpp0) * (D5(1.) - z4) + (e1 - e0) * fz * z4;
  const D5 phi = (dpow(up, 2. / 3.) + dpow(dn, 2. / 3.)) / D5(2.);
  const D5 phi3 = phi * phi * phi;
  D5 grad2 = saa + D5(2.) * sab + sbb;
  if (grad2.v < 1e-32) grad2 = D5(1e-32);
  const D5 kF = dpow(D5(3. * PI * PI) * n, 1. / 3.);
  const D5 ks2 = D5(4. / PI) * kF;
  const D5 t2 = grad2 / (D5(4.) * phi * phi * ks2 * n * n);
  const D5 A = D5(beta / gamma) / (dexp(-(ec / (D5(gamma) * phi3))) - D5(1.));
  const D5 At2 = A * t2;
  const D5 H = D5(gamma) * phi3 * dlog(D5(1.) + D5(beta / gamma) * t2 * (D5(1.) + At2) / (D5(1.) + At2 + At2 * At2));
  return n * (ec + H);
}

// gamma = (sigma_aa, sigma_ab, sigma_bb) interleaved per point like the reference (eval_uvvar_gga_uks)
void eval_func_pol_gga(const Func& f, int npts, const double* rho2, const double* gamma3, double* eps,
                       double* vrho2, double* vgamma3) {
  for (int i = 0; i < npts; ++i) {
    // densities are clipped away from zero so that rho^(negative power) stays finite; such points
    // carry weights * rho ~ 0 anyway
    const D5 ra = D5::var(std::max(rho2[2 * i], 1e-30), 0), rb = D5::var(std::max(rho2[2 * i + 1], 1e-30), 1);
    const D5 saa = D5::var(std::max(gamma3[3 * i], 0.), 2), sab = D5::var(gamma3[3 * i + 1], 3),
             sbb = D5::var(std::max(gamma3[3 * i + 2], 0.), 4);
    D5 E(0.);
    for (int k = 0; k < f.nkern; ++k) {
      D5 e(0.);
      switch (f.kern[k]) {
        case K_B88_X: e = b88_energy(ra, rb, saa, sbb); break;
        case K_LYP_C: e = lyp_energy(ra, rb, saa, sab, sbb); break;
        case K_PBE_C: e = pbe_c_pol_energy(ra, rb, saa, sab, sbb); break;
        case K_SLATER_X:
        case K_VWN5_C: {  // LDA kernels inside a GGA functional (B3LYP): closed forms, d/d sigma = 0
          double el, va, vb;
          if (f.kern[k] == K_SLATER_X) f_slater_pol(ra.v, rb.v, el, va, vb);
          else f_vwn5_pol(ra.v, rb.v, el, va, vb);
          e = D5(el * (ra.v + rb.v));
          e.d[0] = va; e.d[1] = vb;
          break;
        }
        case K_PBE_X:
        case K_REVPBE_X: {  // exact spin scaling E_x[na, nb] = (E_x[2 na] + E_x[2 nb]) / 2
          const double kappa = f.kern[k] == K_PBE_X ? 0.8040 : 1.245;
          double ea, va, sa, eb, vb, sb;
          f_pbe_x(2. * ra.v, 4. * saa.v, ea, va, sa, kappa);
          f_pbe_x(2. * rb.v, 4. * sbb.v, eb, vb, sb, kappa);
          e = D5(ra.v * ea + rb.v * eb);
          e.d[0] = va; e.d[1] = vb; e.d[2] = 2. * sa; e.d[4] = 2. * sb;
          break;
        }
        default: e = D5(std::nan("")); break;  // not restated for UKS GGA
      }
      E = E + D5(f.coeff[k]) * e;
    }
    const double rho = rho2[2 * i] + rho2[2 * i + 1];
    if (rho <= 1e-24) {
      eps[i] = 0.;
      vrho2[2 * i] = vrho2[2 * i + 1] = 0.;
      vgamma3[3 * i] = vgamma3[3 * i + 1] = vgamma3[3 * i + 2] = 0.;
      continue;
    }
    eps[i] = E.v / rho;
    vrho2[2 * i] = E.d[0]; vrho2[2 * i + 1] = E.d[1];
    vgamma3[3 * i] = E.d[2]; vgamma3[3 * i + 1] = E.d[3]; vgamma3[3 * i + 2] = E.d[4];
  }
}

// ----------------------------------------------------------------------------------
// collocation: gau2grid semantics (gg_collocation / gg_collocation_deriv1), host layout
// basis_eval[mu + ipt*nbe]
// ----------------------------------------------------------------------------------
struct Basis {
  int nshells;
  const int32_t *l, *pure, *nprim;
  const double *alpha, *coeff, *origin;  // [nshells][32], [nshells][32], [nshells][3]
  int size(int s) const { return pure[s] ? 2 * l[s] + 1 : (l[s] + 1) * (l[s] + 2) / 2; }
};

// real solid harmonics of order l in CCA order m=-l..l as combinations of cartesian monomials
// (gau2grid "spherical CCA"; (l,0,0)-normalised cartesians)
struct SphTerm { int a, b, c; double f; };
// Built once through a thread-safe function-local static (C++11): the first caller may well be an
// OpenMP worker of oracle_exc_vxc.
std::vector<std::vector<std::vector<SphTerm>>> build_sph_table() {
  std::vector<std::vector<std::vector<SphTerm>>> T;
  {
    T.resize(5);
    const double s3 = std::sqrt(3.);
    T[0] = {{{0, 0, 0, 1.}}};
    T[1] = {{{0, 1, 0, 1.}}, {{0, 0, 1, 1.}}, {{1, 0, 0, 1.}}};
    T[2] = {{{1, 1, 0, s3}},
            {{0, 1, 1, s3}},
            {{0, 0, 2, 1.}, {2, 0, 0, -0.5}, {0, 2, 0, -0.5}},
            {{1, 0, 1, s3}},
            {{2, 0, 0, 0.5 * s3}, {0, 2, 0, -0.5 * s3}}};
    const double c58 = std::sqrt(5. / 8.), c15 = std::sqrt(15.), c38 = std::sqrt(3. / 8.);
    T[3] = {{{2, 1, 0, 3 * c58}, {0, 3, 0, -c58}},
            {{1, 1, 1, c15}},
            {{0, 1, 2, 4 * c38}, {2, 1, 0, -c38}, {0, 3, 0, -c38}},
            {{0, 0, 3, 1.}, {2, 0, 1, -1.5}, {0, 2, 1, -1.5}},
            {{1, 0, 2, 4 * c38}, {3, 0, 0, -c38}, {1, 2, 0, -c38}},
            {{2, 0, 1, 0.5 * c15}, {0, 2, 1, -0.5 * c15}},
            {{3, 0, 0, c58}, {1, 2, 0, -3 * c58}}};
    const double c35 = std::sqrt(35.), c70 = std::sqrt(70.), c5 = std::sqrt(5.), c10 = std::sqrt(10.);
    T[4] = {{{3, 1, 0, c35 / 2}, {1, 3, 0, -c35 / 2}},
            {{2, 1, 1, 3 * c70 / 4}, {0, 3, 1, -c70 / 4}},
            {{1, 1, 2, 6 * c5 / 2}, {3, 1, 0, -c5 / 2}, {1, 3, 0, -c5 / 2}},
            {{0, 1, 3, 4 * c10 / 4}, {2, 1, 1, -3 * c10 / 4}, {0, 3, 1, -3 * c10 / 4}},
            {{0, 0, 4, 1.}, {2, 0, 2, -3.}, {0, 2, 2, -3.}, {4, 0, 0, 0.375}, {2, 2, 0, 0.75}, {0, 4, 0, 0.375}},
            {{1, 0, 3, 4 * c10 / 4}, {3, 0, 1, -3 * c10 / 4}, {1, 2, 1, -3 * c10 / 4}},
            {{2, 0, 2, 6 * c5 / 4}, {0, 2, 2, -6 * c5 / 4}, {4, 0, 0, -c5 / 4}, {0, 4, 0, c5 / 4}},
            {{3, 0, 1, c70 / 4}, {1, 2, 1, -3 * c70 / 4}},
            {{4, 0, 0, c35 / 8}, {2, 2, 0, -6 * c35 / 8}, {0, 4, 0, c35 / 8}}};
  }
  return T;
}
const std::vector<std::vector<SphTerm>>& sph_table(int l) {
  static const std::vector<std::vector<std::vector<SphTerm>>> T = build_sph_table();
  return T[l];
}

double ipow(double x, int n) {
  double r = 1.;
  for (int i = 0; i < n; ++i) r *= x;
  return r;
}

// values (+ gradient) of one shell at one point, written with stride 1 at out[0..size)
void shell_at_point(const Basis& B, int s, const double* p, bool grad, double* v, double* gx, double* gy,
                    double* gz) {
  const double x = p[0] - B.origin[3 * s], y = p[1] - B.origin[3 * s + 1], z = p[2] - B.origin[3 * s + 2];
  const double r2 = x * x + y * y + z * z;
  double S0 = 0, S1 = 0;
  for (int k = 0; k < B.nprim[s]; ++k) {
    const double a = B.alpha[32 * s + k];
    const double e = B.coeff[32 * s + k] * std::exp(-a * r2);
    S0 += e;
    S1 += -2. * a * e;
  }
  const int l = B.l[s];
  auto mono = [&](int a, int b, int c, double& vv, double& dx, double& dy, double& dz) {
    const double f = ipow(x, a) * ipow(y, b) * ipow(z, c);
    vv = f * S0;
    if (grad) {
      dx = (a ? a * ipow(x, a - 1) * ipow(y, b) * ipow(z, c) * S0 : 0.) + f * x * S1;
      dy = (b ? b * ipow(x, a) * ipow(y, b - 1) * ipow(z, c) * S0 : 0.) + f * y * S1;
      dz = (c ? c * ipow(x, a) * ipow(y, b) * ipow(z, c - 1) * S0 : 0.) + f * z * S1;
    }
  };
  if (!B.pure[s]) {
    int c = 0;
    for (int a = l; a >= 0; --a)
      for (int b = l - a; b >= 0; --b, ++c) {
        double vv, dx = 0, dy = 0, dz = 0;
        mono(a, b, l - a - b, vv, dx, dy, dz);
        v[c] = vv;
        if (grad) { gx[c] = dx; gy[c] = dy; gz[c] = dz; }
      }
  } else {
    if (l > 4) { std::fprintf(stderr, "oracle: pure l>4 unsupported\n"); std::abort(); }
    const auto& T = sph_table(l);
    for (int m = 0; m < 2 * l + 1; ++m) {
      double vv = 0, dx = 0, dy = 0, dz = 0;
      for (auto& t : T[m]) {
        double v1, d1 = 0, d2 = 0, d3 = 0;
        mono(t.a, t.b, t.c, v1, d1, d2, d3);
        vv += t.f * v1; dx += t.f * d1; dy += t.f * d2; dz += t.f * d3;
      }
      v[m] = vv;
      if (grad) { gx[m] = dx; gy[m] = dy; gz[m] = dz; }
    }
  }
}

// The reference's own collocation backend: gau2grid, compiled from the reference tree into oracle/_ref/libgau2grid.so
// (oracle/Makefile).  When loaded (oracle_init_gau2grid -- the CPU-baseline / --impl reference legs of bench.py do),
// collocation() makes exactly the calls of gau2grid_collocation[_gradient]
// (local_work_driver/host/reference/gau2grid_collocation.cxx:25-116): one gg_collocation[_deriv1] per shell into a
// [component][point] scratch, then gg_fast_transpose into [point][component].  Otherwise the restatement above is used
// (tests/test_oracle_golden.py checks the two against each other to 1e-14).
typedef void (*gg_colloc_t)(int, unsigned long, const double*, unsigned long, int, const double*, const double*,
                            const double*, int, double*);
typedef void (*gg_deriv1_t)(int, unsigned long, const double*, unsigned long, int, const double*, const double*,
                            const double*, int, double*, double*, double*, double*);
typedef void (*gg_transpose_t)(unsigned long, unsigned long, const double*, double*);
gg_colloc_t f_gg_colloc = nullptr;
gg_deriv1_t f_gg_deriv1 = nullptr;
gg_transpose_t f_gg_transpose = nullptr;
bool use_gau2grid = false;

void collocation(const Basis& B, int nsh, const int32_t* shell_list, int npts, const double* pts, int nbe,
                 bool grad, double* ev, double* dx, double* dy, double* dz) {
  if (use_gau2grid && f_gg_colloc && npts > 0 && nbe > 0) {
    const size_t nn = (size_t)npts * nbe;
    static thread_local std::vector<double> rv;
    rv.resize((grad ? 4 : 1) * nn);
    double *r0 = rv.data(), *rx = r0 + nn, *ry = rx + nn, *rz = ry + nn;
    size_t ncomp = 0;
    for (int q = 0; q < nsh; ++q) {
      const int s = shell_list[q];
      const int order = B.pure[s] ? 300 : 400;  // GG_SPHERICAL_CCA / GG_CARTESIAN_CCA
      const size_t ioff = ncomp * npts;
      if (grad)
        f_gg_deriv1(B.l[s], (unsigned long)npts, pts, 3, B.nprim[s], B.coeff + 32 * s, B.alpha + 32 * s, B.origin + 3 * s,
                    order, r0 + ioff, rx + ioff, ry + ioff, rz + ioff);
      else
        f_gg_colloc(B.l[s], (unsigned long)npts, pts, 3, B.nprim[s], B.coeff + 32 * s, B.alpha + 32 * s, B.origin + 3 * s,
                    order, r0 + ioff);
      ncomp += B.size(s);
    }
    f_gg_transpose(ncomp, (unsigned long)npts, r0, ev);
    if (grad) {
      f_gg_transpose(ncomp, (unsigned long)npts, rx, dx);
      f_gg_transpose(ncomp, (unsigned long)npts, ry, dy);
      f_gg_transpose(ncomp, (unsigned long)npts, rz, dz);
    }
    return;
  }
  for (int i = 0; i < npts; ++i) {
    int off = 0;
    for (int q = 0; q < nsh; ++q) {
      const int s = shell_list[q];
      const size_t o = (size_t)i * nbe + off;
      shell_at_point(B, s, pts + 3 * i, grad, ev + o, grad ? dx + o : nullptr, grad ? dy + o : nullptr,
                     grad ? dz + o : nullptr);
      off += B.size(s);
    }
  }
}

// values, gradient and Hessian (xx,xy,xz,yy,yz,zz) of one shell at one point: the semantics of gau2grid
// gg_collocation_deriv2 as called by gau2grid_collocation_hessian
// (local_work_driver/host/reference/gau2grid_collocation.cxx:118-163).  phi = f(x,y,z) S(r^2):
//   d_i phi  = f_i S0 + f x_i S1
//   d_ij phi = f_ij S0 + (f_i x_j + f_j x_i + f delta_ij) S1 + f x_i x_j S2,
// S0 = sum c e, S1 = sum -2 a c e, S2 = sum 4 a^2 c e.
void shell_at_point_d2(const Basis& B, int s, const double* p, double* out[10]) {
  const double r[3] = {p[0] - B.origin[3 * s], p[1] - B.origin[3 * s + 1], p[2] - B.origin[3 * s + 2]};
  const double r2 = r[0] * r[0] + r[1] * r[1] + r[2] * r[2];
  double S0 = 0, S1 = 0, S2 = 0;
  for (int k = 0; k < B.nprim[s]; ++k) {
    const double a = B.alpha[32 * s + k];
    const double e = B.coeff[32 * s + k] * std::exp(-a * r2);
    S0 += e;
    S1 += -2. * a * e;
    S2 += 4. * a * a * e;
  }
  const int l = B.l[s];
  auto mono = [&](int a, int b, int c, double* o) {
    const int n[3] = {a, b, c};
    auto fpow = [&](int da, int db, int dc) {  // monomial with exponents lowered by (da, db, dc), times the falling factors
      const int e[3] = {n[0] - da, n[1] - db, n[2] - dc};
      if (e[0] < 0 || e[1] < 0 || e[2] < 0) return 0.;
      double pre = 1.;
      const int d[3] = {da, db, dc};
      for (int q = 0; q < 3; ++q)
        for (int k = 0; k < d[q]; ++k) pre *= double(n[q] - k);
      return pre * ipow(r[0], e[0]) * ipow(r[1], e[1]) * ipow(r[2], e[2]);
    };
    const double f = fpow(0, 0, 0);
    const double f1[3] = {fpow(1, 0, 0), fpow(0, 1, 0), fpow(0, 0, 1)};
    o[0] = f * S0;
    for (int i = 0; i < 3; ++i) o[1 + i] = f1[i] * S0 + f * r[i] * S1;
    int q = 4;
    for (int i = 0; i < 3; ++i)
      for (int j = i; j < 3; ++j, ++q) {
        int d[3] = {0, 0, 0};
        d[i] += 1; d[j] += 1;
        const double f2 = fpow(d[0], d[1], d[2]);
        o[q] = f2 * S0 + (f1[i] * r[j] + f1[j] * r[i] + (i == j ? f : 0.)) * S1 + f * r[i] * r[j] * S2;
      }
  };
  if (!B.pure[s]) {
    int c = 0;
    for (int a = l; a >= 0; --a)
      for (int b = l - a; b >= 0; --b, ++c) {
        double o[10];
        mono(a, b, l - a - b, o);
        for (int q = 0; q < 10; ++q) out[q][c] = o[q];
      }
  } else {
    if (l > 4) { std::fprintf(stderr, "oracle: pure l>4 unsupported\n"); std::abort(); }
    const auto& T = sph_table(l);
    for (int m = 0; m < 2 * l + 1; ++m) {
      double acc[10] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0};
      for (auto& t : T[m]) {
        double o[10];
        mono(t.a, t.b, t.c, o);
        for (int q = 0; q < 10; ++q) acc[q] += t.f * o[q];
      }
      for (int q = 0; q < 10; ++q) out[q][m] = acc[q];
    }
  }
}

// mats[q] + mu + ipt*nbe, q = value, x, y, z, xx, xy, xz, yy, yz, zz
void collocation_d2(const Basis& B, int nsh, const int32_t* shell_list, int npts, const double* pts, int nbe,
                    double* const mats[10]) {
  for (int i = 0; i < npts; ++i) {
    int off = 0;
    for (int q = 0; q < nsh; ++q) {
      const int s = shell_list[q];
      double* o[10];
      for (int k = 0; k < 10; ++k) o[k] = mats[k] + (size_t)i * nbe + off;
      shell_at_point_d2(B, s, pts + 3 * i, o);
      off += B.size(s);
    }
  }
}

// Neumaier sum
struct Acc {
  double s = 0, c = 0;
  void add(double x) {
    const double t = s + x;
    if (std::fabs(s) >= std::fabs(x)) c += (s - t) + x;
    else c += (x - t) + s;
    s = t;
  }
  double value() const { return s + c; }
};

}  // namespace

extern "C" {

// returns the BLAS in use
const char* oracle_init_blas(const char* path_hint) {
  if (f_dgemm) return blas_name;
  std::vector<std::string> cands;
  if (path_hint && *path_hint) cands.push_back(path_hint);
  if (const char* e = std::getenv("ORACLE_BLAS")) cands.push_back(e);
  for (auto& c : cands) {
    void* h = dlopen(c.c_str(), RTLD_NOW | RTLD_LOCAL);
    if (!h) continue;
    const char* names[][2] = {{"dgemm_", "dsyr2k_"}, {"scipy_dgemm_", "scipy_dsyr2k_"}};
    for (auto& n : names) {
      auto g = (dgemm_t)dlsym(h, n[0]);
      auto s = (dsyr2k_t)dlsym(h, n[1]);
      if (g && s) {
        f_dgemm = g;
        f_dsyr2k = s;
        std::snprintf(blas_name, sizeof(blas_name), "%s", c.c_str());
        // single-threaded BLAS inside the OpenMP task loop
        typedef void (*setnt_t)(int);
        for (const char* sn : {"openblas_set_num_threads", "scipy_openblas_set_num_threads"}) {
          if (auto f = (setnt_t)dlsym(h, sn)) { f(1); break; }
        }
        return blas_name;
      }
    }
  }
  return blas_name;
}

// loads oracle/_ref/libgau2grid.so (the reference's own collocation code); returns 1 when collocation() will use it
int oracle_init_gau2grid(const char* path, int enable) {
  if (!f_gg_colloc && path && *path) {
    if (void* h = dlopen(path, RTLD_NOW | RTLD_LOCAL)) {
      f_gg_colloc = (gg_colloc_t)dlsym(h, "gg_collocation");
      f_gg_deriv1 = (gg_deriv1_t)dlsym(h, "gg_collocation_deriv1");
      f_gg_transpose = (gg_transpose_t)dlsym(h, "gg_fast_transpose");
      if (!f_gg_colloc || !f_gg_deriv1 || !f_gg_transpose) f_gg_colloc = nullptr;
    }
  }
  use_gau2grid = enable && f_gg_colloc;
  return use_gau2grid ? 1 : 0;
}

int oracle_num_threads() {
#ifdef _OPENMP
  return omp_get_max_threads();
#else
  return 1;
#endif
}
void oracle_set_num_threads(int n) {
#ifdef _OPENMP
  omp_set_num_threads(n);
#endif
}

void oracle_functional(int nkern, const int* kern, const double* coeff, int is_gga, int npts,
                       const double* rho, const double* sigma, double* eps, double* vrho, double* vsigma) {
  Func f{};
  f.nkern = nkern; f.is_gga = is_gga;
  for (int k = 0; k < nkern; ++k) { f.kern[k] = kern[k]; f.coeff[k] = coeff[k]; }
  std::vector<double> zero;
  if (!sigma) { zero.assign(npts, 0.); sigma = zero.data(); }
  eval_func(f, npts, rho, sigma, eps, vrho, vsigma);
}

void oracle_collocation(int nshells_total, const int32_t* l, const int32_t* pure, const int32_t* nprim,
                        const double* alpha, const double* coeff, const double* origin, int nsh,
                        const int32_t* shell_list, int npts, const double* points, int want_grad,
                        double* eval, double* dx, double* dy, double* dz) {
  Basis B{nshells_total, l, pure, nprim, alpha, coeff, origin};
  int nbe = 0;
  for (int q = 0; q < nsh; ++q) nbe += B.size(shell_list[q]);
  collocation(B, nsh, shell_list, npts, points, nbe, want_grad != 0, eval, dx, dy, dz);
}

// weights.cxx:117-237 (reference_ssf_weights_host), RAB from src/molmeta.cxx:28-58
void oracle_ssf_weights(int natoms, const double* coords, int ntasks, const int32_t* task_npts,
                        const int32_t* task_iparent, const double* task_dist_nearest,
                        const double* points, double* weights) {
  const double magic = 0.64, tol = 1e-13;
  std::vector<double> RAB((size_t)natoms * natoms, 0.);
  for (int i = 0; i < natoms; ++i)
    for (int j = 0; j < i; ++j) {
      const double dx = coords[3 * i] - coords[3 * j], dy = coords[3 * i + 1] - coords[3 * j + 1],
                   dz = coords[3 * i + 2] - coords[3 * j + 2];
      RAB[i + (size_t)j * natoms] = std::sqrt(dx * dx + dy * dy + dz * dz);
      RAB[j + (size_t)i * natoms] = RAB[i + (size_t)j * natoms];
    }
  auto gFrisch = [&](double x) {
    const double s_x = x / magic;
    const double s_x2 = s_x * s_x, s_x3 = s_x * s_x2, s_x5 = s_x3 * s_x2, s_x7 = s_x5 * s_x2;
    return (35. * (s_x - s_x3) + 21. * s_x5 - 5. * s_x7) / 16.;
  };
  std::vector<size_t> off(ntasks + 1, 0);
  for (int t = 0; t < ntasks; ++t) off[t + 1] = off[t] + task_npts[t];
#pragma omp parallel
  {
    std::vector<double> part(natoms), dist(natoms);
#pragma omp for schedule(dynamic)
    for (int iT = 0; iT < ntasks; ++iT)
      for (int i = 0; i < task_npts[iT]; ++i) {
        const int par = task_iparent[iT];
        double& weight = weights[off[iT] + i];
        const double* point = points + 3 * (off[iT] + i);
        const double dist_cutoff = 0.5 * (1 - magic) * task_dist_nearest[iT];
        {
          const double dx = point[0] - coords[3 * par], dy = point[1] - coords[3 * par + 1],
                       dz = point[2] - coords[3 * par + 2];
          dist[par] = std::sqrt(dx * dx + dy * dy + dz * dz);
        }
        if (dist[par] < dist_cutoff) continue;
        for (int iA = 0; iA < natoms; ++iA) {
          if (iA == par) continue;
          const double dx = point[0] - coords[3 * iA], dy = point[1] - coords[3 * iA + 1],
                       dz = point[2] - coords[3 * iA + 2];
          dist[iA] = std::sqrt(dx * dx + dy * dy + dz * dz);
        }
        std::fill(part.begin(), part.end(), 1.);
        for (int iA = 0; iA < natoms; ++iA)
          for (int jA = 0; jA < iA; ++jA)
            if (part[iA] > tol || part[jA] > tol) {
              const double mu = (dist[iA] - dist[jA]) / RAB[jA + (size_t)iA * natoms];
              if (mu <= -magic) part[jA] = 0.;
              else if (mu >= magic) part[iA] = 0.;
              else {
                const double g = 0.5 * (1. - gFrisch(mu));
                part[iA] *= g;
                part[jA] *= 1. - g;
              }
            }
        double sum = 0.;
        for (int iA = 0; iA < natoms; ++iA) sum += part[iA];
        weight *= part[par] / sum;
      }
  }
}

// The EXC/VXC host driver (…host_integrator_exc_vxc.hpp:107-601) for RKS LDA/GGA.
// P: nbf x nbf (ld = ldp), the alpha density (X = 2 P B).  VXC: nbf x nbf (ld = nbf), fully
// overwritten and symmetric.  out3 = {EXC, N_EL, F_dense flops}.  task_stride > 1 evaluates
// only every task_stride-th task (bounded CPU-baseline sample in bench.py).
void oracle_exc_vxc(int nshells_total, const int32_t* l, const int32_t* pure, const int32_t* nprim,
                    const double* alpha, const double* coeff, const double* origin, int nbf,
                    const double* P, int ldp, int ntasks, const int32_t* task_npts,
                    const int32_t* task_nshells, const int32_t* shell_lists, const double* points,
                    const double* weights, int nkern, const int* kern, const double* kcoeff, int is_gga,
                    int task_stride, double* VXC, double* out3) {
  Basis B{nshells_total, l, pure, nprim, alpha, coeff, origin};
  Func func{};
  func.nkern = nkern; func.is_gga = is_gga;
  for (int k = 0; k < nkern; ++k) { func.kern[k] = kern[k]; func.coeff[k] = kcoeff[k]; }
  std::vector<int> first_ao(nshells_total + 1, 0);
  for (int s = 0; s < nshells_total; ++s) first_ao[s + 1] = first_ao[s] + B.size(s);
  std::vector<size_t> poff(ntasks + 1, 0), soff(ntasks + 1, 0);
  for (int t = 0; t < ntasks; ++t) {
    poff[t + 1] = poff[t] + task_npts[t];
    soff[t + 1] = soff[t] + task_nshells[t];
  }
  std::fill(VXC, VXC + (size_t)nbf * nbf, 0.);
  std::vector<double> exc_t(ntasks, 0.), nel_t(ntasks, 0.), flops_t(ntasks, 0.);
  if (task_stride < 1) task_stride = 1;

#pragma omp parallel
  {
    std::vector<double> ev, dxv, dyv, dzv, X, Z, scr, Psub, den, ddx, ddy, ddz, gam, eps, vrho, vgam;
    std::vector<int> ao;
#pragma omp for schedule(dynamic)
    for (int iT = 0; iT < ntasks; iT += task_stride) {
      const int npts = task_npts[iT];
      const int nsh = task_nshells[iT];
      const int32_t* sl = shell_lists + soff[iT];
      const double* pts = points + 3 * poff[iT];
      const double* w = weights + poff[iT];
      // local AO list == the cut map of gen_compressed_submat_map (contiguous shell ranges)
      ao.clear();
      for (int q = 0; q < nsh; ++q)
        for (int a = first_ao[sl[q]]; a < first_ao[sl[q] + 1]; ++a) ao.push_back(a);
      const int nbe = (int)ao.size();
      const size_t nn = (size_t)nbe * npts;
      ev.resize(nn); X.resize(nn); Z.resize(nn);
      if (is_gga) { dxv.resize(nn); dyv.resize(nn); dzv.resize(nn); }
      scr.resize((size_t)nbe * nbe); Psub.resize((size_t)nbe * nbe);
      den.resize(npts); eps.resize(npts); vrho.resize(npts); gam.assign(npts, 0.);
      if (is_gga) { ddx.resize(npts); ddy.resize(npts); ddz.resize(npts); vgam.resize(npts); }

      collocation(B, nsh, sl, npts, pts, nbe, is_gga != 0, ev.data(), dxv.data(), dyv.data(), dzv.data());

      // eval_xmat: submat_set + dgemm('N','N', nbe, npts, nbe, 2.0, P_sub, B)
      for (int j = 0; j < nbe; ++j)
        for (int i = 0; i < nbe; ++i) Psub[i + (size_t)j * nbe] = P[ao[i] + (size_t)ao[j] * ldp];
      gemm_nn(nbe, npts, nbe, 2.0, Psub.data(), nbe, ev.data(), nbe, X.data(), nbe);

      // eval_uvvar_{lda,gga}_rks
      for (int i = 0; i < npts; ++i) {
        const double* xi = X.data() + (size_t)i * nbe;
        const double* bi = ev.data() + (size_t)i * nbe;
        double d = 0;
        for (int m = 0; m < nbe; ++m) d += bi[m] * xi[m];
        den[i] = d;
        if (is_gga) {
          double a = 0, b = 0, c = 0;
          const double *bx = dxv.data() + (size_t)i * nbe, *by = dyv.data() + (size_t)i * nbe,
                       *bz = dzv.data() + (size_t)i * nbe;
          for (int m = 0; m < nbe; ++m) { a += bx[m] * xi[m]; b += by[m] * xi[m]; c += bz[m] * xi[m]; }
          ddx[i] = 2. * a; ddy[i] = 2. * b; ddz[i] = 2. * c;
          gam[i] = ddx[i] * ddx[i] + ddy[i] * ddy[i] + ddz[i] * ddz[i];
        }
      }
      eval_func(func, npts, den.data(), gam.data(), eps.data(), vrho.data(), is_gga ? vgam.data() : nullptr);
      // factor weights (:453-466) and scalar integrals (:490-497)
      Acc e_acc, n_acc;
      for (int i = 0; i < npts; ++i) {
        eps[i] *= w[i];
        vrho[i] *= w[i];
        if (is_gga) vgam[i] *= w[i];
        n_acc.add(w[i] * den[i]);
        e_acc.add(eps[i] * den[i]);
      }
      exc_t[iT] = e_acc.value();
      nel_t[iT] = n_acc.value();
      // eval_zmat_{lda,gga}_vxc_rks
      for (int i = 0; i < npts; ++i) {
        double* zi = Z.data() + (size_t)i * nbe;
        const double* bi = ev.data() + (size_t)i * nbe;
        const double lda_fact = 0.5 * vrho[i];
        for (int m = 0; m < nbe; ++m) zi[m] = lda_fact * bi[m];
        if (is_gga) {
          const double gf = 2. * vgam[i];
          const double xf = gf * ddx[i], yf = gf * ddy[i], zf = gf * ddz[i];
          const double *bx = dxv.data() + (size_t)i * nbe, *by = dyv.data() + (size_t)i * nbe,
                       *bz = dzv.data() + (size_t)i * nbe;
          for (int m = 0; m < nbe; ++m) zi[m] += xf * bx[m];
          for (int m = 0; m < nbe; ++m) zi[m] += yf * by[m];
          for (int m = 0; m < nbe; ++m) zi[m] += zf * bz[m];
        }
      }
      // inc_vxc: syr2k('L','N') + inc_by_submat_atomic (lower triangle is what matters)
      syr2k_ln(nbe, npts, ev.data(), nbe, Z.data(), nbe, scr.data(), nbe);
      for (int j = 0; j < nbe; ++j)
        for (int i = j; i < nbe; ++i) {
#pragma omp atomic
          VXC[ao[i] + (size_t)ao[j] * nbf] += scr[i + (size_t)j * nbe];
        }
      flops_t[iT] = 4. * double(nbe) * double(nbe) * double(npts);
    }
  }
  // symmetrise (:577-583): upper <- lower
  for (int j = 0; j < nbf; ++j)
    for (int i = j + 1; i < nbf; ++i) VXC[j + (size_t)i * nbf] = VXC[i + (size_t)j * nbf];
  Acc E, N, F;
  for (int t = 0; t < ntasks; ++t) { E.add(exc_t[t]); N.add(nel_t[t]); F.add(flops_t[t]); }
  out3[0] = E.value(); out3[1] = N.value(); out3[2] = F.value();
}


// UKS, LDA: reference_replicated_xc_host_integrator_exc_vxc.hpp:107-601 with is_uks (Ps = P_alpha +
// P_beta, Pz = P_alpha - P_beta): X_s = 1.0 Ps_sub B, X_z = 1.0 Pz_sub B (:387-396),
// eval_uvvar_lda_uks (driver.cxx:166-188), polarised functional, weights (:453-457), N_EL / EXC with the
// total density (:490-497), eval_zmat_lda_vxc_uks (driver.cxx:607-634), inc_vxc for VXCs and VXCz,
// symmetrise both.  out3 = {EXC, N_EL, dense flops}.
void oracle_exc_vxc_uks_lda(int nshells_total, const int32_t* l, const int32_t* pure, const int32_t* nprim,
                            const double* alpha, const double* coeff, const double* origin, int nbf,
                            const double* Ps, const double* Pz, int ldp, int ntasks, const int32_t* task_npts,
                            const int32_t* task_nshells, const int32_t* shell_lists, const double* points,
                            const double* weights, int nkern, const int* kern, const double* kcoeff,
                            double* VXCs, double* VXCz, double* out3) {
  Basis B{nshells_total, l, pure, nprim, alpha, coeff, origin};
  Func func{};
  func.nkern = nkern; func.is_gga = 0;
  for (int k = 0; k < nkern; ++k) { func.kern[k] = kern[k]; func.coeff[k] = kcoeff[k]; }
  std::vector<int> first_ao(nshells_total + 1, 0);
  for (int s = 0; s < nshells_total; ++s) first_ao[s + 1] = first_ao[s] + B.size(s);
  std::vector<size_t> poff(ntasks + 1, 0), soff(ntasks + 1, 0);
  for (int t = 0; t < ntasks; ++t) {
    poff[t + 1] = poff[t] + task_npts[t];
    soff[t + 1] = soff[t] + task_nshells[t];
  }
  std::fill(VXCs, VXCs + (size_t)nbf * nbf, 0.);
  std::fill(VXCz, VXCz + (size_t)nbf * nbf, 0.);
  std::vector<double> exc_t(ntasks, 0.), nel_t(ntasks, 0.), flops_t(ntasks, 0.);
  sph_table(0);  // built before the parallel region

#pragma omp parallel
  {
    std::vector<double> ev, Xs, Xz, Zs, Zz, scr, Psub, den2, eps, vrho2;
    std::vector<int> ao;
#pragma omp for schedule(dynamic)
    for (int iT = 0; iT < ntasks; ++iT) {
      const int npts = task_npts[iT];
      const int nsh = task_nshells[iT];
      const int32_t* sl = shell_lists + soff[iT];
      const double* pts = points + 3 * poff[iT];
      const double* w = weights + poff[iT];
      ao.clear();
      for (int q = 0; q < nsh; ++q)
        for (int a = first_ao[sl[q]]; a < first_ao[sl[q] + 1]; ++a) ao.push_back(a);
      const int nbe = (int)ao.size();
      const size_t nn = (size_t)nbe * npts;
      ev.resize(nn); Xs.resize(nn); Xz.resize(nn); Zs.resize(nn); Zz.resize(nn);
      scr.resize((size_t)nbe * nbe); Psub.resize((size_t)nbe * nbe);
      den2.resize(2 * (size_t)npts); eps.resize(npts); vrho2.resize(2 * (size_t)npts);
      collocation(B, nsh, sl, npts, pts, nbe, false, ev.data(), nullptr, nullptr, nullptr);
      for (int pass = 0; pass < 2; ++pass) {
        const double* P = pass == 0 ? Ps : Pz;
        for (int j = 0; j < nbe; ++j)
          for (int i = 0; i < nbe; ++i) Psub[i + (size_t)j * nbe] = P[ao[i] + (size_t)ao[j] * ldp];
        gemm_nn(nbe, npts, nbe, 1.0, Psub.data(), nbe, ev.data(), nbe, (pass == 0 ? Xs : Xz).data(), nbe);
      }
      for (int i = 0; i < npts; ++i) {
        const double* bi = ev.data() + (size_t)i * nbe;
        const double *xs = Xs.data() + (size_t)i * nbe, *xz = Xz.data() + (size_t)i * nbe;
        double rs = 0, rz = 0;
        for (int m = 0; m < nbe; ++m) { rs += bi[m] * xs[m]; rz += bi[m] * xz[m]; }
        den2[2 * i] = 0.5 * (rs + rz);
        den2[2 * i + 1] = 0.5 * (rs - rz);
      }
      eval_func_pol_lda(func, npts, den2.data(), eps.data(), vrho2.data());
      Acc e_acc, n_acc;
      for (int i = 0; i < npts; ++i) {
        eps[i] *= w[i];
        vrho2[2 * i] *= w[i];
        vrho2[2 * i + 1] *= w[i];
        const double den = den2[2 * i] + den2[2 * i + 1];
        n_acc.add(w[i] * den);
        e_acc.add(eps[i] * den);
      }
      exc_t[iT] = e_acc.value();
      nel_t[iT] = n_acc.value();
      for (int i = 0; i < npts; ++i) {
        const double* bi = ev.data() + (size_t)i * nbe;
        double *zs = Zs.data() + (size_t)i * nbe, *zz = Zz.data() + (size_t)i * nbe;
        const double factp = 0.5 * vrho2[2 * i], factm = 0.5 * vrho2[2 * i + 1];
        const double fs = 0.5 * (factp + factm), fz = 0.5 * (factp - factm);
        for (int m = 0; m < nbe; ++m) { zs[m] = fs * bi[m]; zz[m] = fz * bi[m]; }
      }
      for (int pass = 0; pass < 2; ++pass) {
        double* V = pass == 0 ? VXCs : VXCz;
        syr2k_ln(nbe, npts, ev.data(), nbe, (pass == 0 ? Zs : Zz).data(), nbe, scr.data(), nbe);
        for (int j = 0; j < nbe; ++j)
          for (int i = j; i < nbe; ++i) {
#pragma omp atomic
            V[ao[i] + (size_t)ao[j] * nbf] += scr[i + (size_t)j * nbe];
          }
      }
      flops_t[iT] = 8. * double(nbe) * double(nbe) * double(npts);
    }
  }
  for (double* V : {VXCs, VXCz})
    for (int j = 0; j < nbf; ++j)
      for (int i = j + 1; i < nbf; ++i) V[j + (size_t)i * nbf] = V[i + (size_t)j * nbf];
  Acc E, N, F;
  for (int t = 0; t < ntasks; ++t) { E.add(exc_t[t]); N.add(nel_t[t]); F.add(flops_t[t]); }
  out3[0] = E.value(); out3[1] = N.value(); out3[2] = F.value();
}

// UKS, GGA: as oracle_exc_vxc_uks_lda plus eval_uvvar_gga_uks (driver.cxx:270-328: grad n, grad Mz with the
// factor 2, gamma_{++,+-,--}), vgamma weights (:459-466) and eval_zmat_gga_vxc_uks (driver.cxx:715-773).
void oracle_exc_vxc_uks_gga(int nshells_total, const int32_t* l, const int32_t* pure, const int32_t* nprim,
                            const double* alpha, const double* coeff, const double* origin, int nbf,
                            const double* Ps, const double* Pz, int ldp, int ntasks, const int32_t* task_npts,
                            const int32_t* task_nshells, const int32_t* shell_lists, const double* points,
                            const double* weights, int nkern, const int* kern, const double* kcoeff,
                            double* VXCs, double* VXCz, double* out3) {
  Basis B{nshells_total, l, pure, nprim, alpha, coeff, origin};
  Func func{};
  func.nkern = nkern; func.is_gga = 1;
  for (int k = 0; k < nkern; ++k) { func.kern[k] = kern[k]; func.coeff[k] = kcoeff[k]; }
  std::vector<int> first_ao(nshells_total + 1, 0);
  for (int s = 0; s < nshells_total; ++s) first_ao[s + 1] = first_ao[s] + B.size(s);
  std::vector<size_t> poff(ntasks + 1, 0), soff(ntasks + 1, 0);
  for (int t = 0; t < ntasks; ++t) {
    poff[t + 1] = poff[t] + task_npts[t];
    soff[t + 1] = soff[t] + task_nshells[t];
  }
  std::fill(VXCs, VXCs + (size_t)nbf * nbf, 0.);
  std::fill(VXCz, VXCz + (size_t)nbf * nbf, 0.);
  std::vector<double> exc_t(ntasks, 0.), nel_t(ntasks, 0.), flops_t(ntasks, 0.);
  sph_table(0);

#pragma omp parallel
  {
    std::vector<double> ev, dxv, dyv, dzv, Xs, Xz, Zs, Zz, scr, Psub, den2, gam3, eps, vrho2, vgam3, dd[3];
    std::vector<int> ao;
#pragma omp for schedule(dynamic)
    for (int iT = 0; iT < ntasks; ++iT) {
      const int npts = task_npts[iT];
      const int nsh = task_nshells[iT];
      const int32_t* sl = shell_lists + soff[iT];
      const double* pts = points + 3 * poff[iT];
      const double* w = weights + poff[iT];
      ao.clear();
      for (int q = 0; q < nsh; ++q)
        for (int a = first_ao[sl[q]]; a < first_ao[sl[q] + 1]; ++a) ao.push_back(a);
      const int nbe = (int)ao.size();
      const size_t nn = (size_t)nbe * npts;
      ev.resize(nn); dxv.resize(nn); dyv.resize(nn); dzv.resize(nn);
      Xs.resize(nn); Xz.resize(nn); Zs.resize(nn); Zz.resize(nn);
      scr.resize((size_t)nbe * nbe); Psub.resize((size_t)nbe * nbe);
      den2.resize(2 * (size_t)npts); gam3.resize(3 * (size_t)npts); eps.resize(npts);
      vrho2.resize(2 * (size_t)npts); vgam3.resize(3 * (size_t)npts);
      for (auto& v : dd) v.resize(2 * (size_t)npts);
      collocation(B, nsh, sl, npts, pts, nbe, true, ev.data(), dxv.data(), dyv.data(), dzv.data());
      for (int pass = 0; pass < 2; ++pass) {
        const double* P = pass == 0 ? Ps : Pz;
        for (int j = 0; j < nbe; ++j)
          for (int i = 0; i < nbe; ++i) Psub[i + (size_t)j * nbe] = P[ao[i] + (size_t)ao[j] * ldp];
        gemm_nn(nbe, npts, nbe, 1.0, Psub.data(), nbe, ev.data(), nbe, (pass == 0 ? Xs : Xz).data(), nbe);
      }
      const double* dB[3] = {dxv.data(), dyv.data(), dzv.data()};
      for (int i = 0; i < npts; ++i) {
        const size_t o = (size_t)i * nbe;
        const double *bi = ev.data() + o, *xs = Xs.data() + o, *xz = Xz.data() + o;
        double rs = 0, rz = 0;
        for (int m = 0; m < nbe; ++m) { rs += bi[m] * xs[m]; rz += bi[m] * xz[m]; }
        den2[2 * i] = 0.5 * (rs + rz);
        den2[2 * i + 1] = 0.5 * (rs - rz);
        double dn[3], dm[3];
        for (int c = 0; c < 3; ++c) {
          double a = 0, b = 0;
          for (int m = 0; m < nbe; ++m) { a += dB[c][o + m] * xs[m]; b += dB[c][o + m] * xz[m]; }
          dn[c] = 2. * a; dm[c] = 2. * b;
          dd[c][2 * i] = dn[c]; dd[c][2 * i + 1] = dm[c];
        }
        const double dn_sq = dn[0] * dn[0] + dn[1] * dn[1] + dn[2] * dn[2];
        const double dm_sq = dm[0] * dm[0] + dm[1] * dm[1] + dm[2] * dm[2];
        const double dn_dm = dn[0] * dm[0] + dn[1] * dm[1] + dn[2] * dm[2];
        gam3[3 * i] = 0.25 * (dn_sq + dm_sq) + 0.5 * dn_dm;
        gam3[3 * i + 1] = 0.25 * (dn_sq - dm_sq);
        gam3[3 * i + 2] = 0.25 * (dn_sq + dm_sq) - 0.5 * dn_dm;
      }
      eval_func_pol_gga(func, npts, den2.data(), gam3.data(), eps.data(), vrho2.data(), vgam3.data());
      Acc e_acc, n_acc;
      for (int i = 0; i < npts; ++i) {
        eps[i] *= w[i];
        vrho2[2 * i] *= w[i]; vrho2[2 * i + 1] *= w[i];
        vgam3[3 * i] *= w[i]; vgam3[3 * i + 1] *= w[i]; vgam3[3 * i + 2] *= w[i];
        const double den = den2[2 * i] + den2[2 * i + 1];
        n_acc.add(w[i] * den);
        e_acc.add(eps[i] * den);
      }
      exc_t[iT] = e_acc.value();
      nel_t[iT] = n_acc.value();
      for (int i = 0; i < npts; ++i) {
        const size_t o = (size_t)i * nbe;
        const double* bi = ev.data() + o;
        double *zs = Zs.data() + o, *zz = Zz.data() + o;
        const double factp = 0.5 * vrho2[2 * i], factm = 0.5 * vrho2[2 * i + 1];
        const double fs = 0.5 * (factp + factm), fz = 0.5 * (factp - factm);
        const double gpp = vgam3[3 * i], gpm = vgam3[3 * i + 1], gmm = vgam3[3 * i + 2];
        const double g1 = 0.5 * (gpp + gpm + gmm), g2 = 0.5 * (gpp - gmm), g3 = 0.5 * (gpp - gpm + gmm);
        for (int m = 0; m < nbe; ++m) { zs[m] = fs * bi[m]; zz[m] = fz * bi[m]; }
        for (int c = 0; c < 3; ++c) {
          const double f_s = g1 * dd[c][2 * i] + g2 * dd[c][2 * i + 1];
          const double f_z = g3 * dd[c][2 * i + 1] + g2 * dd[c][2 * i];
          for (int m = 0; m < nbe; ++m) { zs[m] += f_s * dB[c][o + m]; zz[m] += f_z * dB[c][o + m]; }
        }
      }
      for (int pass = 0; pass < 2; ++pass) {
        double* V = pass == 0 ? VXCs : VXCz;
        syr2k_ln(nbe, npts, ev.data(), nbe, (pass == 0 ? Zs : Zz).data(), nbe, scr.data(), nbe);
        for (int j = 0; j < nbe; ++j)
          for (int i = j; i < nbe; ++i) {
#pragma omp atomic
            V[ao[i] + (size_t)ao[j] * nbf] += scr[i + (size_t)j * nbe];
          }
      }
      flops_t[iT] = 8. * double(nbe) * double(nbe) * double(npts);
    }
  }
  for (double* V : {VXCs, VXCz})
    for (int j = 0; j < nbf; ++j)
      for (int i = j + 1; i < nbf; ++i) V[j + (size_t)i * nbf] = V[i + (size_t)j * nbf];
  Acc E, N, F;
  for (int t = 0; t < ntasks; ++t) { E.add(exc_t[t]); N.add(nel_t[t]); F.add(flops_t[t]); }
  out3[0] = E.value(); out3[1] = N.value(); out3[2] = F.value();
}

void oracle_functional_pol_gga(int nkern, const int* kern, const double* coeff, int npts, const double* rho2,
                               const double* gamma3, double* eps, double* vrho2, double* vgamma3) {
  Func f{};
  f.nkern = nkern; f.is_gga = 1;
  for (int k = 0; k < nkern; ++k) { f.kern[k] = kern[k]; f.coeff[k] = coeff[k]; }
  eval_func_pol_gga(f, npts, rho2, gamma3, eps, vrho2, vgamma3);
}

// polarised LDA functional on its own (unit tests: spin-unpolarised limit, finite differences)
void oracle_functional_pol_lda(int nkern, const int* kern, const double* coeff, int npts, const double* rho2,
                               double* eps, double* vrho2) {
  Func f{};
  f.nkern = nkern; f.is_gga = 0;
  for (int k = 0; k < nkern; ++k) { f.kern[k] = kern[k]; f.coeff[k] = coeff[k]; }
  eval_func_pol_lda(f, npts, rho2, eps, vrho2);
}


void oracle_collocation_d2(int nshells_total, const int32_t* l, const int32_t* pure, const int32_t* nprim,
                           const double* alpha, const double* coeff, const double* origin, int nsh,
                           const int32_t* shell_list, int npts, const double* pts, double* out10) {
  Basis B{nshells_total, l, pure, nprim, alpha, coeff, origin};
  int nbe = 0;
  for (int q = 0; q < nsh; ++q) nbe += B.size(shell_list[q]);
  double* mats[10];
  for (int k = 0; k < 10; ++k) mats[k] = out10 + (size_t)k * nbe * npts;
  collocation_d2(B, nsh, shell_list, npts, pts, nbe, mats);
}

// reference_ssf_weights_1std_contraction_host (host/reference/weights.cxx:806-984): the SSF weight derivatives
// contracted with w_times_f = w * eps * rho per point, accumulated into exc_grad_w[3*natoms].
static void ssf_weights_1std_contraction(int natoms, const double* coords, const std::vector<double>& RAB,
                                         int iParent, double dist_nearest, int npts, const double* points,
                                         const double* w_times_f, double* exc_grad_w) {
  const double magic = 0.64, weight_tol = 1e-13;
  const double safe_magic_ssf_bound = magic - 1.e-4;
  const double w_times_f_thresh = 1.e-12;
  auto gFrisch = [&](double x) {
    const double s_x = x / magic;
    const double s_x2 = s_x * s_x, s_x3 = s_x * s_x2, s_x5 = s_x3 * s_x2, s_x7 = s_x5 * s_x2;
    return (35. * (s_x - s_x3) + 21. * s_x5 - 5. * s_x7) / 16.;
  };
  auto tFrisch = [&](double x) {
    const double s_x = x / magic;
    const double s_x2 = s_x * s_x, s_x3 = s_x * s_x2;
    const double numerator = 35. * (s_x3 + 3. * s_x2 + 3. * s_x + 1.);
    const double denominator = (x - magic) * (5. * s_x3 + 20. * s_x2 + 29. * s_x + 16.);
    return numerator / denominator;
  };
  std::vector<double> part(natoms), dist(natoms);
  auto X = [&](int a, int k) { return coords[3 * a + k]; };
  for (int i = 0; i < npts; ++i) {
    const double wf = w_times_f[i];
    if (std::fabs(wf) < w_times_f_thresh) continue;
    const double* point = points + 3 * i;
    const double dist_cutoff = 0.5 * (1 - magic) * dist_nearest;
    {
      const double dx = point[0] - X(iParent, 0), dy = point[1] - X(iParent, 1), dz = point[2] - X(iParent, 2);
      dist[iParent] = std::sqrt(dx * dx + dy * dy + dz * dz);
    }
    if (dist[iParent] < dist_cutoff) continue;
    for (int iA = 0; iA < natoms; ++iA) {
      if (iA == iParent) continue;
      const double dx = point[0] - X(iA, 0), dy = point[1] - X(iA, 1), dz = point[2] - X(iA, 2);
      dist[iA] = std::sqrt(dx * dx + dy * dy + dz * dz);
    }
    std::fill(part.begin(), part.end(), 1.);
    for (int iA = 0; iA < natoms; ++iA)
      for (int jA = 0; jA < iA; ++jA)
        if (part[iA] > weight_tol || part[jA] > weight_tol) {
          const double mu = (dist[iA] - dist[jA]) / RAB[jA + (size_t)iA * natoms];
          if (mu <= -magic) part[jA] = 0.;
          else if (mu >= magic) part[iA] = 0.;
          else {
            const double g = 0.5 * (1. - gFrisch(mu));
            part[iA] *= g;
            part[jA] *= 1. - g;
          }
        }
    double sum = 0.;
    for (int iA = 0; iA < natoms; ++iA) sum += part[iA];
    for (int iB = 0; iB < natoms; ++iB) {
      if (iB == iParent) continue;
      double gB[3] = {0., 0., 0.};
      const double rAB = RAB[iB + (size_t)iParent * natoms];
      const double rAB_inv = 1.0 / rAB;
      const double mu_AB = (dist[iParent] - dist[iB]) * rAB_inv;
      if (std::fabs(mu_AB) < safe_magic_ssf_bound) {
        const double coef1 = tFrisch(mu_AB) / rAB * (part[iParent] - sum) / sum * wf / dist[iB];
        for (int k = 0; k < 3; ++k) {
          const double uB = X(iB, k) - point[k], uBA = X(iB, k) - X(iParent, k);
          gB[k] = coef1 * (uB + mu_AB * uBA * rAB_inv * dist[iB]);
        }
      }
      if (part[iB] > weight_tol) {
        for (int iC = 0; iC < natoms; ++iC) {
          if (iB == iC) continue;
          const double rBC = RAB[iC + (size_t)iB * natoms];
          const double mu_BC = (dist[iB] - dist[iC]) / rBC;
          if (std::fabs(mu_BC) < safe_magic_ssf_bound) {
            const double t_BC = tFrisch(mu_BC);
            const double coef = part[iB] * t_BC / rBC / sum * wf;
            for (int k = 0; k < 3; ++k)
              gB[k] -= coef * ((X(iB, k) - point[k]) / dist[iB] - mu_BC * (X(iB, k) - X(iC, k)) / rBC);
            if (iC != iParent) {
              for (int k = 0; k < 3; ++k) {
                const double Ck = coef * ((X(iC, k) - point[k]) / dist[iC] + mu_BC * (X(iC, k) - X(iB, k)) / rBC);
#pragma omp atomic
                exc_grad_w[3 * iC + k] += Ck;
#pragma omp atomic
                exc_grad_w[3 * iParent + k] -= Ck;
              }
            }
          }
        }
      }
      for (int k = 0; k < 3; ++k) {
#pragma omp atomic
        exc_grad_w[3 * iB + k] += gB[k];
#pragma omp atomic
        exc_grad_w[3 * iParent + k] -= gB[k];  // translational invariance
      }
    }
  }
}

// The EXC gradient host driver for RKS LDA / GGA
// (reference_replicated_xc_host_integrator_exc_grad.hpp:107-601 with is_rks): collocation gradient / Hessian,
// X (and X_x, X_y, X_z for GGA) = 2 P_sub [B | dB], eval_uvvar_{lda,gga}_rks, functional, optional SSF weight
// derivatives (w_times_f = eps * rho * w), per-shell accumulation g_acc -> EXC_GRAD[shell centre] += -2 g_acc;
// with weight derivatives the shells on the task's parent atom are skipped and the parent receives +2 g_acc
// (translational invariance).  P is the alpha density (ld = ldp).  out: EXC_GRAD[3*natoms].
void oracle_exc_grad(int nshells_total, const int32_t* l, const int32_t* pure, const int32_t* nprim,
                     const double* alpha, const double* coeff, const double* origin, const int32_t* shell_to_center,
                     int natoms, const double* coords, int nbf, const double* P, int ldp, int ntasks,
                     const int32_t* task_npts, const int32_t* task_nshells, const int32_t* shell_lists,
                     const int32_t* task_iparent, const double* task_dist_nearest, const double* points,
                     const double* weights, int nkern, const int* kern, const double* kcoeff, int is_gga,
                     int include_weight_derivatives, double* EXC_GRAD) {
  Basis B{nshells_total, l, pure, nprim, alpha, coeff, origin};
  Func func{};
  func.nkern = nkern; func.is_gga = is_gga;
  for (int k = 0; k < nkern; ++k) { func.kern[k] = kern[k]; func.coeff[k] = kcoeff[k]; }
  std::vector<int> first_ao(nshells_total + 1, 0);
  for (int s = 0; s < nshells_total; ++s) first_ao[s + 1] = first_ao[s] + B.size(s);
  std::vector<size_t> poff(ntasks + 1, 0), soff(ntasks + 1, 0);
  for (int t = 0; t < ntasks; ++t) {
    poff[t + 1] = poff[t] + task_npts[t];
    soff[t + 1] = soff[t] + task_nshells[t];
  }
  std::vector<double> RAB((size_t)natoms * natoms, 0.);
  for (int i = 0; i < natoms; ++i)
    for (int j = 0; j < i; ++j) {
      const double dx = coords[3 * i] - coords[3 * j], dy = coords[3 * i + 1] - coords[3 * j + 1],
                   dz = coords[3 * i + 2] - coords[3 * j + 2];
      RAB[i + (size_t)j * natoms] = RAB[j + (size_t)i * natoms] = std::sqrt(dx * dx + dy * dy + dz * dz);
    }
  for (int i = 0; i < 3 * natoms; ++i) EXC_GRAD[i] = 0.;
  const int nmat = is_gga ? 10 : 4, nx = is_gga ? 4 : 1;

#pragma omp parallel
  {
    std::vector<double> be, X, Psub, den, ddx, ddy, ddz, gam, eps, vrho, vgam;
    std::vector<int> ao;
#pragma omp for schedule(dynamic)
    for (int iT = 0; iT < ntasks; ++iT) {
      const int npts = task_npts[iT];
      const int nsh = task_nshells[iT];
      const int32_t* sl = shell_lists + soff[iT];
      const double* pts = points + 3 * poff[iT];
      const double* w = weights + poff[iT];
      const int iParent = task_iparent[iT];
      ao.clear();
      for (int q = 0; q < nsh; ++q)
        for (int a = first_ao[sl[q]]; a < first_ao[sl[q] + 1]; ++a) ao.push_back(a);
      const int nbe = (int)ao.size();
      const size_t nn = (size_t)nbe * npts;
      be.assign(nmat * nn, 0.);
      X.resize(nx * nn);
      Psub.resize((size_t)nbe * nbe);
      den.resize(npts); eps.resize(npts); vrho.resize(npts); gam.assign(npts, 0.);
      ddx.assign(npts, 0.); ddy.assign(npts, 0.); ddz.assign(npts, 0.); vgam.assign(npts, 0.);
      double* mats[10];
      for (int k = 0; k < 10; ++k) mats[k] = k < nmat ? be.data() + k * nn : nullptr;
      if (is_gga) collocation_d2(B, nsh, sl, npts, pts, nbe, mats);
      else collocation(B, nsh, sl, npts, pts, nbe, true, mats[0], mats[1], mats[2], mats[3]);
      // eval_xmat over [B | dBx | dBy | dBz] (xmat_len * npts columns): X = 2 P_sub B
      for (int j = 0; j < nbe; ++j)
        for (int i = 0; i < nbe; ++i) Psub[i + (size_t)j * nbe] = P[ao[i] + (size_t)ao[j] * ldp];
      gemm_nn(nbe, nx * npts, nbe, 2.0, Psub.data(), nbe, be.data(), nbe, X.data(), nbe);
      const double *xN = X.data(), *xNx = X.data() + nn, *xNy = X.data() + 2 * nn, *xNz = X.data() + 3 * nn;
      // eval_uvvar_{lda,gga}_rks
      for (int i = 0; i < npts; ++i) {
        const double* xi = xN + (size_t)i * nbe;
        double d = 0;
        for (int m = 0; m < nbe; ++m) d += mats[0][(size_t)i * nbe + m] * xi[m];
        den[i] = d;
        if (is_gga) {
          double a = 0, b = 0, c = 0;
          for (int m = 0; m < nbe; ++m) {
            a += mats[1][(size_t)i * nbe + m] * xi[m];
            b += mats[2][(size_t)i * nbe + m] * xi[m];
            c += mats[3][(size_t)i * nbe + m] * xi[m];
          }
          ddx[i] = 2. * a; ddy[i] = 2. * b; ddz[i] = 2. * c;
          gam[i] = ddx[i] * ddx[i] + ddy[i] * ddy[i] + ddz[i] * ddz[i];
        }
      }
      eval_func(func, npts, den.data(), gam.data(), eps.data(), vrho.data(), is_gga ? vgam.data() : nullptr);
      if (include_weight_derivatives) {
        for (int i = 0; i < npts; ++i) eps[i] *= den[i] * w[i];
        ssf_weights_1std_contraction(natoms, coords, RAB, iParent, task_dist_nearest[iT], npts, pts, eps.data(),
                                     EXC_GRAD);
      }
      size_t bf_off = 0;
      for (int ish = 0; ish < nsh; ++ish) {
        const int sh_idx = sl[ish];
        const int sh_sz = B.size(sh_idx);
        const int iAt = shell_to_center[sh_idx];
        if (iAt == iParent && include_weight_derivatives) { bf_off += sh_sz; continue; }
        double g_acc_x = 0, g_acc_y = 0, g_acc_z = 0;
        for (int ibf = 0, mu = (int)bf_off; ibf < sh_sz; ++ibf, ++mu)
          for (int ipt = 0; ipt < npts; ++ipt) {
            const size_t mu_i = mu + (size_t)ipt * nbe;
            const double vrhop_ipt = w[ipt] * vrho[ipt];
            const double xn = xN[mu_i];
            const double dbx = mats[1][mu_i], dby = mats[2][mu_i], dbz = mats[3][mu_i];
            g_acc_x += vrhop_ipt * xn * dbx;
            g_acc_y += vrhop_ipt * xn * dby;
            g_acc_z += vrhop_ipt * xn * dbz;
            if (is_gga) {
              const double vgammapp_ipt = w[ipt] * vgam[ipt];
              const double ddenn_x = ddx[ipt], ddenn_y = ddy[ipt], ddenn_z = ddz[ipt];
              const double xnx = xNx[mu_i], xny = xNy[mu_i], xnz = xNz[mu_i];
              const double d2bxx = mats[4][mu_i], d2bxy = mats[5][mu_i], d2bxz = mats[6][mu_i],
                           d2byy = mats[7][mu_i], d2byz = mats[8][mu_i], d2bzz = mats[9][mu_i];
              const double d2_term_x = d2bxx * ddenn_x + d2bxy * ddenn_y + d2bxz * ddenn_z;
              const double d2_term_y = d2bxy * ddenn_x + d2byy * ddenn_y + d2byz * ddenn_z;
              const double d2_term_z = d2bxz * ddenn_x + d2byz * ddenn_y + d2bzz * ddenn_z;
              const double d11_xmat_term = ddenn_x * xnx + ddenn_y * xny + ddenn_z * xnz;
              g_acc_x += 2 * vgammapp_ipt * (xn * d2_term_x + dbx * d11_xmat_term);
              g_acc_y += 2 * vgammapp_ipt * (xn * d2_term_y + dby * d11_xmat_term);
              g_acc_z += 2 * vgammapp_ipt * (xn * d2_term_z + dbz * d11_xmat_term);
            }
          }
        const double g[3] = {g_acc_x, g_acc_y, g_acc_z};
        for (int k = 0; k < 3; ++k) {
#pragma omp atomic
          EXC_GRAD[3 * iAt + k] += -2 * g[k];
          if (include_weight_derivatives) {
#pragma omp atomic
            EXC_GRAD[3 * iParent + k] -= -2 * g[k];
          }
        }
        bf_off += sh_sz;
      }
    }
  }
}


// UKS EXC gradient (reference_replicated_xc_host_integrator_exc_grad.hpp:107-601 with is_uks): Ps = P_alpha + P_beta,
// Pz = P_alpha - P_beta, X factor 1.0 (:367), eval_uvvar_{lda,gga}_uks, polarised functional, w_times_f = eps (rho_+ +
// rho_-) w, the vrhon / vrhoz and four vgamma combinations of :424-436, 482-513.
void oracle_exc_grad_uks(int nshells_total, const int32_t* l, const int32_t* pure, const int32_t* nprim,
                         const double* alpha, const double* coeff, const double* origin, const int32_t* shell_to_center,
                         int natoms, const double* coords, int nbf, const double* Ps, const double* Pz, int ldp,
                         int ntasks, const int32_t* task_npts, const int32_t* task_nshells, const int32_t* shell_lists,
                         const int32_t* task_iparent, const double* task_dist_nearest, const double* points,
                         const double* weights, int nkern, const int* kern, const double* kcoeff, int is_gga,
                         int include_weight_derivatives, double* EXC_GRAD) {
  Basis B{nshells_total, l, pure, nprim, alpha, coeff, origin};
  Func func{};
  func.nkern = nkern; func.is_gga = is_gga;
  for (int k = 0; k < nkern; ++k) { func.kern[k] = kern[k]; func.coeff[k] = kcoeff[k]; }
  std::vector<int> first_ao(nshells_total + 1, 0);
  for (int s = 0; s < nshells_total; ++s) first_ao[s + 1] = first_ao[s] + B.size(s);
  std::vector<size_t> poff(ntasks + 1, 0), soff(ntasks + 1, 0);
  for (int t = 0; t < ntasks; ++t) {
    poff[t + 1] = poff[t] + task_npts[t];
    soff[t + 1] = soff[t] + task_nshells[t];
  }
  std::vector<double> RAB((size_t)natoms * natoms, 0.);
  for (int i = 0; i < natoms; ++i)
    for (int j = 0; j < i; ++j) {
      const double dx = coords[3 * i] - coords[3 * j], dy = coords[3 * i + 1] - coords[3 * j + 1],
                   dz = coords[3 * i + 2] - coords[3 * j + 2];
      RAB[i + (size_t)j * natoms] = RAB[j + (size_t)i * natoms] = std::sqrt(dx * dx + dy * dy + dz * dz);
    }
  for (int i = 0; i < 3 * natoms; ++i) EXC_GRAD[i] = 0.;
  const int nmat = is_gga ? 10 : 4, nx = is_gga ? 4 : 1;
  sph_table(0);

#pragma omp parallel
  {
    std::vector<double> be, XN, XZ, Psub, den2, gam3, eps, vrho2, vgam3, dd[3];
    std::vector<int> ao;
#pragma omp for schedule(dynamic)
    for (int iT = 0; iT < ntasks; ++iT) {
      const int npts = task_npts[iT];
      const int nsh = task_nshells[iT];
      const int32_t* sl = shell_lists + soff[iT];
      const double* pts = points + 3 * poff[iT];
      const double* w = weights + poff[iT];
      const int iParent = task_iparent[iT];
      ao.clear();
      for (int q = 0; q < nsh; ++q)
        for (int a = first_ao[sl[q]]; a < first_ao[sl[q] + 1]; ++a) ao.push_back(a);
      const int nbe = (int)ao.size();
      const size_t nn = (size_t)nbe * npts;
      be.assign(nmat * nn, 0.);
      XN.resize(nx * nn); XZ.resize(nx * nn);
      Psub.resize((size_t)nbe * nbe);
      den2.assign(2 * (size_t)npts, 0.); gam3.assign(3 * (size_t)npts, 0.); eps.resize(npts);
      vrho2.assign(2 * (size_t)npts, 0.); vgam3.assign(3 * (size_t)npts, 0.);
      for (auto& v : dd) v.assign(2 * (size_t)npts, 0.);
      double* mats[10];
      for (int k = 0; k < 10; ++k) mats[k] = k < nmat ? be.data() + k * nn : nullptr;
      if (is_gga) collocation_d2(B, nsh, sl, npts, pts, nbe, mats);
      else collocation(B, nsh, sl, npts, pts, nbe, true, mats[0], mats[1], mats[2], mats[3]);
      for (int pass = 0; pass < 2; ++pass) {
        const double* P = pass == 0 ? Ps : Pz;
        for (int j = 0; j < nbe; ++j)
          for (int i = 0; i < nbe; ++i) Psub[i + (size_t)j * nbe] = P[ao[i] + (size_t)ao[j] * ldp];
        gemm_nn(nbe, nx * npts, nbe, 1.0, Psub.data(), nbe, be.data(), nbe, (pass == 0 ? XN : XZ).data(), nbe);
      }
      // eval_uvvar_{lda,gga}_uks
      for (int i = 0; i < npts; ++i) {
        const size_t o = (size_t)i * nbe;
        double rs = 0, rz = 0;
        for (int m = 0; m < nbe; ++m) { rs += mats[0][o + m] * XN[o + m]; rz += mats[0][o + m] * XZ[o + m]; }
        den2[2 * i] = 0.5 * (rs + rz);
        den2[2 * i + 1] = 0.5 * (rs - rz);
        if (is_gga) {
          double dn[3], dm[3];
          for (int c = 0; c < 3; ++c) {
            double a = 0, b = 0;
            for (int m = 0; m < nbe; ++m) { a += mats[1 + c][o + m] * XN[o + m]; b += mats[1 + c][o + m] * XZ[o + m]; }
            dn[c] = 2. * a; dm[c] = 2. * b;
            dd[c][2 * i] = dn[c]; dd[c][2 * i + 1] = dm[c];
          }
          const double dn_sq = dn[0] * dn[0] + dn[1] * dn[1] + dn[2] * dn[2];
          const double dm_sq = dm[0] * dm[0] + dm[1] * dm[1] + dm[2] * dm[2];
          const double dn_dm = dn[0] * dm[0] + dn[1] * dm[1] + dn[2] * dm[2];
          gam3[3 * i] = 0.25 * (dn_sq + dm_sq) + 0.5 * dn_dm;
          gam3[3 * i + 1] = 0.25 * (dn_sq - dm_sq);
          gam3[3 * i + 2] = 0.25 * (dn_sq + dm_sq) - 0.5 * dn_dm;
        }
      }
      if (is_gga) eval_func_pol_gga(func, npts, den2.data(), gam3.data(), eps.data(), vrho2.data(), vgam3.data());
      else eval_func_pol_lda(func, npts, den2.data(), eps.data(), vrho2.data());
      if (include_weight_derivatives) {
        for (int i = 0; i < npts; ++i) eps[i] *= (den2[2 * i] + den2[2 * i + 1]) * w[i];
        ssf_weights_1std_contraction(natoms, coords, RAB, iParent, task_dist_nearest[iT], npts, pts, eps.data(),
                                     EXC_GRAD);
      }
      size_t bf_off = 0;
      for (int ish = 0; ish < nsh; ++ish) {
        const int sh_idx = sl[ish];
        const int sh_sz = B.size(sh_idx);
        const int iAt = shell_to_center[sh_idx];
        if (iAt == iParent && include_weight_derivatives) { bf_off += sh_sz; continue; }
        double g[3] = {0, 0, 0};
        for (int ibf = 0, mu = (int)bf_off; ibf < sh_sz; ++ibf, ++mu)
          for (int ipt = 0; ipt < npts; ++ipt) {
            const size_t mu_i = mu + (size_t)ipt * nbe;
            const double vrhop_ipt = w[ipt] * vrho2[2 * ipt], vrhom_ipt = w[ipt] * vrho2[2 * ipt + 1];
            const double xN = XN[mu_i], xZ = XZ[mu_i];
            const double db[3] = {mats[1][mu_i], mats[2][mu_i], mats[3][mu_i]};
            const double vrhon_ipt = vrhop_ipt + vrhom_ipt, vrhoz_ipt = vrhop_ipt - vrhom_ipt;
            for (int c = 0; c < 3; ++c) {
              g[c] += 0.5 * vrhon_ipt * xN * db[c];
              g[c] += 0.5 * vrhoz_ipt * xZ * db[c];
            }
            if (is_gga) {
              const double vpp = w[ipt] * vgam3[3 * ipt], vpm = w[ipt] * vgam3[3 * ipt + 1],
                           vmm = w[ipt] * vgam3[3 * ipt + 2];
              const double dn[3] = {dd[0][2 * ipt], dd[1][2 * ipt], dd[2][2 * ipt]};
              const double dz[3] = {dd[0][2 * ipt + 1], dd[1][2 * ipt + 1], dd[2][2 * ipt + 1]};
              const double xNg[3] = {XN[nn + mu_i], XN[2 * nn + mu_i], XN[3 * nn + mu_i]};
              const double xZg[3] = {XZ[nn + mu_i], XZ[2 * nn + mu_i], XZ[3 * nn + mu_i]};
              const double H[3][3] = {{mats[4][mu_i], mats[5][mu_i], mats[6][mu_i]},
                                      {mats[5][mu_i], mats[7][mu_i], mats[8][mu_i]},
                                      {mats[6][mu_i], mats[8][mu_i], mats[9][mu_i]}};
              const double d11nn = dn[0] * xNg[0] + dn[1] * xNg[1] + dn[2] * xNg[2];
              const double d11nz = dn[0] * xZg[0] + dn[1] * xZg[1] + dn[2] * xZg[2];
              const double d11zn = dz[0] * xNg[0] + dz[1] * xNg[1] + dz[2] * xNg[2];
              const double d11zz = dz[0] * xZg[0] + dz[1] * xZg[1] + dz[2] * xZg[2];
              for (int c = 0; c < 3; ++c) {
                const double d2n = H[c][0] * dn[0] + H[c][1] * dn[1] + H[c][2] * dn[2];
                const double d2z = H[c][0] * dz[0] + H[c][1] * dz[1] + H[c][2] * dz[2];
                g[c] += 0.5 * (vpp + vpm + vmm) * (d2n * xN + d11nn * db[c]);
                g[c] += 0.5 * (vpp - vmm) * (d2z * xN + d11zn * db[c]);
                g[c] += 0.5 * (vpp - vmm) * (d2n * xZ + d11nz * db[c]);
                g[c] += 0.5 * (vpp - vpm + vmm) * (d2z * xZ + d11zz * db[c]);
              }
            }
          }
        for (int k = 0; k < 3; ++k) {
#pragma omp atomic
          EXC_GRAD[3 * iAt + k] += -2 * g[k];
          if (include_weight_derivatives) {
#pragma omp atomic
            EXC_GRAD[3 * iParent + k] -= -2 * g[k];
          }
        }
        bf_off += sh_sz;
      }
    }
  }
}

}  // extern "C"
