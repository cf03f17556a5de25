// Shell normalisation / cutoff radii, BasisSetMap, MolMeta.
#include "types.hpp"
#include <algorithm>

namespace GauXC {

// (K-1)!! as in include/gauxc/shell.hpp:30-36 (only K = 2l, l <= 8 is used here)
static double df_Kminus1(int K) {
  double v = 1.;
  for (int k = K - 1; k > 1; k -= 2) v *= k;
  return v;
}

// include/gauxc/util/gau_rad_eval.hpp:20-30
static double gau_rad_eval(int l, int nprim, const double* alpha, const double* coeff, double r) {
  const double r2 = r * r;
  double tmp = 0.;
  for (int i = 0; i < nprim; ++i) tmp += coeff[i] * std::exp(-alpha[i] * r2);
  return std::pow(r, l) * tmp;
}

// include/gauxc/util/gau_rad_eval.hpp:32-70 -- walk in 0.01 bohr steps from the
// tightest-primitive estimate until |R_l(r)| crosses tol.
double gau_rad_cutoff(int l, int nprim, const double* alpha, const double* coeff, double tol) {
  if (tol <= 0.0) return std::numeric_limits<double>::infinity();
  const double log_tol = -std::log(tol);
  double r = 0;
  for (int i = 0; i < nprim; ++i) {
    const double log_alpha = std::log(alpha[i]);
    const double prim_cutoff = std::sqrt((log_tol + log_alpha / 2.) / alpha[i]);
    r = std::max(r, prim_cutoff);
  }
  std::vector<double> ac(coeff, coeff + nprim);
  for (auto& x : ac) x = std::abs(x);
  double v = gau_rad_eval(l, nprim, alpha, ac.data(), r);
  const double step = 0.01;
  if (v > tol) {
    while (v > tol) {
      r += step;
      v = gau_rad_eval(l, nprim, alpha, ac.data(), r);
    }
  } else {
    while (v < tol) {
      r -= step;
      v = gau_rad_eval(l, nprim, alpha, ac.data(), r);
    }
    r += step;
  }
  return r;
}

Shell::Shell(int nprim_, int l_, int pure_, const double* a, const double* c, const double* o,
             bool do_normalize)
    : nprim(nprim_), l(l_), pure(pure_) {
  if (nprim_ < 0 || nprim_ > shell_nprim_max) GAUXC_GENERIC_EXCEPTION("Invalid NPRIM");
  for (int i = 0; i < nprim_; ++i) {
    alpha[i] = a[i];
    coeff[i] = c[i];
  }
  for (int i = 0; i < 3; ++i) O[i] = o[i];
  if (do_normalize) normalize();
  compute_shell_cutoff();
}

// include/gauxc/shell.hpp:72-109 (Libint-style: unit-normalise the (l,0,0) cartesian)
void Shell::normalize() {
  constexpr double sqrt_Pi_cubed = 5.56832799683170784528481798212;
  const double two_to_l = std::pow(2, l);
  const double df_term = two_to_l / sqrt_Pi_cubed / df_Kminus1(2 * l);
  for (int i = 0; i < nprim; ++i) {
    if (alpha[i] != 0.) {
      const double two_alpha = 2 * alpha[i];
      const double two_alpha_to_am32 = std::pow(two_alpha, l + 1) * std::sqrt(two_alpha);
      coeff[i] *= std::sqrt(df_term * two_alpha_to_am32);
    }
  }
  double norm = 0;
  for (int i = 0; i < nprim; ++i)
    for (int j = 0; j <= i; ++j) {
      const double gamma = alpha[i] + alpha[j];
      const double gamma_to_am32 = std::pow(gamma, l + 1) * std::sqrt(gamma);
      norm += (i == j ? 1 : 2) * coeff[i] * coeff[j] / (df_term * gamma_to_am32);
    }
  const double f = 1. / std::sqrt(norm);
  for (int i = 0; i < nprim; ++i) coeff[i] *= f;
}

BasisSetMap::BasisSetMap(const BasisSet& basis, const Molecule& mol) {
  int32_t st = 0;
  for (auto& sh : basis) {
    shell_to_ao_range.push_back({st, st + sh.size()});
    st += sh.size();
    int32_t c = -1;
    for (size_t a = 0; a < mol.size(); ++a)
      if (mol[a].x == sh.O[0] && mol[a].y == sh.O[1] && mol[a].z == sh.O[2]) {
        c = (int32_t)a;
        break;
      }
    shell_to_center.push_back(c);
  }
  nbf = st;
}

// src/molmeta.cxx:28-58
MolMeta::MolMeta(const Molecule& mol) : natoms(mol.size()) {
  rab.assign(natoms * natoms, 0.);
  for (size_t i = 0; i < natoms; ++i)
    for (size_t j = 0; j < i; ++j) {
      const double dx = mol[i].x - mol[j].x, dy = mol[i].y - mol[j].y, dz = mol[i].z - mol[j].z;
      rab[i + j * natoms] = std::sqrt(dx * dx + dy * dy + dz * dz);
      rab[j + i * natoms] = rab[i + j * natoms];
    }
  dist_nearest.resize(natoms);
  for (size_t i = 0; i < natoms; ++i) {
    double dn = std::numeric_limits<double>::infinity();
    for (size_t j = 0; j < natoms; ++j)
      if (i != j && rab[i * natoms + j] < dn) dn = rab[i * natoms + j];
    dist_nearest[i] = dn;
  }
}

}  // namespace GauXC
