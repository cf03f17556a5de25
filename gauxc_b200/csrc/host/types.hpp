// Core data types of the B200 EXC/VXC path.  Names and field meaning mirror the
// reference's public headers so the host layer reads like GauXC's:
//   Atom/Molecule   include/gauxc/atom.hpp, molecule.hpp
//   Shell/BasisSet  include/gauxc/shell.hpp:49-200, basisset.hpp
//   BasisSetMap     include/gauxc/basisset_map.hpp
//   MolMeta         src/molmeta.cxx:28-58
//   XCTask          include/gauxc/xc_task.hpp:25-119
#pragma once
#include <array>
#include <cmath>
#include <cstdint>
#include <limits>
#include <sstream>
#include <stdexcept>
#include <string>
#include <vector>

namespace GauXC {

// ---- exceptions (include/gauxc/exceptions.hpp:40-97) ------------------------
class generic_gauxc_exception : public std::exception {
  std::string msg_;
public:
  generic_gauxc_exception(const std::string& file, const std::string& func, int line,
                          const std::string& msg) {
    std::ostringstream ss;
    ss << "Generic GauXC Exception (" << msg << ")\n  File     " << file << "\n  Function "
       << func << "\n  Line     " << line;
    msg_ = ss.str();
  }
  const char* what() const noexcept override { return msg_.c_str(); }
};
#define GAUXC_GENERIC_EXCEPTION(MSG) \
  throw ::GauXC::generic_gauxc_exception(__FILE__, __PRETTY_FUNCTION__, __LINE__, MSG)

// ---- enums (include/gauxc/enums.hpp) ---------------------------------------
enum class RadialQuad { Becke, MuraKnowles, MurrayHandyLaming, TreutlerAhlrichs };
enum class AtomicGridSizeDefault {
  FineGrid, UltraFineGrid, SuperFineGrid, GM3, GM5,
  PySCF0, PySCF1, PySCF2, PySCF3, PySCF4, PySCF5, PySCF6, PySCF7, PySCF8, PySCF9
};
enum class XCWeightAlg { NOTPARTITIONED, Becke, SSF, LKO };
enum class ExecutionSpace { Host, Device };
enum class PruningScheme { Unpruned, Robust, Treutler };

// ---- molecule --------------------------------------------------------------
struct Atom {
  int64_t Z;
  double x, y, z;
};
using Molecule = std::vector<Atom>;

inline int64_t molecule_max_Z(const Molecule& m) {
  int64_t z = 0;
  for (auto& a : m) z = std::max(z, a.Z);
  return z;
}

// ---- shells ----------------------------------------------------------------
constexpr int shell_nprim_max = 32;
constexpr double default_shell_tolerance = 1e-10;

double gau_rad_cutoff(int l, int nprim, const double* alpha, const double* coeff, double tol);

struct Shell {
  std::array<double, shell_nprim_max> alpha{};
  std::array<double, shell_nprim_max> coeff{};
  std::array<double, 3> O{};
  int32_t nprim = 0;
  int32_t l = 0;
  int32_t pure = 0;
  double cutoff_radius = 0.;
  double shell_tolerance = default_shell_tolerance;

  Shell() = default;
  Shell(int nprim_, int l_, int pure_, const double* a, const double* c, const double* o,
        bool do_normalize);

  int size() const { return pure ? 2 * l + 1 : (l + 1) * (l + 2) / 2; }
  void normalize();
  void compute_shell_cutoff() {
    cutoff_radius = gau_rad_cutoff(l, nprim, alpha.data(), coeff.data(), shell_tolerance);
  }
  void set_shell_tolerance(double tol) {
    if (tol != shell_tolerance) {
      shell_tolerance = tol;
      compute_shell_cutoff();
    }
  }
};

struct BasisSet : public std::vector<Shell> {
  int nshells() const { return (int)size(); }
  int nbf() const {
    int n = 0;
    for (auto& s : *this) n += s.size();
    return n;
  }
  int max_l() const {
    int l = 0;
    for (auto& s : *this) l = std::max(l, (int)s.l);
    return l;
  }
};

// shell -> AO range and shell -> centre (include/gauxc/basisset_map.hpp)
struct BasisSetMap {
  std::vector<std::pair<int32_t, int32_t>> shell_to_ao_range;  // [first, second)
  std::vector<int32_t> shell_to_center;
  int32_t nbf = 0;
  BasisSetMap() = default;
  BasisSetMap(const BasisSet& basis, const Molecule& mol);
};

// ---- MolMeta ---------------------------------------------------------------
struct MolMeta {
  size_t natoms = 0;
  std::vector<double> rab;           // natoms x natoms, zero diagonal
  std::vector<double> dist_nearest;  // natoms
  MolMeta() = default;
  explicit MolMeta(const Molecule& mol);
};

// ---- XCTask ----------------------------------------------------------------
struct XCTask {
  int32_t iParent = -1;
  std::vector<std::array<double, 3>> points;
  std::vector<double> weights;
  int32_t npts = 0;
  double dist_nearest = 0.;
  double max_weight = 1.;
  struct screening_data {
    std::vector<int32_t> shell_list;
    int32_t nbe = 0;
    bool equiv_with(const screening_data& o) const { return shell_list == o.shell_list; }
  } bfn_screening;

  void merge_with(const XCTask& o) {
    if (!equiv_with(o)) return;
    points.insert(points.end(), o.points.begin(), o.points.end());
    weights.insert(weights.end(), o.weights.begin(), o.weights.end());
    npts = (int32_t)points.size();
  }
  bool equiv_with(const XCTask& o) const {
    return iParent == o.iParent && bfn_screening.equiv_with(o.bfn_screening);
  }
  // include/gauxc/xc_task.hpp:109-114
  size_t cost(size_t n_deriv, size_t natoms) const {
    return (size_t(bfn_screening.nbe) * (1 + bfn_screening.nbe + n_deriv) + natoms * natoms) *
           size_t(npts);
  }
  size_t cost_exc_vxc(size_t n_deriv) const {
    return size_t(bfn_screening.nbe) * (1 + bfn_screening.nbe + n_deriv) * size_t(npts);
  }
};

}  // namespace GauXC
