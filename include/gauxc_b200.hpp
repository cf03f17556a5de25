// C++ facade of libgauxc_b200.so: the slice of GauXC's C++ API that the reference's own drivers use
// for the EXC/VXC path (tests/standalone_driver.cxx:31-477, tests/xc_integrator.cxx:150-260), header
// only, implemented over the C ABI of gauxc_b200.h -- so a C++ host code keeps its source:
//
//   Molecule mol; mol.emplace_back(AtomicNumber(8), x, y, z); ...
//   BasisSet<double> basis; basis.emplace_back(PrimSize(3), AngularMomentum(0), SphericalType(false),
//                                              alpha, coeff, origin); ...
//   auto mg = MolGridFactory::create_default_molgrid(mol, PruningScheme::Unpruned, BatchSize(512),
//                                                    RadialQuad::MuraKnowles,
//                                                    AtomicGridSizeDefault::UltraFineGrid);
//   auto rt = DeviceRuntimeEnvironment(0.9);
//   LoadBalancerFactory lb_factory(ExecutionSpace::Host, "Default");
//   auto lb = lb_factory.get_shared_instance(rt, mol, mg, basis);
//   MolecularWeightsFactory mw_factory(ExecutionSpace::Device, "Default", MolecularWeightsSettings{});
//   auto mw = mw_factory.get_instance();   mw.modify_weights(*lb);
//   functional_type func("PBE");
//   XCIntegratorFactory<matrix_type> integrator_factory(ExecutionSpace::Device, "Replicated",
//                                                        "Default", "Default", "Default");
//   auto integrator = integrator_factory.get_instance(func, lb);
//   auto [EXC, VXC] = integrator.eval_exc_vxc(P);
//
// Reference declarations mirrored here: include/gauxc/atom.hpp:20-49, molecule.hpp, shell.hpp:44-160,
// basisset.hpp, molgrid/defaults.hpp:78-92, runtime_environment/decl.hpp:26-86, load_balancer.hpp:
// 120-200, molecular_weights.hpp:24-115, xc_integrator.hpp:29-94, xc_integrator/integrator_factory.hpp:
// 26-104, exceptions.hpp.  Differences: functional_type is a name ("SVWN5", "PBE", ...) instead of an
// ExchCXX::XCFunctional (ExchCXX is not vendored), no MPI_Comm arguments (rank / size through
// set_comm), MatrixType needs (rows, cols) construction, rows(), cols(), data() and column-major
// storage exactly as in include/gauxc/xc_integrator/replicated/impl.hpp:108-120.
//
// Everything sits in an inline namespace so that these inline definitions can never interpose the
// library's internal GauXC:: classes of the same names.
#pragma once
#include "gauxc_b200.h"

#include <algorithm>
#include <array>
#include <cctype>
#include <cmath>
#include <cstdint>
#include <limits>
#include <memory>
#include <stdexcept>
#include <string>
#include <tuple>
#include <type_traits>
#include <utility>
#include <vector>

namespace GauXC {
inline namespace b200_capi {

// ---- exceptions (include/gauxc/exceptions.hpp) ----------------------------------------------------
class generic_gauxc_exception : public std::runtime_error {
public:
  explicit generic_gauxc_exception(const std::string& msg) : std::runtime_error(msg) {}
};

namespace detail {
// every C-ABI call goes through check(): status code 1 -> exception with the library's message
struct StatusGuard {
  GauXCStatus st{0, nullptr};
  ~StatusGuard() { gauxc_status_delete(&st); }
  void check() {
    if (st.code != 0) {
      const std::string msg = st.message ? st.message : "GauXC error";
      gauxc_status_delete(&st);
      throw generic_gauxc_exception(msg);
    }
  }
};
template <typename T, typename Tag>
class NamedType {
  T v_;

public:
  constexpr explicit NamedType(T v) : v_(v) {}
  constexpr NamedType() : v_() {}
  constexpr T get() const { return v_; }
  friend bool operator==(NamedType a, NamedType b) { return a.v_ == b.v_; }
};
}  // namespace detail

// ---- enums (include/gauxc/enums.hpp) ---------------------------------------------------------------
enum class RadialQuad { Becke, MuraKnowles, MurrayHandyLaming, TreutlerAhlrichs };
enum class AtomicGridSizeDefault { FineGrid, UltraFineGrid, SuperFineGrid, GM3, GM5 };
enum class XCWeightAlg { NOTPARTITIONED, Becke, SSF, LKO };
enum class ExecutionSpace { Host, Device };
enum class PruningScheme { Unpruned, Robust, Treutler };
using BatchSize = detail::NamedType<int64_t, struct BatchSizeType>;

// ---- atom / molecule --------------------------------------------------------------------------------
using AtomicNumber = detail::NamedType<int64_t, struct AtomicNumberType>;
struct Atom {
  AtomicNumber Z;
  double x = 0., y = 0., z = 0.;  // bohr
  Atom() = default;
  Atom(AtomicNumber _Z, double _x, double _y, double _z) : Z(_Z), x(_x), y(_y), z(_z) {}
};
class Molecule : public std::vector<Atom> {
public:
  using std::vector<Atom>::vector;
  size_t natoms() const { return size(); }
};

// ---- shell / basis set -----------------------------------------------------------------------------
using PrimSize = detail::NamedType<int32_t, struct PrimSizeType>;
using AngularMomentum = detail::NamedType<int32_t, struct AngularMomentumType>;
using SphericalType = detail::NamedType<int32_t, struct SphericalTypeType>;

template <typename F>
class Shell {
public:
  static constexpr size_t shell_nprim_max = 32;
  using prim_array = std::array<F, shell_nprim_max>;
  using cart_array = std::array<double, 3>;

private:
  prim_array alpha_{}, coeff_{};
  cart_array O_{};
  int32_t nprim_ = 0, l_ = 0, pure_ = 0;
  double tol_ = 1e-10;
  bool normalize_ = true;

public:
  Shell() = default;
  Shell(PrimSize nprim, AngularMomentum l, SphericalType pure, prim_array alpha, prim_array coeff,
        cart_array O, bool _normalize = true)
      : alpha_(alpha), coeff_(coeff), O_(O), nprim_(nprim.get()), l_(l.get()), pure_(pure.get()),
        normalize_(_normalize) {}
  void set_shell_tolerance(double tol) { tol_ = tol; }
  int32_t nprim() const { return nprim_; }
  int32_t l() const { return l_; }
  int32_t pure() const { return pure_; }
  int32_t size() const { return pure_ ? 2 * l_ + 1 : (l_ + 1) * (l_ + 2) / 2; }
  const prim_array& alpha() const { return alpha_; }
  const prim_array& coeff() const { return coeff_; }
  const cart_array& O() const { return O_; }
  double shell_tolerance() const { return tol_; }
  bool normalize_requested() const { return normalize_; }
};

template <typename F>
class BasisSet : public std::vector<Shell<F>> {
public:
  using std::vector<Shell<F>>::vector;
  int32_t nshells() const { return (int32_t)this->size(); }
  int32_t nbf() const {
    int32_t n = 0;
    for (const auto& s : *this) n += s.size();
    return n;
  }
};

// ---- handles over the C objects ------------------------------------------------------------------------
namespace detail {
inline GauXCMolecule to_c(const Molecule& mol) {
  std::vector<GauXCAtom> a(mol.size());
  for (size_t i = 0; i < mol.size(); ++i) a[i] = GauXCAtom{mol[i].Z.get(), mol[i].x, mol[i].y, mol[i].z};
  StatusGuard g;
  GauXCMolecule h = gauxc_molecule_new_from_atoms(&g.st, a.data(), a.size());
  g.check();
  return h;
}
template <typename F>
inline GauXCBasisSet to_c(const BasisSet<F>& basis) {
  std::vector<GauXCShell> sh(basis.size());
  bool normalize = true;
  for (size_t i = 0; i < basis.size(); ++i) {
    const auto& s = basis[i];
    GauXCShell c{};
    c.l = s.l();
    c.pure = s.pure() != 0;
    c.nprim = s.nprim();
    for (int k = 0; k < s.nprim(); ++k) {
      c.exponents[k] = (double)s.alpha()[k];
      c.coefficients[k] = (double)s.coeff()[k];
    }
    for (int k = 0; k < 3; ++k) c.origin[k] = s.O()[k];
    c.shell_tolerance = s.shell_tolerance();
    normalize = normalize && s.normalize_requested();
    sh[i] = c;
  }
  StatusGuard g;
  GauXCBasisSet h = gauxc_basisset_new_from_shells(&g.st, sh.data(), sh.size(), normalize);
  g.check();
  return h;
}
}  // namespace detail

class MolGrid {
  std::shared_ptr<GauXCMolGrid> h_;
  std::shared_ptr<GauXCMolecule> mol_;

public:
  MolGrid(const Molecule& mol, PruningScheme scheme, BatchSize bsz, RadialQuad rq, AtomicGridSizeDefault gs) {
    mol_ = std::shared_ptr<GauXCMolecule>(new GauXCMolecule(detail::to_c(mol)), [](GauXCMolecule* p) {
      GauXCStatus st{0, nullptr};
      gauxc_molecule_delete(&st, p);
      gauxc_status_delete(&st);
      delete p;
    });
    detail::StatusGuard g;
    GauXCMolGrid h = gauxc_molgrid_new_default(&g.st, *mol_, (GauXC_PruningScheme)(int)scheme, bsz.get(),
                                               (GauXC_RadialQuad)(int)rq, (GauXC_AtomicGridSizeDefault)(int)gs);
    g.check();
    h_ = std::shared_ptr<GauXCMolGrid>(new GauXCMolGrid(h), [](GauXCMolGrid* p) {
      GauXCStatus st{0, nullptr};
      gauxc_molgrid_delete(&st, p);
      gauxc_status_delete(&st);
      delete p;
    });
  }
  const GauXCMolGrid& c_handle() const { return *h_; }
};
struct MolGridFactory {
  static MolGrid create_default_molgrid(const Molecule& mol, PruningScheme scheme, BatchSize bsz, RadialQuad rq,
                                        AtomicGridSizeDefault gs) {
    return MolGrid(mol, scheme, bsz, rq, gs);
  }
};

// ---- runtime environment (copies share the underlying object, decl.hpp:26-86) ------------------------
class RuntimeEnvironment {
protected:
  std::shared_ptr<GauXCRuntimeEnvironment> h_;
  static std::shared_ptr<GauXCRuntimeEnvironment> adopt(GauXCRuntimeEnvironment h) {
    return std::shared_ptr<GauXCRuntimeEnvironment>(new GauXCRuntimeEnvironment(h), [](GauXCRuntimeEnvironment* p) {
      GauXCStatus st{0, nullptr};
      gauxc_runtime_environment_delete(&st, p);
      gauxc_status_delete(&st);
      delete p;
    });
  }
  struct no_init {};
  explicit RuntimeEnvironment(no_init) {}

public:
  RuntimeEnvironment() {
    detail::StatusGuard g;
    GauXCRuntimeEnvironment h = gauxc_runtime_environment_new(&g.st);
    g.check();
    h_ = adopt(h);
  }
  int comm_rank() const {
    detail::StatusGuard g;
    const int r = gauxc_runtime_environment_comm_rank(&g.st, *h_);
    g.check();
    return r;
  }
  int comm_size() const {
    detail::StatusGuard g;
    const int r = gauxc_runtime_environment_comm_size(&g.st, *h_);
    g.check();
    return r;
  }
  // stands in for the MPI_Comm argument of the reference constructors (one process per GPU)
  void set_comm(int rank, int size) {
    detail::StatusGuard g;
    gauxc_b200_runtime_environment_set_comm(&g.st, *h_, rank, size);
    g.check();
  }
  const GauXCRuntimeEnvironment& c_handle() const { return *h_; }
};
class DeviceRuntimeEnvironment : public RuntimeEnvironment {
public:
  explicit DeviceRuntimeEnvironment(double fill_fraction) : RuntimeEnvironment(no_init{}) {
    detail::StatusGuard g;
    GauXCRuntimeEnvironment h = gauxc_device_runtime_environment_new(&g.st, fill_fraction);
    g.check();
    h_ = adopt(h);
  }
  DeviceRuntimeEnvironment(void* mem, size_t mem_sz) : RuntimeEnvironment(no_init{}) {
    detail::StatusGuard g;
    GauXCRuntimeEnvironment h = gauxc_device_runtime_environment_new_mem(&g.st, mem, mem_sz);
    g.check();
    h_ = adopt(h);
  }
};

// ---- tasks, molecular meta data (include/gauxc/xc_task.hpp:25-62, include/gauxc/molmeta.hpp) ------------------
struct XCTask {
  struct screening_data {
    std::vector<int32_t> shell_list;
    int32_t nbe = 0;
  };
  int32_t iParent = -1;
  std::vector<std::array<double, 3>> points;
  std::vector<double> weights;
  int32_t npts = 0;
  double dist_nearest = 0.;
  screening_data bfn_screening;
};

class MolMeta {
  size_t natoms_ = 0;
  std::vector<double> rab_, dist_nearest_;

public:
  MolMeta() = default;
  explicit MolMeta(const std::vector<Atom>& mol) : natoms_(mol.size()), rab_(mol.size() * mol.size(), 0.),
                                                   dist_nearest_(mol.size(), 0.) {
    // src/molmeta.cxx:28-58
    for (size_t i = 0; i < natoms_; ++i) {
      double best = std::numeric_limits<double>::infinity();
      for (size_t j = 0; j < natoms_; ++j) {
        const double dx = mol[i].x - mol[j].x, dy = mol[i].y - mol[j].y, dz = mol[i].z - mol[j].z;
        const double r = std::sqrt(dx * dx + dy * dy + dz * dz);
        rab_[i + j * natoms_] = r;
        if (i != j) best = std::min(best, r);
      }
      dist_nearest_[i] = best;
    }
  }
  size_t natoms() const { return natoms_; }
  const std::vector<double>& rab() const { return rab_; }
  const std::vector<double>& dist_nearest() const { return dist_nearest_; }
};

struct LoadBalancerState {
  bool modified_weights_are_stored = false;
  XCWeightAlg weight_alg = XCWeightAlg::NOTPARTITIONED;
};

// ---- load balancer -----------------------------------------------------------------------------------
class LoadBalancer {
  std::shared_ptr<GauXCLoadBalancer> h_;
  // the C objects the balancer was built from must outlive it
  std::shared_ptr<GauXCMolecule> mol_;
  std::shared_ptr<GauXCBasisSet> basis_;
  MolGrid mg_;
  RuntimeEnvironment rt_;
  MolMeta meta_;
  friend class LoadBalancerFactory;
  LoadBalancer(const MolGrid& mg, const RuntimeEnvironment& rt) : mg_(mg), rt_(rt) {}

public:
  size_t total_npts() const {
    detail::StatusGuard g;
    const int64_t n = gauxc_b200_load_balancer_total_npts(&g.st, *h_);
    g.check();
    return (size_t)n;
  }
  size_t ntasks() const {
    detail::StatusGuard g;
    const int64_t n = gauxc_b200_load_balancer_ntasks(&g.st, *h_);
    g.check();
    return (size_t)n;
  }
  const RuntimeEnvironment& runtime() const { return rt_; }
  const GauXCLoadBalancer& c_handle() const { return *h_; }

  // include/gauxc/load_balancer.hpp:71-119.  The task list lives behind the C ABI (and, once the weights
  // are modified, on the device): get_tasks() materialises a copy.
  std::vector<XCTask> get_tasks() const {
    const size_t nt = ntasks();
    std::vector<int32_t> ip(nt), np(nt), nbe(nt), nsh(nt);
    std::vector<double> dn(nt);
    detail::StatusGuard g;
    gauxc_b200_load_balancer_task_info(&g.st, *h_, ip.data(), np.data(), nbe.data(), nsh.data(), dn.data());
    g.check();
    std::vector<XCTask> tasks(nt);
    for (size_t t = 0; t < nt; ++t) {
      auto& x = tasks[t];
      x.iParent = ip[t]; x.npts = np[t]; x.dist_nearest = dn[t];
      x.points.resize(np[t]); x.weights.resize(np[t]);
      x.bfn_screening.nbe = nbe[t]; x.bfn_screening.shell_list.resize(nsh[t]);
      gauxc_b200_load_balancer_get_task(&g.st, *h_, (int64_t)t, x.points.empty() ? nullptr : x.points[0].data(),
                                        x.weights.data(), x.bfn_screening.shell_list.data());
      g.check();
    }
    return tasks;
  }
  LoadBalancerState state() const {
    int mod = 0, alg = 0;
    detail::StatusGuard g;
    gauxc_b200_load_balancer_state(&g.st, *h_, &mod, &alg);
    g.check();
    return LoadBalancerState{mod != 0, (XCWeightAlg)alg};
  }
  size_t max_npts() const {
    size_t m = 0;
    for (auto& t : get_tasks()) m = std::max(m, (size_t)t.npts);
    return m;
  }
  size_t max_nbe() const {
    size_t m = 0;
    for (auto& t : get_tasks()) m = std::max(m, (size_t)t.bfn_screening.nbe);
    return m;
  }
  const MolGrid& molgrid() const { return mg_; }
  const MolMeta& molmeta() const { return meta_; }
};

class LoadBalancerFactory {
  ExecutionSpace ex_;
  std::string kernel_;

public:
  LoadBalancerFactory() = delete;
  LoadBalancerFactory(ExecutionSpace ex, std::string kernel_name) : ex_(ex), kernel_(std::move(kernel_name)) {}
  template <typename F>
  std::shared_ptr<LoadBalancer> get_shared_instance(const RuntimeEnvironment& rt, const Molecule& mol,
                                                    const MolGrid& mg, const BasisSet<F>& basis) {
    auto del_status = [](GauXCStatus& st) { gauxc_status_delete(&st); };
    (void)del_status;
    std::shared_ptr<LoadBalancer> lb(new LoadBalancer(mg, rt));
    lb->meta_ = MolMeta(mol);
    lb->mol_ = std::shared_ptr<GauXCMolecule>(new GauXCMolecule(detail::to_c(mol)), [](GauXCMolecule* p) {
      GauXCStatus st{0, nullptr};
      gauxc_molecule_delete(&st, p);
      gauxc_status_delete(&st);
      delete p;
    });
    lb->basis_ = std::shared_ptr<GauXCBasisSet>(new GauXCBasisSet(detail::to_c(basis)), [](GauXCBasisSet* p) {
      GauXCStatus st{0, nullptr};
      gauxc_basisset_delete(&st, p);
      gauxc_status_delete(&st);
      delete p;
    });
    detail::StatusGuard g;
    GauXCLoadBalancerFactory f = gauxc_load_balancer_factory_new(&g.st, (GauXC_ExecutionSpace)(int)ex_, kernel_.c_str());
    g.check();
    GauXCLoadBalancer h =
        gauxc_load_balancer_factory_get_instance(&g.st, f, rt.c_handle(), *lb->mol_, mg.c_handle(), *lb->basis_);
    GauXCStatus st2{0, nullptr};
    gauxc_load_balancer_factory_delete(&st2, &f);
    gauxc_status_delete(&st2);
    g.check();
    lb->h_ = std::shared_ptr<GauXCLoadBalancer>(new GauXCLoadBalancer(h), [](GauXCLoadBalancer* p) {
      GauXCStatus st{0, nullptr};
      gauxc_load_balancer_delete(&st, p);
      gauxc_status_delete(&st);
      delete p;
    });
    return lb;
  }
  template <typename... Args>
  LoadBalancer get_instance(Args&&... args) {
    return *get_shared_instance(std::forward<Args>(args)...);
  }
};

// ---- molecular weights --------------------------------------------------------------------------------
struct MolecularWeightsSettings {
  XCWeightAlg weight_alg = XCWeightAlg::SSF;
  bool becke_size_adjustment = false;
};
class MolecularWeights {
  std::shared_ptr<GauXCMolecularWeights> h_;
  friend class MolecularWeightsFactory;

public:
  void modify_weights(LoadBalancer& lb) const {
    detail::StatusGuard g;
    gauxc_molecular_weights_modify_weights(&g.st, *h_, lb.c_handle());
    g.check();
  }
  double last_ms() const {  // device time of the SSF kernel (CUDA events)
    detail::StatusGuard g;
    const double ms = gauxc_b200_molecular_weights_last_ms(&g.st, *h_);
    g.check();
    return ms;
  }
};
class MolecularWeightsFactory {
  ExecutionSpace ex_;
  std::string lwd_;
  MolecularWeightsSettings settings_;

public:
  MolecularWeightsFactory() = delete;
  MolecularWeightsFactory(ExecutionSpace ex, std::string local_work_kernel_name, MolecularWeightsSettings s)
      : ex_(ex), lwd_(std::move(local_work_kernel_name)), settings_(s) {}
  MolecularWeights get_instance() {
    detail::StatusGuard g;
    GauXCMolecularWeightsSettings cs{(GauXC_XCWeightAlg)(int)settings_.weight_alg, settings_.becke_size_adjustment};
    GauXCMolecularWeightsFactory f =
        gauxc_molecular_weights_factory_new(&g.st, (GauXC_ExecutionSpace)(int)ex_, lwd_.c_str(), cs);
    g.check();
    GauXCMolecularWeights h = gauxc_molecular_weights_factory_get_instance(&g.st, f);
    GauXCStatus st2{0, nullptr};
    gauxc_molecular_weights_factory_delete(&st2, &f);
    gauxc_status_delete(&st2);
    g.check();
    MolecularWeights mw;
    mw.h_ = std::shared_ptr<GauXCMolecularWeights>(new GauXCMolecularWeights(h), [](GauXCMolecularWeights* p) {
      GauXCStatus st{0, nullptr};
      gauxc_molecular_weights_delete(&st, p);
      gauxc_status_delete(&st);
      delete p;
    });
    return mw;
  }
  std::shared_ptr<MolecularWeights> get_shared_instance() { return std::make_shared<MolecularWeights>(get_instance()); }
};

// ---- reduction driver (include/gauxc/reduction_driver.hpp:26-70) ------------------------------------------
enum class ReductionOp { Sum };
// "Default" / "NCCL": in-place sum over the ranks of the job of a DEVICE buffer through the library's NCCL
// communicator (gauxc_b200_nccl_init); a single rank reduces to a no-op.  "BasicMPI" does not exist here.
class ReductionDriver {
  int size_ = 1;

public:
  explicit ReductionDriver(int comm_size) : size_(comm_size) {}
  bool takes_host_memory() const { return false; }
  bool takes_device_memory() const { return true; }
  int comm_size() const { return size_; }
  // queue (the reference's std::any holding a stream) is ignored: the call returns when the sum is complete
  template <typename T, typename... Queue>
  void allreduce_inplace(T* data, size_t n, ReductionOp, Queue&&...) {
    static_assert(std::is_same<T, double>::value, "FP64 buffers only");
    detail::StatusGuard g;
    gauxc_b200_allreduce_device(&g.st, data, n);
    g.check();
  }
};
struct ReductionDriverFactory {
  static std::shared_ptr<ReductionDriver> get_shared_instance(const RuntimeEnvironment& rt,
                                                              std::string kernel_name = "Default") {
    for (auto& c : kernel_name) c = (char)std::toupper((unsigned char)c);
    if (kernel_name == "BASICMPI")
      throw generic_gauxc_exception("BasicMPI ReductionDriver unavailable: no MPI in this build (use NCCL)");
    if (kernel_name != "DEFAULT" && kernel_name != "NCCL")
      throw generic_gauxc_exception("ReductionDriver Not Recognized: " + kernel_name);
    return std::make_shared<ReductionDriver>(rt.comm_size());
  }
  static ReductionDriver get_instance(const RuntimeEnvironment& rt, std::string kernel_name = "Default") {
    return *get_shared_instance(rt, std::move(kernel_name));
  }
};

// ---- functional ----------------------------------------------------------------------------------------
// Stands in for ExchCXX::XCFunctional: a functional NAME of the reference's functional_map
// (tests/standalone_driver.cxx:428-433); SVWN5, PBE, PBE0, LDA/SLATER, VWN5, SPW92 are implemented.
class functional_type {
  std::string spec_;
  bool polarized_ = false;

public:
  functional_type() = default;
  explicit functional_type(std::string spec, bool polarized = false) : spec_(std::move(spec)), polarized_(polarized) {}
  const std::string& spec() const { return spec_; }
  bool polarized() const { return polarized_; }
};

// ---- XC integrator ----------------------------------------------------------------------------------------
// include/gauxc/xc_integrator_settings.hpp:14-30
struct IntegratorSettingsXC {
  virtual ~IntegratorSettingsXC() noexcept = default;
};
struct IntegratorSettingsEXC_GRAD : public IntegratorSettingsXC {
  bool include_weight_derivatives = true;  // grid-weight contribution + translational invariance; false: Hellmann-Feynman
};

template <typename MatrixType>
class XCIntegrator {
public:
  using matrix_type = MatrixType;
  using value_type = typename MatrixType::value_type;
  using exc_vxc_type_rks = std::tuple<value_type, matrix_type>;
  using exc_vxc_type_uks = std::tuple<value_type, matrix_type, matrix_type>;

private:
  std::shared_ptr<GauXCIntegrator> h_;
  std::shared_ptr<GauXCFunctional> f_;
  std::shared_ptr<LoadBalancer> lb_;
  template <typename M>
  friend class XCIntegratorFactory;

public:
  // include/gauxc/xc_integrator/replicated/impl.hpp:108-120: VXC is (rows, cols) of P, fully overwritten
  exc_vxc_type_rks eval_exc_vxc(const MatrixType& P) {
    matrix_type VXC(P.rows(), P.cols());
    value_type EXC = 0;
    detail::StatusGuard g;
    gauxc_integrator_eval_exc_vxc_rks(&g.st, *h_, (int64_t)P.rows(), (int64_t)P.cols(), P.data(), (int64_t)P.rows(),
                                      &EXC, VXC.data(), (int64_t)VXC.rows());
    g.check();
    return std::make_tuple(EXC, std::move(VXC));
  }
  // UKS (include/gauxc/xc_integrator.hpp: eval_exc_vxc(Ps, Pz)): Ps = P_alpha + P_beta, Pz = P_alpha - P_beta;
  // LDA functionals constructed with polarized = true
  exc_vxc_type_uks eval_exc_vxc(const MatrixType& Ps, const MatrixType& Pz) {
    matrix_type VXCs(Ps.rows(), Ps.cols()), VXCz(Pz.rows(), Pz.cols());
    value_type EXC = 0;
    detail::StatusGuard g;
    gauxc_integrator_eval_exc_vxc_uks(&g.st, *h_, (int64_t)Ps.rows(), (int64_t)Ps.cols(), Ps.data(),
                                      (int64_t)Ps.rows(), Pz.data(), (int64_t)Pz.rows(), &EXC, VXCs.data(),
                                      (int64_t)VXCs.rows(), VXCz.data(), (int64_t)VXCz.rows());
    g.check();
    return std::make_tuple(EXC, std::move(VXCs), std::move(VXCz));
  }
  value_type eval_exc(const MatrixType& P) {
    value_type EXC = 0;
    detail::StatusGuard g;
    gauxc_integrator_eval_exc_rks(&g.st, *h_, (int64_t)P.rows(), (int64_t)P.cols(), P.data(), (int64_t)P.rows(), &EXC);
    g.check();
    return EXC;
  }
  // UKS EXC only
  value_type eval_exc(const MatrixType& Ps, const MatrixType& Pz) {
    value_type EXC = 0;
    detail::StatusGuard g;
    gauxc_integrator_eval_exc_uks(&g.st, *h_, (int64_t)Ps.rows(), (int64_t)Ps.cols(), Ps.data(), (int64_t)Ps.rows(),
                                  Pz.data(), (int64_t)Pz.rows(), &EXC);
    g.check();
    return EXC;
  }
  // EXC gradient w.r.t. the nuclear coordinates, 3 * natoms (include/gauxc/xc_integrator.hpp eval_exc_grad)
  std::vector<value_type> eval_exc_grad(const MatrixType& P, size_t natoms) {
    std::vector<value_type> grad(3 * natoms, 0);
    detail::StatusGuard g;
    gauxc_integrator_eval_exc_grad_rks(&g.st, *h_, (int64_t)P.rows(), (int64_t)P.cols(), P.data(), (int64_t)P.rows(),
                                       grad.data());
    g.check();
    return grad;
  }
  // UKS (include/gauxc/xc_integrator.hpp eval_exc_grad(Ps, Pz))
  std::vector<value_type> eval_exc_grad(const MatrixType& Ps, const MatrixType& Pz, size_t natoms,
                                        const IntegratorSettingsEXC_GRAD& settings = IntegratorSettingsEXC_GRAD{}) {
    std::vector<value_type> grad(3 * natoms, 0);
    detail::StatusGuard g;
    gauxc_b200_integrator_eval_exc_grad_uks(&g.st, *h_, (int64_t)Ps.rows(), (int64_t)Ps.cols(), Ps.data(),
                                            (int64_t)Ps.rows(), Pz.data(), (int64_t)Pz.rows(), grad.data(),
                                            settings.include_weight_derivatives ? 1 : 0);
    g.check();
    return grad;
  }
  // with IntegratorSettingsEXC_GRAD (include/gauxc/xc_integrator_settings.hpp:28-30)
  std::vector<value_type> eval_exc_grad(const MatrixType& P, size_t natoms, const IntegratorSettingsEXC_GRAD& settings) {
    std::vector<value_type> grad(3 * natoms, 0);
    detail::StatusGuard g;
    gauxc_b200_integrator_eval_exc_grad_rks(&g.st, *h_, (int64_t)P.rows(), (int64_t)P.cols(), P.data(),
                                            (int64_t)P.rows(), grad.data(), settings.include_weight_derivatives ? 1 : 0);
    g.check();
    return grad;
  }
  value_type integrate_den(const MatrixType& P) {
    value_type N_EL = 0;
    detail::StatusGuard g;
    gauxc_integrator_integrate_den(&g.st, *h_, (int64_t)P.rows(), (int64_t)P.cols(), P.data(), (int64_t)P.rows(), &N_EL);
    g.check();
    return N_EL;
  }
  const LoadBalancer& load_balancer() const { return *lb_; }
  const GauXCIntegrator& c_handle() const { return *h_; }
};

template <typename MatrixType>
class XCIntegratorFactory {
  ExecutionSpace ex_;
  std::string input_type_, integrator_kernel_, lwd_kernel_, reduction_kernel_;

public:
  using integrator_type = XCIntegrator<MatrixType>;
  XCIntegratorFactory() = delete;
  XCIntegratorFactory(ExecutionSpace ex, std::string integrator_input_type, std::string integrator_kernel_name,
                      std::string local_work_kernel_name, std::string reduction_kernel_name)
      : ex_(ex), input_type_(std::move(integrator_input_type)), integrator_kernel_(std::move(integrator_kernel_name)),
        lwd_kernel_(std::move(local_work_kernel_name)), reduction_kernel_(std::move(reduction_kernel_name)) {}

  std::shared_ptr<integrator_type> get_shared_instance(const functional_type& func, std::shared_ptr<LoadBalancer> lb) {
    auto xi = std::make_shared<integrator_type>();
    xi->lb_ = lb;
    detail::StatusGuard g;
    GauXCFunctional f = gauxc_functional_from_string(&g.st, func.spec().c_str(), func.polarized());
    g.check();
    xi->f_ = std::shared_ptr<GauXCFunctional>(new GauXCFunctional(f), [](GauXCFunctional* p) {
      GauXCStatus st{0, nullptr};
      gauxc_functional_delete(&st, p);
      gauxc_status_delete(&st);
      delete p;
    });
    GauXCIntegrator h = gauxc_integrator_new(&g.st, *xi->f_, lb->c_handle(), (GauXC_ExecutionSpace)(int)ex_,
                                             input_type_.c_str(), integrator_kernel_.c_str(), lwd_kernel_.c_str(),
                                             reduction_kernel_.c_str());
    g.check();
    xi->h_ = std::shared_ptr<GauXCIntegrator>(new GauXCIntegrator(h), [](GauXCIntegrator* p) {
      GauXCStatus st{0, nullptr};
      gauxc_integrator_delete(&st, p);
      gauxc_status_delete(&st);
      delete p;
    });
    return xi;
  }
  integrator_type get_instance(const functional_type& func, std::shared_ptr<LoadBalancer> lb) {
    return *get_shared_instance(func, std::move(lb));
  }
};

}  // namespace b200_capi
}  // namespace GauXC
