// The in-tree binding a GauXC maintainer would add, as a program that compiles and runs.
//
// The reference selects its device integrator through XCIntegratorFactory -> ReplicatedXCDeviceIntegrator<double>,
// whose hook is (include/gauxc/xc_integrator/replicated/replicated_xc_integrator_impl.hpp:33-196,
// replicated_xc_device_integrator.hpp:21-67)
//     virtual void eval_exc_vxc_(int64_t m, int64_t n, const value_type* P, int64_t ldp,
//                                value_type* VXC, int64_t ldvxc, value_type* EXC,
//                                const IntegratorSettingsXC&) = 0;
// The reference headers cannot be included here (they pull in ExchCXX / IntegratorXX, which are not vendored), so
// `ref::` below re-declares exactly that slice -- the base class with the hook, XCTask with the fields of
// include/gauxc/xc_task.hpp:25-62 -- and B200ReplicatedXCDeviceIntegrator is the adaptor, written against the C ABI
// of libgauxc_b200 only.  It hands the reference's OWN task list (grid, batching, screening, weights) to the device
// path through gauxc_b200_load_balancer_set_tasks, so the reference's LoadBalancer / MolecularWeights stay in charge.
#include <gauxc_b200.h>

#include <array>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <stdexcept>
#include <string>
#include <vector>

namespace ref {  // stand-ins with the reference's signatures
struct IntegratorSettingsXC { virtual ~IntegratorSettingsXC() = default; };
struct XCTask {
  int32_t iParent;
  std::vector<std::array<double, 3>> points;
  std::vector<double> weights;
  struct { std::vector<int32_t> shell_list; int32_t nbe; } bfn_screening;
  double dist_nearest;
};
template <typename T>
struct ReplicatedXCDeviceIntegrator {
  virtual ~ReplicatedXCDeviceIntegrator() = default;
  virtual void eval_exc_vxc_(int64_t m, int64_t n, const T* P, int64_t ldp, T* VXC, int64_t ldvxc, T* EXC,
                             const IntegratorSettingsXC&) = 0;
};
}  // namespace ref

static void check(GauXCStatus& st) {
  if (st.code) {
    std::string msg = st.message ? st.message : "unknown";
    gauxc_status_delete(&st);
    throw std::runtime_error(msg);  // GAUXC_GENERIC_EXCEPTION in tree
  }
}

struct B200ReplicatedXCDeviceIntegrator : ref::ReplicatedXCDeviceIntegrator<double> {
  GauXCIntegrator handle_{};
  GauXCLoadBalancer lb_{};
  GauXCFunctional func_{};

  // lb: a libgauxc_b200 LoadBalancer built from the same Molecule / BasisSet (its own task list is replaced);
  // tasks: the reference LoadBalancer's get_tasks(), weights already modified (lb.state().modified_weights_are_stored)
  B200ReplicatedXCDeviceIntegrator(GauXCLoadBalancer lb, const std::vector<ref::XCTask>& tasks, const char* functional)
      : lb_(lb) {
    GauXCStatus st{0, nullptr};
    std::vector<int32_t> npts, ipar, nsh, sl;
    std::vector<double> dn, pts, w;
    for (auto& t : tasks) {
      npts.push_back((int32_t)t.points.size());
      ipar.push_back(t.iParent);
      dn.push_back(t.dist_nearest);
      nsh.push_back((int32_t)t.bfn_screening.shell_list.size());
      for (auto& p : t.points) pts.insert(pts.end(), p.begin(), p.end());
      w.insert(w.end(), t.weights.begin(), t.weights.end());
      sl.insert(sl.end(), t.bfn_screening.shell_list.begin(), t.bfn_screening.shell_list.end());
    }
    gauxc_b200_load_balancer_set_tasks(&st, lb_, (int64_t)tasks.size(), npts.data(), ipar.data(), dn.data(), pts.data(),
                                       w.data(), nsh.data(), sl.data(), /*weights_are_modified=*/1);
    check(st);
    func_ = gauxc_functional_from_string(&st, functional, false);
    check(st);
    handle_ = gauxc_integrator_new(&st, func_, lb_, GauXC_ExecutionSpace_Device, "Replicated", "Default", "Default",
                                   "Default");
    check(st);
  }
  ~B200ReplicatedXCDeviceIntegrator() override {
    GauXCStatus st{0, nullptr};
    gauxc_integrator_delete(&st, &handle_);
    gauxc_functional_delete(&st, &func_);
    gauxc_status_delete(&st);
  }
  void eval_exc_vxc_(int64_t m, int64_t n, const double* P, int64_t ldp, double* VXC, int64_t ldvxc, double* EXC,
                     const ref::IntegratorSettingsXC&) override {
    GauXCStatus st{0, nullptr};
    gauxc_integrator_eval_exc_vxc_rks(&st, handle_, m, n, P, ldp, EXC, VXC, ldvxc);
    check(st);
  }
};

// ---- demo: the "reference" task list is produced here by a second libgauxc_b200 LoadBalancer + Device weights and
//      carried over as ref::XCTask objects; the adaptor must reproduce the direct call exactly -------------------------
int main() {
  try {
    GauXCStatus st{0, nullptr};
    const GauXCAtom atoms[3] = {{8, 0., -0.07579, 0.}, {1, 0.86681, 0.60144, 0.}, {1, -0.86681, 0.60144, 0.}};
    GauXCMolecule mol = gauxc_molecule_new_from_atoms(&st, atoms, 3); check(st);
    GauXCShell sh[5] = {};
    const double a_o1[3] = {130.70932, 23.808861, 6.4436083}, c_s[3] = {0.15432897, 0.53532814, 0.44463454};
    const double a_o2[3] = {5.0331513, 1.1695961, 0.3803890}, c_2s[3] = {-0.09996723, 0.39951283, 0.70011547};
    const double c_2p[3] = {0.15591627, 0.60768372, 0.39195739}, a_h[3] = {3.42525091, 0.62391373, 0.16885540};
    auto set = [&](GauXCShell& s, int l, const double* a, const double* c, const GauXCAtom& at) {
      s.l = l; s.pure = true; s.nprim = 3; s.shell_tolerance = 1e-10;
      for (int k = 0; k < 3; ++k) { s.exponents[k] = a[k]; s.coefficients[k] = c[k]; }
      s.origin[0] = at.x; s.origin[1] = at.y; s.origin[2] = at.z;
    };
    set(sh[0], 0, a_o1, c_s, atoms[0]); set(sh[1], 0, a_o2, c_2s, atoms[0]); set(sh[2], 1, a_o2, c_2p, atoms[0]);
    set(sh[3], 0, a_h, c_s, atoms[1]); set(sh[4], 0, a_h, c_s, atoms[2]);
    GauXCBasisSet basis = gauxc_basisset_new_from_shells(&st, sh, 5, true); check(st);
    const int64_t nbf = gauxc_b200_basisset_nbf(&st, basis); check(st);
    GauXCMolGrid mg = gauxc_molgrid_new_default(&st, mol, GauXC_PruningScheme_Unpruned, 512, GauXC_RadialQuad_MuraKnowles,
                                                GauXC_AtomicGridSizeDefault_FineGrid); check(st);
    GauXCRuntimeEnvironment rt = gauxc_device_runtime_environment_new(&st, 0.5); check(st);
    GauXCLoadBalancerFactory lbf = gauxc_load_balancer_factory_new(&st, GauXC_ExecutionSpace_Host, "Default"); check(st);
    GauXCLoadBalancer lb_ref = gauxc_load_balancer_factory_get_instance(&st, lbf, rt, mol, mg, basis); check(st);
    GauXCLoadBalancer lb_b200 = gauxc_load_balancer_factory_get_instance(&st, lbf, rt, mol, mg, basis); check(st);
    GauXCMolecularWeightsSettings ws = {GauXC_XCWeightAlg_SSF, false};
    GauXCMolecularWeightsFactory mwf = gauxc_molecular_weights_factory_new(&st, GauXC_ExecutionSpace_Device, "Default", ws);
    check(st);
    GauXCMolecularWeights mw = gauxc_molecular_weights_factory_get_instance(&st, mwf); check(st);
    gauxc_molecular_weights_modify_weights(&st, mw, lb_ref); check(st);

    std::vector<double> P((size_t)(nbf * nbf), 0.), V1(P.size()), V2(P.size());
    for (int64_t i = 0; i < nbf; ++i) P[(size_t)(i * nbf + i)] = 0.7;
    double exc_direct = 0., exc_adaptor = 0.;
    GauXCFunctional f = gauxc_functional_from_string(&st, "PBE", false); check(st);
    GauXCIntegrator direct = gauxc_integrator_new(&st, f, lb_ref, GauXC_ExecutionSpace_Device, "Replicated", "Default",
                                                  "Default", "Default"); check(st);
    gauxc_integrator_eval_exc_vxc_rks(&st, direct, nbf, nbf, P.data(), nbf, &exc_direct, V1.data(), nbf); check(st);

    // the "reference" tasks
    const int64_t nt = gauxc_b200_load_balancer_ntasks(&st, lb_ref); check(st);
    std::vector<int32_t> ip(nt), np(nt), nbe(nt), nsh(nt);
    std::vector<double> dn(nt);
    gauxc_b200_load_balancer_task_info(&st, lb_ref, ip.data(), np.data(), nbe.data(), nsh.data(), dn.data()); check(st);
    std::vector<ref::XCTask> tasks((size_t)nt);
    for (int64_t t = 0; t < nt; ++t) {
      auto& x = tasks[(size_t)t];
      x.iParent = ip[t]; x.dist_nearest = dn[t]; x.bfn_screening.nbe = nbe[t];
      x.points.resize(np[t]); x.weights.resize(np[t]); x.bfn_screening.shell_list.resize(nsh[t]);
      gauxc_b200_load_balancer_get_task(&st, lb_ref, t, x.points[0].data(), x.weights.data(),
                                        x.bfn_screening.shell_list.data()); check(st);
    }
    B200ReplicatedXCDeviceIntegrator adaptor(lb_b200, tasks, "PBE");
    ref::ReplicatedXCDeviceIntegrator<double>& base = adaptor;
    base.eval_exc_vxc_(nbf, nbf, P.data(), nbf, V2.data(), nbf, &exc_adaptor, ref::IntegratorSettingsXC{});
    double dv = 0.;
    for (size_t i = 0; i < V1.size(); ++i) dv = std::fmax(dv, std::fabs(V1[i] - V2[i]));
    std::printf("EXC direct %.12f adaptor %.12f |dEXC| %.2e max|dVXC| %.2e\n", exc_direct, exc_adaptor,
                std::fabs(exc_direct - exc_adaptor), dv);
    return (std::fabs(exc_direct - exc_adaptor) < 1e-12 && dv < 1e-12) ? 0 : 1;
  } catch (const std::exception& e) {
    std::fprintf(stderr, "adaptor: %s\n", e.what());
    return 1;
  }
}
