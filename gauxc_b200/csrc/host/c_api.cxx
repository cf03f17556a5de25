// extern "C" shims over the C++ host layer; same pattern as the reference's src/c-api/
// (try/catch -> GauXCStatus, objects = {GauXCHeader, void*} validated by type tag;
// src/c-api/c_status.hpp:23-43, src/c-api/c_xc_integrator.cxx:86-147).
#include "../../../include/gauxc_b200.h"
#include "../cuda/xc_functionals.cuh"
#include "../cuda/xc_functionals_pol_gga.cuh"
#include "hdf5_io.hpp"
#include "xc_integrator.hpp"
#include <cstdlib>
#include <cstring>

using namespace GauXC;

namespace GauXC {
void device_eval_collocation(const BasisSet& basis, const std::vector<int32_t>& shell_list,
                             int64_t npts, const double* points, double* eval, double* dx,
                             double* dy, double* dz, double* hess6 = nullptr);
double device_probe_peak(int which);
int device_count();
void device_set(int dev);
void device_allreduce(double* dptr, size_t n);
}  // namespace GauXC

namespace {

void status_init(GauXCStatus* st) {
  if (st) {
    st->code = 0;
    std::free(st->message);
    st->message = nullptr;
  }
}
void status_fail(GauXCStatus* st, const char* msg) {
  if (st) {
    st->code = 1;
    if (st->message) std::free(st->message);
    const size_t len = std::strlen(msg) + 1;
    st->message = (char*)std::malloc(len);
    std::memcpy(st->message, msg, len);
  } else {
    GAUXC_GENERIC_EXCEPTION(msg);
  }
}

#define C_TRY(st) \
  status_init(st); \
  try {
#define C_CATCH(st)                       \
  }                                       \
  catch (const std::exception& e) {       \
    status_fail(st, e.what());            \
  }                                       \
  catch (...) {                           \
    status_fail(st, "Unknown exception"); \
  }

template <typename T>
T* checked(const void* ptr, const GauXCHeader& hdr, GauXC_Type type, const char* what) {
  if (hdr.type != type || ptr == nullptr)
    GAUXC_GENERIC_EXCEPTION(std::string("Invalid handle: expected ") + what);
  return (T*)ptr;
}

struct LBFactory {
  ExecutionSpace ex;
  std::string kernel;
};
struct MWFactory {
  ExecutionSpace ex;
  std::string kernel;
  MolecularWeightsSettings settings;
};
using RuntimePtr = std::shared_ptr<RuntimeEnvironment>;
using LBPtr = std::shared_ptr<LoadBalancer>;
using FuncPtr = std::shared_ptr<XCFunctional>;

std::string upper(std::string s) {
  for (auto& c : s) c = (char)::toupper(c);
  return s;
}

#define MOL(m) checked<Molecule>((m).ptr, (m).hdr, GauXC_Type_Molecule, "Molecule")
#define BAS(b) checked<BasisSet>((b).ptr, (b).hdr, GauXC_Type_BasisSet, "BasisSet")
#define MG(g) checked<MolGrid>((g).ptr, (g).hdr, GauXC_Type_MolGrid, "MolGrid")
#define RT(r) checked<RuntimePtr>((r).ptr, (r).hdr, GauXC_Type_RuntimeEnvironment, "RuntimeEnvironment")
#define LB(l) checked<LBPtr>((l).ptr, (l).hdr, GauXC_Type_LoadBalancer, "LoadBalancer")
#define FN(f) checked<FuncPtr>((f).ptr, (f).hdr, GauXC_Type_Functional, "Functional")
#define INTG(i) checked<XCIntegrator>((i).ptr, (i).hdr, GauXC_Type_Integrator, "Integrator")
#define MW(w) checked<MolecularWeights>((w).ptr, (w).hdr, GauXC_Type_MolecularWeights, "MolecularWeights")

void delete_by_type(GauXC_Type type, void* ptr) {
  switch (type) {
    case GauXC_Type_Molecule: delete (Molecule*)ptr; break;
    case GauXC_Type_BasisSet: delete (BasisSet*)ptr; break;
    case GauXC_Type_MolGrid: delete (MolGrid*)ptr; break;
    case GauXC_Type_RuntimeEnvironment: delete (RuntimePtr*)ptr; break;
    case GauXC_Type_LoadBalancer: delete (LBPtr*)ptr; break;
    case GauXC_Type_LoadBalancerFactory: delete (LBFactory*)ptr; break;
    case GauXC_Type_MolecularWeights: delete (MolecularWeights*)ptr; break;
    case GauXC_Type_MolecularWeightsFactory: delete (MWFactory*)ptr; break;
    case GauXC_Type_Functional: delete (FuncPtr*)ptr; break;
    case GauXC_Type_Integrator: delete (XCIntegrator*)ptr; break;
    default: GAUXC_GENERIC_EXCEPTION("Unknown object type");
  }
}

}  // namespace

extern "C" {

void gauxc_status_delete(GauXCStatus* status) {
  if (status) {
    std::free(status->message);
    status->message = nullptr;
    status->code = 0;
  }
}

// every handle struct starts with {GauXCHeader hdr; void* ptr;}
struct GenericHandle {
  GauXCHeader hdr;
  void* ptr;
};

void gauxc_object_delete(GauXCStatus* status, void** handle) {
  C_TRY(status)
  if (!handle || !*handle) return;
  auto* h = (GenericHandle*)(*handle);
  if (h->ptr) delete_by_type(h->hdr.type, h->ptr);
  h->ptr = nullptr;
  C_CATCH(status)
}
void gauxc_objects_delete(GauXCStatus* status, void** handles, size_t n) {
  for (size_t i = 0; i < n; ++i) gauxc_object_delete(status, &handles[i]);
}

#define DEFINE_DELETE(NAME, CTYPE)                          \
  void NAME(GauXCStatus* status, CTYPE* h) {                \
    C_TRY(status)                                           \
    if (h && h->ptr) delete_by_type(h->hdr.type, h->ptr);   \
    if (h) h->ptr = nullptr;                                \
    C_CATCH(status)                                         \
  }
DEFINE_DELETE(gauxc_molecule_delete, GauXCMolecule)
DEFINE_DELETE(gauxc_basisset_delete, GauXCBasisSet)
DEFINE_DELETE(gauxc_molgrid_delete, GauXCMolGrid)
DEFINE_DELETE(gauxc_load_balancer_delete, GauXCLoadBalancer)
DEFINE_DELETE(gauxc_load_balancer_factory_delete, GauXCLoadBalancerFactory)
DEFINE_DELETE(gauxc_molecular_weights_delete, GauXCMolecularWeights)
DEFINE_DELETE(gauxc_molecular_weights_factory_delete, GauXCMolecularWeightsFactory)
DEFINE_DELETE(gauxc_functional_delete, GauXCFunctional)
DEFINE_DELETE(gauxc_integrator_delete, GauXCIntegrator)

void gauxc_runtime_environment_delete(GauXCStatus* status, GauXCRuntimeEnvironment* env) {
  C_TRY(status)
  if (env && env->ptr) delete (RuntimePtr*)env->ptr;
  if (env) env->ptr = env->device_ptr = nullptr;
  C_CATCH(status)
}

// ---- molecule -------------------------------------------------------------------------
GauXCMolecule gauxc_molecule_new(GauXCStatus* status) {
  GauXCMolecule m{{GauXC_Type_Molecule}, nullptr};
  C_TRY(status)
  m.ptr = new Molecule();
  C_CATCH(status)
  return m;
}
GauXCMolecule gauxc_molecule_new_from_atoms(GauXCStatus* status, const GauXCAtom* atoms, size_t natoms) {
  GauXCMolecule m{{GauXC_Type_Molecule}, nullptr};
  C_TRY(status)
  auto* mol = new Molecule();
  for (size_t i = 0; i < natoms; ++i) mol->push_back({atoms[i].Z, atoms[i].x, atoms[i].y, atoms[i].z});
  m.ptr = mol;
  C_CATCH(status)
  return m;
}
size_t gauxc_molecule_natoms(GauXCStatus* status, const GauXCMolecule mol) {
  size_t n = 0;
  C_TRY(status)
  n = MOL(mol)->size();
  C_CATCH(status)
  return n;
}
bool gauxc_molecule_equal(GauXCStatus* status, const GauXCMolecule a, const GauXCMolecule b) {
  bool eq = false;
  C_TRY(status)
  auto *x = MOL(a), *y = MOL(b);
  eq = x->size() == y->size();
  for (size_t i = 0; eq && i < x->size(); ++i)
    eq = (*x)[i].Z == (*y)[i].Z && (*x)[i].x == (*y)[i].x && (*x)[i].y == (*y)[i].y && (*x)[i].z == (*y)[i].z;
  C_CATCH(status)
  return eq;
}

// ---- basis ----------------------------------------------------------------------------
GauXCBasisSet gauxc_basisset_new(GauXCStatus* status) {
  GauXCBasisSet b{{GauXC_Type_BasisSet}, nullptr};
  C_TRY(status)
  b.ptr = new BasisSet();
  C_CATCH(status)
  return b;
}
GauXCBasisSet gauxc_basisset_new_from_shells(GauXCStatus* status, const GauXCShell* shells, size_t nshells,
                                             bool normalize) {
  GauXCBasisSet b{{GauXC_Type_BasisSet}, nullptr};
  C_TRY(status)
  auto* bs = new BasisSet();
  try {
    for (size_t i = 0; i < nshells; ++i) {
      const auto& s = shells[i];
      Shell sh(s.nprim, s.l, s.pure ? 1 : 0, s.exponents, s.coefficients, s.origin, normalize);
      if (s.shell_tolerance > 0.) sh.set_shell_tolerance(s.shell_tolerance);
      bs->push_back(sh);
    }
  } catch (...) {
    delete bs;
    throw;
  }
  b.ptr = bs;
  C_CATCH(status)
  return b;
}

// ---- molgrid ---------------------------------------------------------------------------
GauXCMolGrid gauxc_molgrid_new_default(GauXCStatus* status, const GauXCMolecule mol,
                                       enum GauXC_PruningScheme ps, int64_t batchsize,
                                       enum GauXC_RadialQuad rq, enum GauXC_AtomicGridSizeDefault gs) {
  GauXCMolGrid g{{GauXC_Type_MolGrid}, nullptr};
  C_TRY(status)
  g.ptr = new MolGrid(create_default_molgrid(*MOL(mol), (PruningScheme)ps, batchsize, (RadialQuad)rq,
                                             (AtomicGridSizeDefault)gs));
  C_CATCH(status)
  return g;
}

// ---- runtime -----------------------------------------------------------------------------
GauXCRuntimeEnvironment gauxc_runtime_environment_new(GauXCStatus* status) {
  GauXCRuntimeEnvironment r{{GauXC_Type_RuntimeEnvironment}, nullptr, nullptr};
  C_TRY(status)
  r.ptr = new RuntimePtr(std::make_shared<RuntimeEnvironment>());
  C_CATCH(status)
  return r;
}
GauXCRuntimeEnvironment gauxc_device_runtime_environment_new(GauXCStatus* status, double fill_fraction) {
  GauXCRuntimeEnvironment r{{GauXC_Type_RuntimeEnvironment}, nullptr, nullptr};
  C_TRY(status)
  auto d = std::make_shared<DeviceRuntimeEnvironment>(fill_fraction);
  r.ptr = new RuntimePtr(d);
  r.device_ptr = d.get();
  C_CATCH(status)
  return r;
}
GauXCRuntimeEnvironment gauxc_device_runtime_environment_new_mem(GauXCStatus* status, void* mem, size_t sz) {
  GauXCRuntimeEnvironment r{{GauXC_Type_RuntimeEnvironment}, nullptr, nullptr};
  C_TRY(status)
  auto d = std::make_shared<DeviceRuntimeEnvironment>(mem, sz);
  r.ptr = new RuntimePtr(d);
  r.device_ptr = d.get();
  C_CATCH(status)
  return r;
}
int gauxc_runtime_environment_comm_rank(GauXCStatus* status, const GauXCRuntimeEnvironment env) {
  int v = 0;
  C_TRY(status)
  v = (*RT(env))->comm_rank();
  C_CATCH(status)
  return v;
}
int gauxc_runtime_environment_comm_size(GauXCStatus* status, const GauXCRuntimeEnvironment env) {
  int v = 1;
  C_TRY(status)
  v = (*RT(env))->comm_size();
  C_CATCH(status)
  return v;
}
void gauxc_b200_runtime_environment_set_comm(GauXCStatus* status, GauXCRuntimeEnvironment env, int rank,
                                             int size) {
  C_TRY(status)
  (*RT(env))->set_comm(rank, size);
  C_CATCH(status)
}

// ---- load balancer ---------------------------------------------------------------------------
GauXCLoadBalancerFactory gauxc_load_balancer_factory_new(GauXCStatus* status, enum GauXC_ExecutionSpace ex,
                                                         const char* kernel_name) {
  GauXCLoadBalancerFactory f{{GauXC_Type_LoadBalancerFactory}, nullptr};
  C_TRY(status)
  // src/load_balancer/host/load_balancer_host_factory.cxx:28-40 names
  const std::string k = upper(kernel_name ? kernel_name : "Default");
  if (k != "DEFAULT" && k != "REPLICATED" && k != "REPLICATED-PETITE" && k != "REPLICATED-FILLIN")
    GAUXC_GENERIC_EXCEPTION("LoadBalancer Kernel Not Recognized: " + k);
  f.ptr = new LBFactory{(ExecutionSpace)ex, k};
  C_CATCH(status)
  return f;
}
GauXCLoadBalancer gauxc_load_balancer_factory_get_instance(GauXCStatus* status,
                                                           const GauXCLoadBalancerFactory factory,
                                                           const GauXCRuntimeEnvironment env,
                                                           const GauXCMolecule mol, const GauXCMolGrid mg,
                                                           const GauXCBasisSet basis) {
  GauXCLoadBalancer lb{{GauXC_Type_LoadBalancer}, nullptr};
  C_TRY(status)
  auto* f = checked<LBFactory>(factory.ptr, factory.hdr, GauXC_Type_LoadBalancerFactory, "LoadBalancerFactory");
  lb.ptr = new LBPtr(std::make_shared<LoadBalancer>(*RT(env), *MOL(mol), *MG(mg), *BAS(basis), f->kernel, f->ex));
  C_CATCH(status)
  return lb;
}

// ---- molecular weights ------------------------------------------------------------------------
GauXCMolecularWeightsFactory gauxc_molecular_weights_factory_new(GauXCStatus* status,
                                                                 enum GauXC_ExecutionSpace ex,
                                                                 const char* lwd,
                                                                 GauXCMolecularWeightsSettings settings) {
  GauXCMolecularWeightsFactory f{{GauXC_Type_MolecularWeightsFactory}, nullptr};
  C_TRY(status)
  MolecularWeightsSettings s;
  s.weight_alg = (XCWeightAlg)settings.weight_alg;
  s.becke_size_adjustment = settings.becke_size_adjustment;
  f.ptr = new MWFactory{(ExecutionSpace)ex, lwd ? lwd : "Default", s};
  C_CATCH(status)
  return f;
}
GauXCMolecularWeights gauxc_molecular_weights_factory_get_instance(GauXCStatus* status,
                                                                   const GauXCMolecularWeightsFactory factory) {
  GauXCMolecularWeights w{{GauXC_Type_MolecularWeights}, nullptr};
  C_TRY(status)
  auto* f = checked<MWFactory>(factory.ptr, factory.hdr, GauXC_Type_MolecularWeightsFactory,
                               "MolecularWeightsFactory");
  w.ptr = new MolecularWeights(f->ex, f->kernel, f->settings);
  C_CATCH(status)
  return w;
}
void gauxc_molecular_weights_modify_weights(GauXCStatus* status, const GauXCMolecularWeights mw,
                                            const GauXCLoadBalancer lb) {
  C_TRY(status)
  MW(mw)->modify_weights(**LB(lb));
  C_CATCH(status)
}
double gauxc_b200_molecular_weights_last_ms(GauXCStatus* status, const GauXCMolecularWeights mw) {
  double v = 0;
  C_TRY(status)
  auto& t = MW(mw)->get_timings().ms;
  auto it = t.find("MolecularWeights");
  v = it == t.end() ? 0. : it->second;
  C_CATCH(status)
  return v;
}

// ---- functional --------------------------------------------------------------------------------
GauXCFunctional gauxc_functional_from_string(GauXCStatus* status, const char* spec, bool polarized) {
  GauXCFunctional f{{GauXC_Type_Functional}, nullptr};
  C_TRY(status)
  f.ptr = new FuncPtr(std::make_shared<XCFunctional>(functional_from_string(spec ? spec : "", polarized)));
  C_CATCH(status)
  return f;
}
GauXCFunctional gauxc_functional_from_enum(GauXCStatus* status, enum GauXC_Functional functional_type,
                                           bool polarized) {
  // src/c-api/c_functional.cxx:96-114 (ExchCXX::Functional enumerators in the same order)
  GauXCFunctional f{{GauXC_Type_Functional}, nullptr};
  C_TRY(status)
  const char* name = nullptr;
  switch (functional_type) {
    case GauXC_Functional_SVWN5: name = "SVWN5"; break;
    case GauXC_Functional_BLYP: name = "BLYP"; break;
    case GauXC_Functional_B3LYP: name = "B3LYP"; break;
    case GauXC_Functional_PBE: name = "PBE"; break;
    case GauXC_Functional_revPBE: name = "REVPBE"; break;
    case GauXC_Functional_PBE0: name = "PBE0"; break;
    case GauXC_Functional_LDA: name = "LDA"; break;
    case GauXC_Functional_SPW92: name = "SPW92"; break;
    case GauXC_Functional_VWN5: name = "VWN5"; break;
    case GauXC_Functional_revPBE0: name = "REVPBE0"; break;
    default: GAUXC_GENERIC_EXCEPTION("Functional NYI in B200 path: enum " + std::to_string((int)functional_type));
  }
  f.ptr = new FuncPtr(std::make_shared<XCFunctional>(functional_from_string(name, polarized)));
  C_CATCH(status)
  return f;
}
void gauxc_b200_functional_eval_host(GauXCStatus* status, const GauXCFunctional functional, int64_t npts,
                                     const double* rho, const double* sigma, double* eps, double* vrho,
                                     double* vsigma) {
  C_TRY(status)
  const auto& d = (*FN(functional))->desc;
  for (int64_t i = 0; i < npts; ++i) {
    const auto o = gxb::eval_functional(d, rho[i], sigma ? sigma[i] : 0.);
    eps[i] = o.eps;
    vrho[i] = o.vrho;
    if (vsigma) vsigma[i] = o.vsigma;
  }
  C_CATCH(status)
}

void gauxc_b200_functional_eval_host_pol(GauXCStatus* status, const GauXCFunctional functional, int64_t npts,
                                         const double* rho_a, const double* rho_b, double* eps, double* vrho_a,
                                         double* vrho_b) {
  C_TRY(status)
  const auto& f = **FN(functional);
  if (f.is_gga()) GAUXC_GENERIC_EXCEPTION("Polarized GGA NYI in B200 path");
  for (int64_t i = 0; i < npts; ++i) {
    const auto o = gxb::eval_functional_pol_lda(f.desc, rho_a[i], rho_b[i]);
    eps[i] = o.eps;
    vrho_a[i] = o.va;
    vrho_b[i] = o.vb;
  }
  C_CATCH(status)
}

void gauxc_b200_functional_eval_host_pol_full(GauXCStatus* status, const GauXCFunctional functional, int64_t npts,
                                              const double* rho2, const double* gamma3, double* eps, double* vrho2,
                                              double* vgamma3) {
  C_TRY(status)
  const auto& f = **FN(functional);
  for (int64_t i = 0; i < npts; ++i) {
    const auto o = gxb::eval_functional_pol(f.desc, rho2[2 * i], rho2[2 * i + 1], gamma3[3 * i], gamma3[3 * i + 1],
                                            gamma3[3 * i + 2]);
    eps[i] = o.eps;
    vrho2[2 * i] = o.va; vrho2[2 * i + 1] = o.vb;
    vgamma3[3 * i] = o.vaa; vgamma3[3 * i + 1] = o.vab; vgamma3[3 * i + 2] = o.vbb;
  }
  C_CATCH(status)
}

void gauxc_b200_functional_eval_host_pol_gga(GauXCStatus* status, int nkern, const int* kern, const double* coeff,
                                             int64_t npts, const double* rho2, const double* gamma3, double* eps,
                                             double* vrho2, double* vgamma3) {
  C_TRY(status)
  for (int64_t i = 0; i < npts; ++i) {
    const auto o = gxb::eval_pol_gga(nkern, kern, coeff, rho2[2 * i], rho2[2 * i + 1], gamma3[3 * i],
                                     gamma3[3 * i + 1], gamma3[3 * i + 2]);
    eps[i] = o.eps;
    vrho2[2 * i] = o.va; vrho2[2 * i + 1] = o.vb;
    vgamma3[3 * i] = o.vaa; vgamma3[3 * i + 1] = o.vab; vgamma3[3 * i + 2] = o.vbb;
  }
  C_CATCH(status)
}

// ---- integrator ----------------------------------------------------------------------------------
GauXCIntegrator gauxc_integrator_new(GauXCStatus* status, const GauXCFunctional functional,
                                     const GauXCLoadBalancer lb, enum GauXC_ExecutionSpace ex,
                                     const char* input_type, const char* integrator_kernel,
                                     const char* lwd_kernel, const char* reduction_kernel) {
  GauXCIntegrator h{{GauXC_Type_Integrator}, nullptr};
  C_TRY(status)
  h.ptr = new XCIntegrator((ExecutionSpace)ex, input_type ? input_type : "Replicated",
                           integrator_kernel ? integrator_kernel : "Default",
                           lwd_kernel ? lwd_kernel : "Default",
                           reduction_kernel ? reduction_kernel : "Default", *FN(functional), *LB(lb));
  C_CATCH(status)
  return h;
}
void gauxc_integrator_integrate_den(GauXCStatus* status, const GauXCIntegrator integrator, const int64_t m,
                                    const int64_t n, const double* P, const int64_t ldp, double* den) {
  C_TRY(status)
  INTG(integrator)->integrate_den(m, n, P, ldp, den);
  C_CATCH(status)
}
void gauxc_integrator_eval_exc_rks(GauXCStatus* status, const GauXCIntegrator integrator, const int64_t m,
                                   const int64_t n, const double* P, const int64_t ldp, double* exc) {
  C_TRY(status)
  INTG(integrator)->eval_exc(m, n, P, ldp, exc);
  C_CATCH(status)
}
void gauxc_integrator_eval_exc_vxc_rks(GauXCStatus* status, const GauXCIntegrator integrator,
                                       const int64_t m, const int64_t n, const double* P, const int64_t ldp,
                                       double* exc, double* vxc, const int64_t vxc_ld) {
  C_TRY(status)
  INTG(integrator)->eval_exc_vxc(m, n, P, ldp, vxc, vxc_ld, exc);
  C_CATCH(status)
}
void gauxc_integrator_eval_exc_vxc_uks(GauXCStatus* status, const GauXCIntegrator integrator, const int64_t m,
                                       const int64_t n, const double* Ps, const int64_t ldps, const double* Pz,
                                       const int64_t ldpz, double* exc, double* vxc_s, const int64_t ldvs,
                                       double* vxc_z, const int64_t ldvz) {
  C_TRY(status)
  INTG(integrator)->eval_exc_vxc_uks(m, n, Ps, ldps, Pz, ldpz, vxc_s, ldvs, vxc_z, ldvz, exc);
  C_CATCH(status)
}
void gauxc_integrator_eval_exc_grad_rks(GauXCStatus* status, const GauXCIntegrator integrator, const int64_t m,
                                        const int64_t n, const double* P, const int64_t ldp, double* exc_grad) {
  C_TRY(status)
  INTG(integrator)->eval_exc_grad(m, n, P, ldp, exc_grad);
  C_CATCH(status)
}
void gauxc_b200_integrator_eval_exc_grad_rks(GauXCStatus* status, const GauXCIntegrator integrator, const int64_t m,
                                             const int64_t n, const double* P, const int64_t ldp, double* exc_grad,
                                             int include_weight_derivatives) {
  C_TRY(status)
  INTG(integrator)->eval_exc_grad(m, n, P, ldp, exc_grad, include_weight_derivatives != 0);
  C_CATCH(status)
}
void gauxc_integrator_eval_exc_uks(GauXCStatus* status, const GauXCIntegrator integrator, const int64_t m,
                                   const int64_t n, const double* Ps, const int64_t ldps, const double* Pz,
                                   const int64_t ldpz, double* exc) {
  C_TRY(status)
  INTG(integrator)->eval_exc_uks(m, n, Ps, ldps, Pz, ldpz, exc);
  C_CATCH(status)
}
// Entry points of include/gauxc/c/xc_integrator.h that lie outside the LDA/GGA RKS/UKS EXC/VXC path
// (SURVEY.md section 8f): exported for link compatibility, status code 1 with a "NYI" message.
#define GAUXC_B200_NYI(what)                                          \
  C_TRY(status)                                                       \
  GAUXC_GENERIC_EXCEPTION(what " NYI in B200 path");                  \
  C_CATCH(status)
void gauxc_integrator_eval_exc_gks(GauXCStatus* status, const GauXCIntegrator, const int64_t, const int64_t,
                                   const double*, const int64_t, const double*, const int64_t, const double*,
                                   const int64_t, const double*, const int64_t, double*) {
  GAUXC_B200_NYI("GKS EXC")
}
void gauxc_integrator_eval_exc_vxc_gks(GauXCStatus* status, const GauXCIntegrator, const int64_t, const int64_t,
                                       const double*, const int64_t, const double*, const int64_t, const double*,
                                       const int64_t, const double*, const int64_t, double*, double*, const int64_t,
                                       double*, const int64_t, double*, const int64_t, double*, const int64_t) {
  GAUXC_B200_NYI("GKS EXC/VXC")
}
void gauxc_integrator_eval_exc_grad_uks(GauXCStatus* status, const GauXCIntegrator integrator, const int64_t m,
                                        const int64_t n, const double* Ps, const int64_t ldps, const double* Pz,
                                        const int64_t ldpz, double* exc_grad) {
  C_TRY(status)
  INTG(integrator)->eval_exc_grad_uks(m, n, Ps, ldps, Pz, ldpz, exc_grad);
  C_CATCH(status)
}
void gauxc_b200_integrator_eval_exc_grad_uks(GauXCStatus* status, const GauXCIntegrator integrator, const int64_t m,
                                             const int64_t n, const double* Ps, const int64_t ldps, const double* Pz,
                                             const int64_t ldpz, double* exc_grad, int include_weight_derivatives) {
  C_TRY(status)
  INTG(integrator)->eval_exc_grad_uks(m, n, Ps, ldps, Pz, ldpz, exc_grad, include_weight_derivatives != 0);
  C_CATCH(status)
}
void gauxc_integrator_eval_exx_rks(GauXCStatus* status, const GauXCIntegrator, const int64_t, const int64_t,
                                   const double*, const int64_t, double*, const int64_t) {
  GAUXC_B200_NYI("EXX")
}
void gauxc_integrator_eval_fxc_contraction_rks(GauXCStatus* status, const GauXCIntegrator, const int64_t,
                                               const int64_t, const double*, const int64_t, const double*,
                                               const int64_t, double*, const int64_t) {
  GAUXC_B200_NYI("FXC Contraction")
}
void gauxc_integrator_eval_fxc_contraction_uks(GauXCStatus* status, const GauXCIntegrator, const int64_t,
                                               const int64_t, const double*, const int64_t, const double*,
                                               const int64_t, const double*, const int64_t, const double*,
                                               const int64_t, double*, const int64_t, double*, const int64_t) {
  GAUXC_B200_NYI("FXC Contraction")
}
void gauxc_b200_integrator_eval_exc_vxc_rks_device(GauXCStatus* status, const GauXCIntegrator integrator,
                                                   const double* dP, double* dVXC, double* d_out2) {
  C_TRY(status)
  INTG(integrator)->eval_exc_vxc_device(dP, dVXC, d_out2, true);
  C_CATCH(status)
}
void gauxc_b200_integrator_stats(GauXCStatus* status, const GauXCIntegrator integrator, double* o) {
  C_TRY(status)
  const auto& s = INTG(integrator)->stats();
  o[0] = s.last_local_work_ms; o[1] = s.last_total_ms;
  for (int k = 0; k < 4; ++k) o[2 + k] = s.kernel_ms[k];
  o[6] = (double)s.kernel_launches; o[7] = s.f_dense; o[8] = s.sum_nbe_npts; o[9] = (double)s.npts;
  o[10] = (double)s.ntiles; o[11] = (double)s.nbatches; o[12] = (double)s.nitems; o[13] = s.n_el;
  o[14] = o[15] = 0.;
  C_CATCH(status)
}
void gauxc_b200_integrator_set_vxc_root_only(GauXCStatus* status, const GauXCIntegrator integrator, int on) {
  C_TRY(status)
  INTG(integrator)->set_vxc_root_only(on != 0);
  C_CATCH(status)
}
void gauxc_b200_integrator_set_profile(GauXCStatus* status, const GauXCIntegrator integrator, int on) {
  C_TRY(status)
  INTG(integrator)->set_profile(on != 0);
  C_CATCH(status)
}

// ---- include/gauxc/c/hdf5.h ----------------------------------------------------------------------------
void gauxc_molecule_write_hdf5_record(GauXCStatus* status, GauXCMolecule mol, const char* fname, const char* dset) {
  C_TRY(status)
  write_hdf5_record(*MOL(mol), fname ? fname : "", dset ? dset : "");
  C_CATCH(status)
}
void gauxc_basisset_write_hdf5_record(GauXCStatus* status, GauXCBasisSet basis, const char* fname, const char* dset) {
  C_TRY(status)
  write_hdf5_record(*BAS(basis), fname ? fname : "", dset ? dset : "");
  C_CATCH(status)
}
void gauxc_molecule_read_hdf5_record(GauXCStatus* status, GauXCMolecule mol, const char* fname, const char* dset) {
  C_TRY(status)
  read_hdf5_record(*MOL(mol), fname ? fname : "", dset ? dset : "");
  C_CATCH(status)
}
void gauxc_basisset_read_hdf5_record(GauXCStatus* status, GauXCBasisSet basis, const char* fname, const char* dset) {
  C_TRY(status)
  read_hdf5_record(*BAS(basis), fname ? fname : "", dset ? dset : "");
  C_CATCH(status)
}

// dense FP64 datasets (/DENSITY, /VXC, /EXC ... of the reference's fixtures)
int64_t gauxc_b200_hdf5_dataset_size(GauXCStatus* status, const char* fname, const char* dset, int64_t* dims4, int* rank) {
  int64_t n = 0;
  C_TRY(status)
  std::vector<double> data;
  std::vector<size_t> dims;
  read_hdf5_dataset(fname ? fname : "", dset ? dset : "", data, dims);
  if (dims.size() > 4) GAUXC_GENERIC_EXCEPTION("HDF5: rank > 4");
  if (rank) *rank = (int)dims.size();
  for (size_t i = 0; i < dims.size() && dims4; ++i) dims4[i] = (int64_t)dims[i];
  n = (int64_t)data.size();
  C_CATCH(status)
  return n;
}
void gauxc_b200_hdf5_read_dataset(GauXCStatus* status, const char* fname, const char* dset, double* out, int64_t n) {
  C_TRY(status)
  std::vector<double> data;
  std::vector<size_t> dims;
  read_hdf5_dataset(fname ? fname : "", dset ? dset : "", data, dims);
  if ((int64_t)data.size() != n) GAUXC_GENERIC_EXCEPTION("HDF5: dataset size mismatch");
  std::copy(data.begin(), data.end(), out);
  C_CATCH(status)
}
void gauxc_b200_hdf5_write_dataset(GauXCStatus* status, const char* fname, const char* dset, const double* data,
                                   const int64_t* dims, int rank) {
  C_TRY(status)
  std::vector<size_t> d(dims, dims + rank);
  write_hdf5_dataset(fname ? fname : "", dset ? dset : "", data, d);
  C_CATCH(status)
}
// molecule / basis accessors for records read from a file
void gauxc_b200_molecule_get_atoms(GauXCStatus* status, const GauXCMolecule mol, GauXCAtom* atoms) {
  C_TRY(status)
  auto* m = MOL(mol);
  for (size_t i = 0; i < m->size(); ++i) atoms[i] = GauXCAtom{(*m)[i].Z, (*m)[i].x, (*m)[i].y, (*m)[i].z};
  C_CATCH(status)
}

// ---- NCCL ---------------------------------------------------------------------------------------
void gauxc_b200_nccl_get_unique_id(GauXCStatus* status, char id[128]) {
  C_TRY(status)
  nccl_get_unique_id(id);
  C_CATCH(status)
}
void gauxc_b200_nccl_init(GauXCStatus* status, const char id[128], int rank, int size) {
  C_TRY(status)
  nccl_init_global(id, rank, size);
  C_CATCH(status)
}
void gauxc_b200_nccl_finalize(GauXCStatus* status) {
  C_TRY(status)
  nccl_finalize_global();
  C_CATCH(status)
}
void gauxc_b200_allreduce_device(GauXCStatus* status, double* dptr, size_t n) {
  C_TRY(status)
  device_allreduce(dptr, n);
  C_CATCH(status)
}

// ---- introspection ---------------------------------------------------------------------------------
int64_t gauxc_b200_basisset_nbf(GauXCStatus* status, const GauXCBasisSet basis) {
  int64_t v = 0;
  C_TRY(status)
  v = BAS(basis)->nbf();
  C_CATCH(status)
  return v;
}
int64_t gauxc_b200_basisset_nshells(GauXCStatus* status, const GauXCBasisSet basis) {
  int64_t v = 0;
  C_TRY(status)
  v = BAS(basis)->nshells();
  C_CATCH(status)
  return v;
}
void gauxc_b200_basisset_set_shell_tolerance(GauXCStatus* status, GauXCBasisSet basis, double tol) {
  C_TRY(status)
  for (auto& s : *BAS(basis)) s.set_shell_tolerance(tol);
  C_CATCH(status)
}
void gauxc_b200_basisset_get_shell(GauXCStatus* status, const GauXCBasisSet basis, int64_t s, int32_t* l,
                                   int32_t* pure, int32_t* nprim, double* cutoff, double* origin,
                                   double* alpha, double* coeff) {
  C_TRY(status)
  const auto& sh = BAS(basis)->at((size_t)s);
  *l = sh.l; *pure = sh.pure; *nprim = sh.nprim; *cutoff = sh.cutoff_radius;
  for (int i = 0; i < 3; ++i) origin[i] = sh.O[i];
  for (int i = 0; i < 32; ++i) { alpha[i] = sh.alpha[i]; coeff[i] = sh.coeff[i]; }
  C_CATCH(status)
}
int64_t gauxc_b200_load_balancer_ntasks(GauXCStatus* status, const GauXCLoadBalancer lb) {
  int64_t v = 0;
  C_TRY(status)
  v = (int64_t)(*LB(lb))->get_tasks().size();
  C_CATCH(status)
  return v;
}
int64_t gauxc_b200_load_balancer_total_npts(GauXCStatus* status, const GauXCLoadBalancer lb) {
  int64_t v = 0;
  C_TRY(status)
  v = (int64_t)(*LB(lb))->total_npts();
  C_CATCH(status)
  return v;
}
void gauxc_b200_load_balancer_task_info(GauXCStatus* status, const GauXCLoadBalancer lb, int32_t* iParent,
                                        int32_t* npts, int32_t* nbe, int32_t* nshells, double* dn) {
  C_TRY(status)
  auto& tasks = (*LB(lb))->get_tasks();
  for (size_t i = 0; i < tasks.size(); ++i) {
    iParent[i] = tasks[i].iParent;
    npts[i] = (int32_t)tasks[i].points.size();
    nbe[i] = tasks[i].bfn_screening.nbe;
    nshells[i] = (int32_t)tasks[i].bfn_screening.shell_list.size();
    dn[i] = tasks[i].dist_nearest;
  }
  C_CATCH(status)
}
void gauxc_b200_load_balancer_state(GauXCStatus* status, const GauXCLoadBalancer lb, int* modified_weights_are_stored,
                                    int* weight_alg) {
  C_TRY(status)
  const auto& st = (*LB(lb))->state();
  *modified_weights_are_stored = st.modified_weights_are_stored ? 1 : 0;
  *weight_alg = (int)st.weight_alg;
  C_CATCH(status)
}
void gauxc_b200_load_balancer_get_task(GauXCStatus* status, const GauXCLoadBalancer lb, int64_t it,
                                       double* points, double* weights, int32_t* shell_list) {
  C_TRY(status)
  (*LB(lb))->sync_host_tasks();  // weights modified on the device are fetched on first use
  auto& t = (*LB(lb))->get_tasks().at((size_t)it);
  for (size_t i = 0; i < t.points.size(); ++i) {
    if (points) { points[3 * i] = t.points[i][0]; points[3 * i + 1] = t.points[i][1]; points[3 * i + 2] = t.points[i][2]; }
    if (weights) weights[i] = t.weights[i];
  }
  if (shell_list) std::copy(t.bfn_screening.shell_list.begin(), t.bfn_screening.shell_list.end(), shell_list);
  C_CATCH(status)
}
void gauxc_b200_load_balancer_set_task_weights(GauXCStatus* status, GauXCLoadBalancer lb, int64_t it,
                                               const double* weights) {
  C_TRY(status)
  auto& l = **LB(lb);
  l.sync_host_tasks();
  auto& t = l.get_tasks().at((size_t)it);
  std::copy(weights, weights + t.weights.size(), t.weights.begin());
  l.touch();
  C_CATCH(status)
}
void gauxc_b200_load_balancer_set_tasks(GauXCStatus* status, GauXCLoadBalancer lb, int64_t ntasks,
                                        const int32_t* npts, const int32_t* iParent, const double* dn,
                                        const double* points, const double* weights,
                                        const int32_t* nshells, const int32_t* shell_lists,
                                        int weights_are_modified) {
  C_TRY(status)
  auto& l = **LB(lb);
  std::vector<XCTask> tasks;
  tasks.reserve((size_t)ntasks);
  const int32_t nsh_total = (int32_t)l.basis().size();
  const int32_t natoms = (int32_t)l.molecule().size();
  size_t po = 0, so = 0;
  for (int64_t i = 0; i < ntasks; ++i) {
    if (npts[i] < 0 || nshells[i] < 0) GAUXC_GENERIC_EXCEPTION("Invalid Task: negative size");
    if (iParent[i] < 0 || iParent[i] >= natoms) GAUXC_GENERIC_EXCEPTION("Invalid Task: iParent out of range");
    XCTask t;
    t.iParent = iParent[i];
    t.npts = npts[i];
    t.dist_nearest = dn[i];
    t.points.resize(npts[i]);
    t.weights.assign(weights + po, weights + po + npts[i]);
    for (int p = 0; p < npts[i]; ++p)
      t.points[p] = {points[3 * (po + p)], points[3 * (po + p) + 1], points[3 * (po + p) + 2]};
    t.bfn_screening.shell_list.assign(shell_lists + so, shell_lists + so + nshells[i]);
    int nbe = 0;
    int32_t prev = -1;
    for (int s : t.bfn_screening.shell_list) {
      // the kernels gather the lower triangle of P' / scatter distinct AO pairs: the list must be
      // strictly ascending, as every list the LoadBalancer produces is
      if (s <= prev || s >= nsh_total)
        GAUXC_GENERIC_EXCEPTION("Invalid Task: shell_list must be strictly ascending and within the basis");
      prev = s;
      nbe += l.basis().at(s).size();
    }
    t.bfn_screening.nbe = nbe;
    po += npts[i];
    so += nshells[i];
    tasks.push_back(std::move(t));
  }
  l.replace_tasks(std::move(tasks));
  l.state().modified_weights_are_stored = weights_are_modified != 0;
  l.state().weight_alg = weights_are_modified ? XCWeightAlg::SSF : XCWeightAlg::NOTPARTITIONED;
  C_CATCH(status)
}

int64_t gauxc_b200_lebedev(GauXCStatus* status, int npts, double* xyz, double* w) {
  int64_t n = 0;
  C_TRY(status)
  const auto& r = lebedev_rule(npts);
  n = (int64_t)r.pts.size();
  for (size_t i = 0; i < r.pts.size(); ++i) {
    xyz[3 * i] = r.pts[i][0]; xyz[3 * i + 1] = r.pts[i][1]; xyz[3 * i + 2] = r.pts[i][2];
    w[i] = r.wts[i];
  }
  C_CATCH(status)
  return n;
}
void gauxc_b200_radial(GauXCStatus* status, enum GauXC_RadialQuad rq, int n, double R, double* r, double* w) {
  C_TRY(status)
  std::vector<double> rr, ww;
  radial_quadrature((RadialQuad)rq, n, R, rr, ww);
  std::copy(rr.begin(), rr.end(), r);
  std::copy(ww.begin(), ww.end(), w);
  C_CATCH(status)
}
void gauxc_b200_eval_collocation(GauXCStatus* status, const GauXCBasisSet basis, int64_t nshells,
                                 const int32_t* shell_list, int64_t npts, const double* points,
                                 double* eval, double* dx, double* dy, double* dz) {
  C_TRY(status)
  std::vector<int32_t> sl(shell_list, shell_list + nshells);
  device_eval_collocation(*BAS(basis), sl, npts, points, eval, dx, dy, dz);
  C_CATCH(status)
}
void gauxc_b200_eval_collocation_hessian(GauXCStatus* status, const GauXCBasisSet basis, int64_t nshells,
                                         const int32_t* shell_list, int64_t npts, const double* points, double* eval,
                                         double* dx, double* dy, double* dz, double* hess6) {
  C_TRY(status)
  std::vector<int32_t> sl(shell_list, shell_list + nshells);
  device_eval_collocation(*BAS(basis), sl, npts, points, eval, dx, dy, dz, hess6);
  C_CATCH(status)
}
double gauxc_b200_probe_peak(GauXCStatus* status, int which) {
  double v = 0;
  C_TRY(status)
  v = device_probe_peak(which);
  C_CATCH(status)
  return v;
}
int gauxc_b200_device_count(void) { return device_count(); }
void gauxc_b200_set_device(GauXCStatus* status, int device) {
  C_TRY(status)
  device_set(device);
  C_CATCH(status)
}
void* gauxc_b200_integrator_stream(GauXCStatus* status, const GauXCIntegrator integrator) {
  C_TRY(status)
  return INTG(integrator)->stream();
  C_CATCH(status)
  return nullptr;
}
const char* gauxc_b200_version(void) { return "gauxc_b200 0.1 (sm_100a)"; }

}  // extern "C"
