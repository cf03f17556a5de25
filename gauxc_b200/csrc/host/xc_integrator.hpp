// Host-side objects of the Device execution space: functional descriptor, reduction driver,
// molecular weights, XC integrator.  Names and call contracts mirror the reference façade:
//   XCIntegrator / ReplicatedXCIntegratorImpl   include/gauxc/xc_integrator.hpp:29-94,
//       include/gauxc/xc_integrator/replicated/replicated_xc_integrator_impl.hpp:33-196
//   IncoreReplicatedXCDeviceIntegrator::eval_exc_vxc_
//       src/xc_integrator/replicated/device/incore_replicated_xc_device_integrator_exc_vxc.hpp:46-386
//   MolecularWeights   include/gauxc/molecular_weights.hpp:24-115,
//       src/molecular_weights/device/device_molecular_weights.cxx:18-88
//   ReductionDriver    include/gauxc/reduction_driver.hpp:26-70,
//       src/reduction_driver/device/nccl_reduction_driver.cxx:58-103
#pragma once
#include "../cuda/device_plan.hpp"
#include "load_balancer.hpp"
#include <map>
#include <memory>
#include <string>

namespace GauXC {

// ---- functional (stands in for ExchCXX::XCFunctional) ---------------------------------
struct XCFunctional {
  std::string name;
  gxb::FunctionalDesc desc{};
  bool polarized = false;
  bool is_lda() const { return !desc.is_gga; }
  bool is_gga() const { return desc.is_gga != 0; }
  double hyb_exx = 0.;  // informational (PBE0 = 0.25)
};
XCFunctional functional_from_string(const std::string& spec, bool polarized);

// ---- timer (include/gauxc/util/timer.hpp:42-97) ----------------------------------------
struct Timer {
  std::map<std::string, double> ms;
  void add(const std::string& k, double v) { ms[k] += v; }
};

// ---- reduction driver ------------------------------------------------------------------
enum class ReductionOp { Sum };
class ReductionDriver {
public:
  virtual ~ReductionDriver() = default;
  virtual bool takes_host_memory() const = 0;
  virtual bool takes_device_memory() const = 0;
  // stream: cudaStream_t passed as void* (the reference passes std::any queue)
  virtual void allreduce_inplace(double* data, size_t n, ReductionOp op, void* stream) = 0;
  virtual int comm_size() const = 0;
  // extension used for the replicated density upload: rank r contributes data[r * count .. (r+1) * count)
  virtual bool can_allgather() const { return false; }
  virtual void allgather_inplace(double*, size_t /*count_per_rank*/, void* /*stream*/) {
    GAUXC_GENERIC_EXCEPTION("allgather NYI for this ReductionDriver");
  }
};
// "Default"/"NCCL": NCCL over NVLink when comm_size > 1, a no-op driver for a single rank.
// "BasicMPI" is rejected (no MPI in this build).
std::shared_ptr<ReductionDriver> make_reduction_driver(const RuntimeEnvironment& rt,
                                                       const std::string& name);
// NCCL bootstrap without MPI: rank 0 creates the id, the launcher broadcasts the bytes
// (torch.distributed store / file), every rank then joins.
void nccl_get_unique_id(char out[128]);
void nccl_init_global(const char id[128], int rank, int size);
void nccl_finalize_global();

// ---- device-resident task data ----------------------------------------------------------
struct DevicePlan;  // defined in device_integrator.cu
std::shared_ptr<DevicePlan> get_device_plan(LoadBalancer& lb);

// ---- molecular weights ------------------------------------------------------------------
struct MolecularWeightsSettings {
  XCWeightAlg weight_alg = XCWeightAlg::SSF;
  bool becke_size_adjustment = false;
};
class MolecularWeights {
  ExecutionSpace ex_;
  MolecularWeightsSettings settings_;
  Timer timer_;

public:
  MolecularWeights(ExecutionSpace ex, const std::string& lwd_kernel, MolecularWeightsSettings s);
  void modify_weights(LoadBalancer& lb);
  const Timer& get_timings() const { return timer_; }
};

// ---- XC integrator ------------------------------------------------------------------------
struct XCIntegratorStats {
  double last_local_work_ms = 0.;   // device time of the batch loop (CUDA events)
  double last_total_ms = 0.;        // device time incl. H2D/D2H and reduction
  double kernel_ms[4] = {0, 0, 0, 0};  // collocation, xmat+density, func+zmat, vxc (profiled mode)
  long long kernel_launches = 0;
  double f_dense = 0.;              // sum_t 4 nbe^2 npts
  double sum_nbe_npts = 0.;
  long long npts = 0;
  long long ntiles = 0, nbatches = 0, nitems = 0;
  double n_el = 0.;
};

class XCIntegrator {
  std::shared_ptr<XCFunctional> func_;
  std::shared_ptr<LoadBalancer> lb_;
  std::shared_ptr<ReductionDriver> red_;
  struct Impl;
  std::shared_ptr<Impl> impl_;
  Timer timer_;
  XCIntegratorStats stats_;
  bool vxc_root_only_ = false;

  void reduce_and_symmetrize_(double* dV, double* dVz, double* d_out2, int nbf, bool do_vxc);
  void upload_density_(const double* P, int64_t ldp, double* dP, size_t nbf);
  void eval_exc_grad_(int64_t m, int64_t n, const double* P, int64_t ldp, const double* Pz, int64_t ldpz,
                      double* EXC_GRAD, bool include_weight_derivatives);
  void eval_uks_(int64_t m, int64_t n, const double* Ps, int64_t ldps, const double* Pz, int64_t ldpz,
                 double* VXCs, int64_t ldvxcs, double* VXCz, int64_t ldvxcz, double* EXC, bool do_vxc);

public:
  XCIntegrator(ExecutionSpace ex, const std::string& input_type, const std::string& integrator_kernel,
               const std::string& lwd_kernel, const std::string& reduction_kernel,
               std::shared_ptr<XCFunctional> func, std::shared_ptr<LoadBalancer> lb);
  ~XCIntegrator();

  // RKS: P = P_alpha; VXC fully overwritten (column-major, ldvxc >= nbf)
  void eval_exc_vxc(int64_t m, int64_t n, const double* P, int64_t ldp, double* VXC,
                    int64_t ldvxc, double* EXC);
  // UKS (LDA functionals): Ps = P_alpha + P_beta, Pz = P_alpha - P_beta; VXCs / VXCz fully overwritten
  void eval_exc_vxc_uks(int64_t m, int64_t n, const double* Ps, int64_t ldps, const double* Pz, int64_t ldpz,
                        double* VXCs, int64_t ldvxcs, double* VXCz, int64_t ldvxcz, double* EXC);
  void eval_exc_uks(int64_t m, int64_t n, const double* Ps, int64_t ldps, const double* Pz, int64_t ldpz,
                    double* EXC);
  void eval_exc(int64_t m, int64_t n, const double* P, int64_t ldp, double* EXC);
  // EXC gradient w.r.t. the nuclear coordinates (3 * natoms, atom-major), RKS.  include_weight_derivatives:
  // IntegratorSettingsEXC_GRAD (include/gauxc/xc_integrator_settings.hpp:28-30), default true = full gradient with
  // the grid-weight contribution and translational invariance; false = Hellmann-Feynman-like gradient
  void eval_exc_grad(int64_t m, int64_t n, const double* P, int64_t ldp, double* EXC_GRAD,
                     bool include_weight_derivatives = true);
  void eval_exc_grad_uks(int64_t m, int64_t n, const double* Ps, int64_t ldps, const double* Pz, int64_t ldpz,
                         double* EXC_GRAD, bool include_weight_derivatives = true);
  void integrate_den(int64_t m, int64_t n, const double* P, int64_t ldp, double* N_EL);
  // device-resident variant: dP (nbf x nbf, ld nbf) and dVXC live in HBM, out2 = {EXC, N_EL}
  // device scalars; no host<->device traffic in the call.
  void eval_exc_vxc_device(const double* dP, double* dVXC, double* d_out2, bool do_vxc = true);

  void set_profile(bool on);
  // Extension: only rank 0 copies VXC back to its host buffer (the other ranks' buffers are left untouched).  The
  // replicated contract of the reference makes every rank of a box pull nbf^2 * 8 bytes through the host's PCIe
  // root / memory controllers at once; a caller that diagonalises on one rank does not need that.
  void set_vxc_root_only(bool on) { vxc_root_only_ = on; }
  const XCIntegratorStats& stats() const { return stats_; }
  const Timer& get_timings() const { return timer_; }
  LoadBalancer& load_balancer() { return *lb_; }
  void* stream() const;
};

}  // namespace GauXC
