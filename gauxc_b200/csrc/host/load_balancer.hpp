// RuntimeEnvironment + LoadBalancer (host task creation).
//   include/gauxc/runtime_environment/decl.hpp:26-86
//   include/gauxc/load_balancer.hpp:71-119
//   src/load_balancer/host/replicated_host_load_balancer.cxx:22-193
//   src/load_balancer/host/petite_replicated_load_balancer.cxx:31-65
#pragma once
#include "grid.hpp"
#include "types.hpp"
#include <functional>
#include <memory>
#include <string>

namespace GauXC {

// There is no MPI in this build: rank/size are supplied by the launcher (torchrun's
// RANK/WORLD_SIZE through the C ABI) instead of an MPI communicator.
class RuntimeEnvironment {
protected:
  int rank_ = 0, size_ = 1;

public:
  RuntimeEnvironment() = default;
  RuntimeEnvironment(int rank, int size) : rank_(rank), size_(size) {}
  virtual ~RuntimeEnvironment() = default;
  int comm_rank() const { return rank_; }
  int comm_size() const { return size_; }
  void set_comm(int rank, int size) {
    if (size < 1 || rank < 0 || rank >= size) GAUXC_GENERIC_EXCEPTION("Invalid Rank/Size");
    rank_ = rank;
    size_ = size;
  }
};

class DeviceRuntimeEnvironment : public RuntimeEnvironment {
  double fill_fraction_ = 0.9;
  void* user_mem_ = nullptr;
  size_t user_mem_sz_ = 0;

public:
  explicit DeviceRuntimeEnvironment(double fill_fraction);
  DeviceRuntimeEnvironment(void* mem, size_t sz);
  double fill_fraction() const { return fill_fraction_; }
  void* device_memory() const { return user_mem_; }
  size_t device_memory_size() const { return user_mem_sz_; }
};

struct LoadBalancerState {
  bool modified_weights_are_stored = false;
  XCWeightAlg weight_alg = XCWeightAlg::NOTPARTITIONED;
};

class LoadBalancer {
  std::shared_ptr<RuntimeEnvironment> runtime_;
  std::shared_ptr<Molecule> mol_;
  std::shared_ptr<MolGrid> mg_;
  std::shared_ptr<BasisSet> basis_;
  std::shared_ptr<MolMeta> molmeta_;
  std::shared_ptr<BasisSetMap> basis_map_;
  std::vector<XCTask> local_tasks_;
  bool tasks_created_ = false;
  LoadBalancerState state_;
  uint64_t version_ = 0;  // bumped whenever tasks / weights change (device caches key on it)

  // pending refresh of host-side task data that currently lives on the device (SSF weights)
  std::function<void(std::vector<XCTask>&)> host_sync_;
  bool fill_in_ = false;
  bool device_screen_ = false;  // ExecutionSpace::Device: shell screening on the GPU (cuda/lb_screen.cu)  // REPLICATED-FILLIN: contiguous shell range first..last instead of the exact list

  std::vector<XCTask> create_local_tasks_() const;

public:
  // kernel: "DEFAULT" / "REPLICATED" / "REPLICATED-PETITE" (exact shell lists,
  // petite_replicated_load_balancer.cxx:31-65) or "REPLICATED-FILLIN" (fillin_replicated_load_balancer.cxx)
  LoadBalancer(std::shared_ptr<RuntimeEnvironment> rt, const Molecule& mol, const MolGrid& mg,
               const BasisSet& basis, const std::string& kernel = "DEFAULT",
               ExecutionSpace ex = ExecutionSpace::Host);

  std::vector<XCTask>& get_tasks();
  // install a user supplied task list without generating the default one first
  void replace_tasks(std::vector<XCTask> tasks);
  bool tasks_created() const { return tasks_created_; }
  // Device MolecularWeights leave the partitioned weights on the device; the host copy is refreshed on demand
  void set_host_sync(std::function<void(std::vector<XCTask>&)> f) { host_sync_ = std::move(f); }
  void sync_host_tasks() {
    if (host_sync_) {
      auto f = std::move(host_sync_);
      host_sync_ = nullptr;
      f(local_tasks_);
    }
  }
  const Molecule& molecule() const { return *mol_; }
  const MolGrid& molgrid() const { return *mg_; }
  const BasisSet& basis() const { return *basis_; }
  const MolMeta& molmeta() const { return *molmeta_; }
  const BasisSetMap& basis_map() const { return *basis_map_; }
  const RuntimeEnvironment& runtime() const { return *runtime_; }
  std::shared_ptr<RuntimeEnvironment> runtime_ptr() const { return runtime_; }
  LoadBalancerState& state() { return state_; }
  uint64_t version() const { return version_; }
  void touch() { ++version_; }

  // device-resident copy of the task data (owned by the Device execution space objects);
  // valid while device_cache_version == version()
  std::shared_ptr<void> device_cache;
  uint64_t device_cache_version = ~0ull;

  size_t total_npts();
  size_t max_npts();
  size_t max_nbe();
};

// cube/sphere test of include/gauxc/util/geometry.hpp:36-54
bool cube_sphere_intersect(const double* lo, const double* up, const double* center, double rad);

}  // namespace GauXC
