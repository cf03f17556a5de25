#include "load_balancer.hpp"
#include <algorithm>
#include <numeric>

namespace GauXC {

bool cube_sphere_intersect(const double* lo, const double* up, const double* center, double rad) {
  double dist = rad * rad;
  for (int i = 0; i < 3; ++i) {
    double r = 0.;
    if (center[i] < lo[i]) r = lo[i] - center[i];
    else if (center[i] > up[i]) r = center[i] - up[i];
    dist -= r * r;
    if (dist < 0.) return false;
  }
  return true;
}

LoadBalancer::LoadBalancer(std::shared_ptr<RuntimeEnvironment> rt, const Molecule& mol,
                           const MolGrid& mg, const BasisSet& basis)
    : runtime_(std::move(rt)),
      mol_(std::make_shared<Molecule>(mol)),
      mg_(std::make_shared<MolGrid>(mg)),
      basis_(std::make_shared<BasisSet>(basis)),
      molmeta_(std::make_shared<MolMeta>(mol)),
      basis_map_(std::make_shared<BasisSetMap>(basis, mol)) {}

std::vector<XCTask>& LoadBalancer::get_tasks() {
  if (!tasks_created_) {
    local_tasks_ = create_local_tasks_();
    tasks_created_ = true;
    ++version_;
  }
  return local_tasks_;
}

size_t LoadBalancer::total_npts() {
  size_t n = 0;
  for (auto& t : get_tasks()) n += t.points.size();
  return n;
}
size_t LoadBalancer::max_npts() {
  size_t n = 0;
  for (auto& t : get_tasks()) n = std::max(n, t.points.size());
  return n;
}
size_t LoadBalancer::max_nbe() {
  size_t n = 0;
  for (auto& t : get_tasks()) n = std::max(n, (size_t)t.bfn_screening.nbe);
  return n;
}

// Same pipeline as HostReplicatedLoadBalancer::create_local_tasks_
// (src/load_balancer/host/replicated_host_load_balancer.cxx:22-193): per atom, per batch:
// box/sphere shell screening -> deal batches to ranks greedily by XCTask::cost -> sort by
// (iParent, shell_list) -> merge batches with identical screening into one task.
std::vector<XCTask> LoadBalancer::create_local_tasks_() const {
  const int32_t n_deriv = 1;
  const int32_t world_rank = runtime_->comm_rank();
  const int32_t world_size = runtime_->comm_size();
  const auto& basis = *basis_;
  const auto& mol = *mol_;
  const size_t natoms = mol.size();
  const size_t nsh = basis.size();

  // coarse pre-filter: shells grouped by centre with the largest cutoff of the group
  std::vector<std::vector<int32_t>> center_shells(natoms);
  std::vector<int32_t> loose_shells;
  std::vector<double> center_maxrad(natoms, 0.);
  for (size_t s = 0; s < nsh; ++s) {
    const int c = basis_map_->shell_to_center[s];
    if (c < 0) loose_shells.push_back((int32_t)s);
    else {
      center_shells[c].push_back((int32_t)s);
      center_maxrad[c] = std::max(center_maxrad[c], basis[s].cutoff_radius);
    }
  }

  std::vector<XCTask> local_work;
  std::vector<size_t> global_workload(world_size, 0);

  for (size_t iAtom = 0; iAtom < natoms; ++iAtom) {
    const auto& atom = mol[iAtom];
    const Grid& grid = mg_->get_grid(atom.Z);
    const size_t nb = grid.nbatches();
    std::vector<XCTask> temp(nb);
    std::vector<char> keep(nb, 0);

#pragma omp parallel for schedule(dynamic, 4)
    for (size_t ib = 0; ib < nb; ++ib) {
      const GridBatch& gb = grid.batch(ib);
      if (gb.points.empty()) continue;
      const double lo[3] = {gb.lo[0] + atom.x, gb.lo[1] + atom.y, gb.lo[2] + atom.z};
      const double up[3] = {gb.up[0] + atom.x, gb.up[1] + atom.y, gb.up[2] + atom.z};

      std::vector<int32_t> shell_list;
      for (size_t c = 0; c < natoms; ++c) {
        if (center_shells[c].empty()) continue;
        const double cen[3] = {mol[c].x, mol[c].y, mol[c].z};
        if (!cube_sphere_intersect(lo, up, cen, center_maxrad[c])) continue;
        for (int32_t s : center_shells[c])
          if (cube_sphere_intersect(lo, up, basis[s].O.data(), basis[s].cutoff_radius))
            shell_list.push_back(s);
      }
      for (int32_t s : loose_shells)
        if (cube_sphere_intersect(lo, up, basis[s].O.data(), basis[s].cutoff_radius))
          shell_list.push_back(s);
      if (shell_list.empty()) continue;
      std::sort(shell_list.begin(), shell_list.end());

      size_t nbe = 0;
      for (auto s : shell_list) nbe += basis[s].size();

      XCTask& task = temp[ib];
      task.iParent = (int32_t)iAtom;
      task.npts = (int32_t)gb.points.size();
      task.points.resize(gb.points.size());
      for (size_t i = 0; i < gb.points.size(); ++i)
        task.points[i] = {gb.points[i][0] + atom.x, gb.points[i][1] + atom.y,
                          gb.points[i][2] + atom.z};
      task.weights = gb.weights;
      task.bfn_screening.shell_list = std::move(shell_list);
      task.bfn_screening.nbe = (int32_t)nbe;
      task.dist_nearest = molmeta_->dist_nearest[iAtom];
      keep[ib] = 1;
    }

    // deterministic greedy deal in batch order
    for (size_t ib = 0; ib < nb; ++ib) {
      if (!keep[ib]) continue;
      auto min_it = std::min_element(global_workload.begin(), global_workload.end());
      const int64_t min_rank = std::distance(global_workload.begin(), min_it);
      global_workload[min_rank] += temp[ib].cost(n_deriv, natoms);
      if (world_rank == min_rank) local_work.push_back(std::move(temp[ib]));
    }
  }

  auto task_order = [](const XCTask& a, const XCTask& b) {
    if (a.iParent < b.iParent) return true;
    if (a.iParent > b.iParent) return false;
    return a.bfn_screening.shell_list < b.bfn_screening.shell_list;
  };
  std::stable_sort(local_work.begin(), local_work.end(), task_order);

  // merge runs of equivalent tasks
  std::vector<XCTask> merged;
  for (auto& t : local_work) {
    if (!merged.empty() && merged.back().equiv_with(t)) merged.back().merge_with(t);
    else merged.push_back(std::move(t));
  }
  return merged;
}

}  // namespace GauXC
