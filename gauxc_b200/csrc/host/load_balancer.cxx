#include "load_balancer.hpp"
#include "../cuda/lb_screen.hpp"
#include <algorithm>
#include <cmath>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <map>
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
                           const MolGrid& mg, const BasisSet& basis, const std::string& kernel, ExecutionSpace ex)
    : runtime_(std::move(rt)),
      mol_(std::make_shared<Molecule>(mol)),
      mg_(std::make_shared<MolGrid>(mg)),
      basis_(std::make_shared<BasisSet>(basis)),
      molmeta_(std::make_shared<MolMeta>(mol)),
      basis_map_(std::make_shared<BasisSetMap>(basis, mol)) {
  // src/load_balancer/host/load_balancer_host_factory.cxx:28-40
  // the fill-in variant changes nbe (hence the deal over the ranks) after the screening: it keeps the host path
  device_screen_ = ex == ExecutionSpace::Device && kernel != "REPLICATED-FILLIN";
  if (kernel == "REPLICATED-FILLIN") fill_in_ = true;
  else if (kernel != "DEFAULT" && kernel != "REPLICATED" && kernel != "REPLICATED-PETITE")
    GAUXC_GENERIC_EXCEPTION("LoadBalancer Kernel Not Recognized: " + kernel);
}

void LoadBalancer::replace_tasks(std::vector<XCTask> tasks) {
  host_sync_ = nullptr;
  local_tasks_ = std::move(tasks);
  tasks_created_ = true;
  ++version_;
}

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

  const bool dbg_t = std::getenv("GAUXC_B200_LB_TIMING") != nullptr;
  auto tnow = [] { return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count(); };
  double t_scr = 0, t_deal = 0; const double t_begin = tnow();
  std::vector<XCTask> local_work;
  {
    size_t nbatch_total = 0;
    for (size_t a = 0; a < natoms; ++a) nbatch_total += mg_->get_grid(mol[a].Z).nbatches();
    local_work.reserve(nbatch_total / (size_t)world_size + 64);  // grows if the deal is uneven
  }
  std::vector<size_t> global_workload(world_size, 0);

  // Uniform cell list over the shell centres: a batch only tests the centres whose cell can reach its
  // box (the reference tests every shell against every batch, O(natoms) per batch -- 2.5e9 tests for
  // the 2499-atom water cluster).  Candidates are visited in ascending centre index, so the shell
  // lists come out exactly as before.
  double cmin[3] = {1e300, 1e300, 1e300}, cmax[3] = {-1e300, -1e300, -1e300}, rmax_all = 0.;
  for (size_t c = 0; c < natoms; ++c) {
    const double cen[3] = {mol[c].x, mol[c].y, mol[c].z};
    for (int d = 0; d < 3; ++d) { cmin[d] = std::min(cmin[d], cen[d]); cmax[d] = std::max(cmax[d], cen[d]); }
    rmax_all = std::max(rmax_all, center_maxrad[c]);
  }
  // cell edge: at least half the largest cutoff; enlarged until the grid holds at most ~8 cells per atom,
  // so far-apart fragments (dissociation scans) cannot blow the cell array up
  double cell = std::max(4.0, 0.5 * rmax_all);
  int ncell[3] = {1, 1, 1};
  for (;;) {
    double total = 1.;
    for (int d = 0; d < 3; ++d) {
      const double nd = natoms ? std::max(1., std::floor((cmax[d] - cmin[d]) / cell) + 1.) : 1.;
      total *= nd;
      ncell[d] = (int)std::min(nd, 1e6);
    }
    if (total <= 8. * double(natoms) + 64.) break;
    cell *= 1.5;
  }
  std::vector<std::vector<int32_t>> cells((size_t)ncell[0] * ncell[1] * ncell[2]);
  auto cell_of = [&](double v, int d) {
    return std::min(ncell[d] - 1, std::max(0, (int)std::floor((v - cmin[d]) / cell)));
  };
  for (size_t c = 0; c < natoms; ++c)
    cells[((size_t)cell_of(mol[c].x, 0) * ncell[1] + cell_of(mol[c].y, 1)) * ncell[2] + cell_of(mol[c].z, 2)]
        .push_back((int32_t)c);

  // ---- ExecutionSpace::Device: the box/sphere tests and the shell-list compaction run on the GPU
  //      (cuda/lb_screen.cu; reference: replicated_cuda_load_balancer.cxx:71-323); the deal over the ranks, the
  //      sort and the merge below are shared with the Host path, so both produce the same task list bit for bit.
  std::vector<long long> dev_off;       // per atom: first pair; per pair: offset into dev_lists
  std::vector<long long> dev_atom_first(natoms + 1, 0);
  std::vector<int> dev_nshell, dev_nbe, dev_lists;
  std::vector<unsigned char> want;
  if (device_screen_) {
    const double t_a = tnow();
    gxb::LbScreenInput in;
    std::map<int64_t, int> box_first;  // grid type -> first box
    for (size_t a = 0; a < natoms; ++a) {
      in.atoms.insert(in.atoms.end(), {mol[a].x, mol[a].y, mol[a].z});
      const Grid& grid = mg_->get_grid(mol[a].Z);
      if (!box_first.count(mol[a].Z)) {
        box_first[mol[a].Z] = (int)(in.box_lo.size() / 3);
        for (size_t ib = 0; ib < grid.nbatches(); ++ib) {
          const auto& gb = grid.batch(ib);
          in.box_lo.insert(in.box_lo.end(), gb.lo.begin(), gb.lo.end());
          in.box_up.insert(in.box_up.end(), gb.up.begin(), gb.up.end());
        }
      }
      dev_atom_first[a] = (long long)in.pair_atom.size();
      const int b0 = box_first[mol[a].Z];
      for (size_t ib = 0; ib < grid.nbatches(); ++ib) {
        in.pair_atom.push_back((int)a);
        in.pair_box.push_back(b0 + (int)ib);
      }
    }
    dev_atom_first[natoms] = (long long)in.pair_atom.size();
    for (size_t sidx = 0; sidx < nsh; ++sidx) {
      in.shell_xyz.insert(in.shell_xyz.end(), basis[sidx].O.begin(), basis[sidx].O.end());
      in.shell_rad.push_back(basis[sidx].cutoff_radius);
      in.shell_size.push_back(basis[sidx].size());
    }
    std::unique_ptr<gxb::LbScreen> scr_p;
    try {
      scr_p.reset(new gxb::LbScreen(in));
    } catch (const std::exception& e) {
      GAUXC_GENERIC_EXCEPTION(std::string("No CUDA device: the Device LoadBalancer has no CPU fallback in this build (") +
                              e.what() + ")");
    }
    gxb::LbScreen& scr = *scr_p;
    scr.count(dev_nshell, dev_nbe);
    // the greedy deal below needs nbe of EVERY batch but the shell lists only of this rank's: replay the deal
    // on the counts to find them
    const size_t np = in.pair_atom.size();
    want.assign(np, 0);
    {
      std::vector<size_t> wl(world_size, 0);
      for (size_t a = 0; a < natoms; ++a) {
        const Grid& grid = mg_->get_grid(mol[a].Z);
        for (size_t ib = 0; ib < grid.nbatches(); ++ib) {
          const size_t p = (size_t)dev_atom_first[a] + ib;
          if (grid.batch(ib).points.empty() || dev_nshell[p] == 0) continue;
          auto min_it = std::min_element(wl.begin(), wl.end());
          XCTask probe;
          probe.npts = (int32_t)grid.batch(ib).points.size();
          probe.bfn_screening.nbe = dev_nbe[p];
          *min_it += probe.cost(n_deriv, natoms);
          if ((int32_t)std::distance(wl.begin(), min_it) == world_rank) want[p] = 1;
        }
      }
    }
    dev_off.assign(np, 0);
    long long total = 0;
    for (size_t p = 0; p < np; ++p) {
      dev_off[p] = total;
      if (want[p]) total += dev_nshell[p];
    }
    scr.fill(want, dev_off, total, dev_lists);
    t_scr += tnow() - t_a;
  }

  for (size_t iAtom = 0; iAtom < natoms; ++iAtom) {
    const auto& atom = mol[iAtom];
    const Grid& grid = mg_->get_grid(atom.Z);
    const size_t nb = grid.nbatches();
    std::vector<XCTask> temp(nb);
    std::vector<char> keep(nb, 0);
    std::vector<size_t> cost_of(nb, 0);

    const double t_a = tnow();
#pragma omp parallel for schedule(dynamic, 4)
    for (size_t ib = 0; ib < nb; ++ib) {
      const GridBatch& gb = grid.batch(ib);
      if (gb.points.empty()) continue;
      if (device_screen_) {
        // lists of batches dealt to other ranks were not fetched: those tasks only carry their cost
        const size_t p = (size_t)dev_atom_first[iAtom] + ib;
        if (dev_nshell[p] == 0) continue;
        XCTask& task = temp[ib];
        task.iParent = (int32_t)iAtom;
        task.npts = (int32_t)gb.points.size();
        task.bfn_screening.nbe = dev_nbe[p];
        task.dist_nearest = molmeta_->dist_nearest[iAtom];
        keep[ib] = 1;
        cost_of[ib] = task.cost(n_deriv, natoms);
        if (want[p]) {  // this rank's batch (the deal was replayed on the counts): materialise it
          task.points.resize(gb.points.size());
          for (size_t i = 0; i < gb.points.size(); ++i)
            task.points[i] = {gb.points[i][0] + atom.x, gb.points[i][1] + atom.y, gb.points[i][2] + atom.z};
          task.weights = gb.weights;
          task.bfn_screening.shell_list.assign(dev_lists.begin() + dev_off[p],
                                               dev_lists.begin() + dev_off[p] + dev_nshell[p]);
        }
        continue;
      }
      const double lo[3] = {gb.lo[0] + atom.x, gb.lo[1] + atom.y, gb.lo[2] + atom.z};
      const double up[3] = {gb.up[0] + atom.x, gb.up[1] + atom.y, gb.up[2] + atom.z};

      std::vector<int32_t> shell_list;
      std::vector<int32_t> cand;
      {
        int c0[3], c1[3];
        for (int d = 0; d < 3; ++d) {
          c0[d] = cell_of(lo[d] - rmax_all, d);
          c1[d] = cell_of(up[d] + rmax_all, d);
        }
        for (int ix = c0[0]; ix <= c1[0]; ++ix)
          for (int iy = c0[1]; iy <= c1[1]; ++iy)
            for (int iz = c0[2]; iz <= c1[2]; ++iz) {
              const auto& cl = cells[((size_t)ix * ncell[1] + iy) * ncell[2] + iz];
              cand.insert(cand.end(), cl.begin(), cl.end());
            }
        std::sort(cand.begin(), cand.end());
      }
      for (int32_t c : cand) {
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
      if (fill_in_) {
        // fillin_replicated_load_balancer.cxx: every shell between the first and the last hit
        const int32_t first = shell_list.front(), last = shell_list.back();
        shell_list.resize((size_t)(last - first + 1));
        std::iota(shell_list.begin(), shell_list.end(), first);
      }

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
      cost_of[ib] = task.cost(n_deriv, natoms);
    }

    const double t_b = tnow(); t_scr += t_b - t_a;
    // deterministic greedy deal in batch order.  The serial part only reads two flat arrays (cost, keep): the tasks
    // themselves were just written by the worker threads, and pulling 1.5e6 of them through the master's cache one by
    // one cost ~1 us each on a multi-chiplet host ((H2O)833: 1.5-2.3 s).  The owned tasks are then moved to their slots
    // by all threads.
    std::vector<size_t> slot(nb, (size_t)-1);
    size_t n_mine = 0;
    for (size_t ib = 0; ib < nb; ++ib) {
      if (!keep[ib]) continue;
      auto min_it = std::min_element(global_workload.begin(), global_workload.end());
      const int64_t min_rank = std::distance(global_workload.begin(), min_it);
      global_workload[min_rank] += cost_of[ib];
      if (world_rank == min_rank) {
        if (device_screen_ && !want[(size_t)dev_atom_first[iAtom] + ib])
          GAUXC_GENERIC_EXCEPTION("Device LoadBalancer: replayed deal disagrees with the deal");
        slot[ib] = n_mine++;
      }
    }
    const size_t base = local_work.size();
    local_work.resize(base + n_mine);
#pragma omp parallel for schedule(static)
    for (size_t ib = 0; ib < nb; ++ib)
      if (slot[ib] != (size_t)-1) local_work[base + slot[ib]] = std::move(temp[ib]);
    t_deal += tnow() - t_b;
  }
  const double t_loop = tnow();

  // Sort by (iParent, shell_list), stable, then merge runs of equivalent tasks (same parent, same shell
  // list).  Tasks are created parent by parent, so both steps act inside each parent's stretch of the
  // list: the stretches are processed independently and in parallel, then concatenated in order.
  std::vector<XCTask> merged;
  {
    std::vector<size_t> run_begin;
    for (size_t i = 0; i < local_work.size(); ++i)
      if (i == 0 || local_work[i].iParent != local_work[i - 1].iParent) run_begin.push_back(i);
    run_begin.push_back(local_work.size());
    bool ascending = true;
    for (size_t r = 0; r + 2 < run_begin.size(); ++r)
      ascending = ascending && local_work[run_begin[r]].iParent < local_work[run_begin[r + 1]].iParent;
    auto merge_range = [](std::vector<XCTask>& out, std::vector<XCTask>::iterator b, std::vector<XCTask>::iterator e) {
      for (auto it = b; it != e; ++it) {
        if (!out.empty() && out.back().equiv_with(*it)) out.back().merge_with(*it);
        else out.push_back(std::move(*it));
      }
    };
    if (ascending && !local_work.empty()) {
      auto by_list = [](const XCTask& a, const XCTask& b) {
        return a.bfn_screening.shell_list < b.bfn_screening.shell_list;
      };
      const long nruns = (long)run_begin.size() - 1;
      std::vector<std::vector<XCTask>> per_run(nruns);
#pragma omp parallel for schedule(dynamic, 1)
      for (long r = 0; r < nruns; ++r) {
        std::stable_sort(local_work.begin() + run_begin[r], local_work.begin() + run_begin[r + 1], by_list);
        merge_range(per_run[r], local_work.begin() + run_begin[r], local_work.begin() + run_begin[r + 1]);
      }
      size_t total = 0;
      for (auto& v : per_run) total += v.size();
      merged.reserve(total);
      for (auto& v : per_run)
        for (auto& t : v) merged.push_back(std::move(t));
    } else {
      auto task_order = [](const XCTask& a, const XCTask& b) {
        if (a.iParent < b.iParent) return true;
        if (a.iParent > b.iParent) return false;
        return a.bfn_screening.shell_list < b.bfn_screening.shell_list;
      };
      std::stable_sort(local_work.begin(), local_work.end(), task_order);
      merge_range(merged, local_work.begin(), local_work.end());
    }
  }
  const double t_sort = tnow();
  if (dbg_t) std::fprintf(stderr, "[lb] screen %.2f deal %.2f (loop %.2f) sort+merge %.2f s\n", t_scr, t_deal, t_loop - t_begin, t_sort - t_loop);
  return merged;
}

}  // namespace GauXC
