// Atomic / molecular quadrature grids and the spatial micro-batcher.
// Replaces what the reference obtains from IntegratorXX (un-vendored dependency):
//   src/grid_factory.cxx:35-247, src/grid_impl.cxx:32-36, src/molgrid_defaults.cxx:52-199
#pragma once
#include "types.hpp"
#include <map>
#include <memory>

namespace GauXC {

// ---- Lebedev-Laikov rules (orbit tables generated from scipy, see tools/gen_lebedev.py)
struct LebedevRule {
  std::vector<std::array<double, 3>> pts;
  std::vector<double> wts;  // sum = 4*pi
};
const LebedevRule& lebedev_rule(int npts);
int lebedev_algebraic_order_by_npts(int npts);   // -1 if unknown
int lebedev_npts_by_algebraic_order(int order);  // -1 if unknown
int lebedev_next_algebraic_order(int order);     // smallest tabulated order >= order

// ---- grid specifications (include/gauxc/grid_factory.hpp)
struct PruningRegion {
  size_t idx_st, idx_en;
  int angular_size;
};
struct UnprunedAtomicGridSpecification {
  RadialQuad radial_quad;
  int radial_size;
  double radial_scale;
  int angular_size;
};
struct PrunedAtomicGridSpecification {
  RadialQuad radial_quad;
  int radial_size;
  double radial_scale;
  std::vector<PruningRegion> pruning_regions;
};

double default_radial_scaling_factor(RadialQuad rq, int64_t Z);
std::pair<int, int> default_grid_size(int64_t Z, RadialQuad rq, AtomicGridSizeDefault s);
PrunedAtomicGridSpecification create_pruned_spec(PruningScheme, UnprunedAtomicGridSpecification);

// radial nodes r_i (ascending index i = 0..n-1) and weights including r^2 Jacobian
void radial_quadrature(RadialQuad rq, int n, double R, std::vector<double>& r,
                       std::vector<double>& w);

// ---- one atom-centred grid, already cut into spatial batches ------------------
struct GridBatch {
  std::array<double, 3> lo, up;  // bounding box (relative to the atom centre)
  std::vector<std::array<double, 3>> points;
  std::vector<double> weights;
};

class Grid {
  std::vector<GridBatch> batches_;
  size_t npts_ = 0;
  int64_t max_batch_sz_ = 512;

public:
  Grid() = default;
  Grid(const PrunedAtomicGridSpecification& spec, int64_t max_batch_sz);
  size_t nbatches() const { return batches_.size(); }
  size_t npts() const { return npts_; }
  int64_t max_batch_sz() const { return max_batch_sz_; }
  const GridBatch& batch(size_t i) const { return batches_[i]; }
};

// ---- MolGrid (include/gauxc/molgrid.hpp): Z -> Grid
class MolGrid {
  std::map<int64_t, std::shared_ptr<Grid>> grids_;

public:
  MolGrid() = default;
  explicit MolGrid(std::map<int64_t, std::shared_ptr<Grid>> g) : grids_(std::move(g)) {}
  size_t natoms_uniq() const { return grids_.size(); }
  const Grid& get_grid(int64_t Z) const {
    auto it = grids_.find(Z);
    if (it == grids_.end()) GAUXC_GENERIC_EXCEPTION("No Grid For Atomic Number");
    return *it->second;
  }
  size_t max_nbatches() const {
    size_t n = 0;
    for (auto& g : grids_) n = std::max(n, g.second->nbatches());
    return n;
  }
};

// MolGridFactory::create_default_molgrid (src/molgrid_defaults.cxx:201+, tests/standalone_driver.cxx:206)
MolGrid create_default_molgrid(const Molecule& mol, PruningScheme scheme, int64_t batch_size,
                               RadialQuad rq, AtomicGridSizeDefault size);

}  // namespace GauXC
