// Device-side shell screening for the LoadBalancer (definition in lb_screen.cu).
#pragma once
#include <vector>

namespace gxb {

struct LbScreenView {
  const double *atoms, *box_lo, *box_up, *shell_xyz, *shell_rad;
  const int *shell_size, *pair_atom, *pair_box;
  int *nshell_out, *nbe_out;
  const unsigned char* want;
  long long npairs;
  int nshells;
};

// pair = (atom, bounding box of one batch of the atom's grid, relative to the atom centre)
struct LbScreenInput {
  std::vector<double> atoms;               // natoms x 3
  std::vector<double> box_lo, box_up;      // nboxes x 3 (all grid types concatenated)
  std::vector<double> shell_xyz, shell_rad;  // nshells x 3, nshells (cutoff radii)
  std::vector<int> shell_size;             // basis functions per shell
  std::vector<int> pair_atom, pair_box;    // npairs
};

class LbScreen {
  struct State;
  State* st_;

public:
  explicit LbScreen(const LbScreenInput& in);
  ~LbScreen();
  LbScreen(const LbScreen&) = delete;
  // pass 1: number of shells / basis functions that reach each pair's box
  void count(std::vector<int>& nshell, std::vector<int>& nbe);
  // pass 2: the ascending shell lists of the wanted pairs, written at offsets[p] of `lists` (total entries)
  void fill(const std::vector<unsigned char>& want, const std::vector<long long>& offsets, long long total,
            std::vector<int>& lists);
};

}  // namespace gxb
