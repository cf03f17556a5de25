// Shell screening of the LoadBalancer on the device (ExecutionSpace::Device of LoadBalancerFactory).
//
// Replaces the reference's device load balancer kernels (src/load_balancer/device/cuda/
// cuda_collision_detection.cu:74-203: bitvector per batch + popcount + position list) and reproduces, bit for bit,
// the host screening (petite_replicated_load_balancer.cxx:31-65 over geometry.hpp:36-54): a shell belongs to a grid
// batch iff the sphere (shell centre, cutoff radius) meets the batch's bounding box.  One warp owns one (atom,
// batch) pair: the 32 lanes test 32 shells at a time and append the hits through a ballot, so every list comes
// out ascending -- the order the host produces -- with no sort.  Two passes (count, then fill at offsets from an
// exclusive scan); the arithmetic of cube_sphere_intersect is kept contraction-free (__dmul_rn / __dsub_rn) so
// that borderline spheres fall on the same side as on the host.
#include "lb_screen.hpp"
#include <cuda_runtime.h>
#include <stdexcept>
#include <string>

namespace gxb {

namespace {

#define LB_CUDA_CHECK(expr)                                                                       \
  do {                                                                                            \
    cudaError_t _e = (expr);                                                                      \
    if (_e != cudaSuccess) throw std::runtime_error(std::string("CUDA Failed: ") + cudaGetErrorString(_e) + " in " #expr); \
  } while (0)

// geometry.hpp:36-54 (same operation order as the host's cube_sphere_intersect; no FMA contraction)
__device__ __forceinline__ bool cube_sphere(const double (&lo)[3], const double (&up)[3], double cx, double cy,
                                            double cz, double rad) {
  double dist = __dmul_rn(rad, rad);
  const double c[3] = {cx, cy, cz};
#pragma unroll
  for (int i = 0; i < 3; ++i) {
    double r = 0.;
    if (c[i] < lo[i]) r = __dsub_rn(lo[i], c[i]);
    else if (c[i] > up[i]) r = __dsub_rn(c[i], up[i]);
    dist = __dsub_rn(dist, __dmul_rn(r, r));
    if (dist < 0.) return false;
  }
  return true;
}

// pair p -> (atom, batch of the atom's grid); boxes are stored per grid type relative to the atom centre
template <bool FILL>
__global__ void __launch_bounds__(256) screen_kernel(LbScreenView v, const long long* __restrict__ offsets,
                                                     int* __restrict__ lists) {
  const int warp_in_block = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const long long p = (long long)blockIdx.x * (blockDim.x >> 5) + warp_in_block;
  if (p >= v.npairs) return;
  if (FILL && v.want && !v.want[p]) return;
  const int a = v.pair_atom[p];
  const int b = v.pair_box[p];
  const double ax = v.atoms[3 * a], ay = v.atoms[3 * a + 1], az = v.atoms[3 * a + 2];
  // same expression as the host: box bound + atom coordinate
  const double lo[3] = {__dadd_rn(v.box_lo[3 * b], ax), __dadd_rn(v.box_lo[3 * b + 1], ay),
                        __dadd_rn(v.box_lo[3 * b + 2], az)};
  const double up[3] = {__dadd_rn(v.box_up[3 * b], ax), __dadd_rn(v.box_up[3 * b + 1], ay),
                        __dadd_rn(v.box_up[3 * b + 2], az)};
  int n = 0, nbe = 0;
  int* out = FILL ? lists + offsets[p] : nullptr;
  for (int s0 = 0; s0 < v.nshells; s0 += 32) {
    const int s = s0 + lane;
    bool hit = false;
    if (s < v.nshells)
      hit = cube_sphere(lo, up, v.shell_xyz[3 * s], v.shell_xyz[3 * s + 1], v.shell_xyz[3 * s + 2], v.shell_rad[s]);
    const unsigned m = __ballot_sync(0xffffffffu, hit);
    if (FILL) {
      if (hit) out[n + __popc(m & ((1u << lane) - 1u))] = s;
    } else {
      int sz = hit ? v.shell_size[s] : 0;
#pragma unroll
      for (int d = 16; d > 0; d >>= 1) sz += __shfl_xor_sync(0xffffffffu, sz, d);
      nbe += sz;
    }
    n += __popc(m);
  }
  if (!FILL && lane == 0) {
    v.nshell_out[p] = n;
    v.nbe_out[p] = nbe;
  }
}

template <typename T>
struct Dev {
  T* p = nullptr;
  explicit Dev(size_t n) { if (n) LB_CUDA_CHECK(cudaMalloc((void**)&p, n * sizeof(T))); }
  Dev(const std::vector<T>& h) : Dev(h.size()) {
    if (!h.empty()) LB_CUDA_CHECK(cudaMemcpy(p, h.data(), h.size() * sizeof(T), cudaMemcpyHostToDevice));
  }
  Dev(const Dev&) = delete;
  ~Dev() { if (p) cudaFree(p); }
};

}  // namespace

struct LbScreen::State {
  Dev<double> atoms, box_lo, box_up, shell_xyz, shell_rad;
  Dev<int> shell_size, pair_atom, pair_box, nshell, nbe;
  long long npairs;
  int nshells;
  State(const LbScreenInput& in)
      : atoms(in.atoms), box_lo(in.box_lo), box_up(in.box_up), shell_xyz(in.shell_xyz), shell_rad(in.shell_rad),
        shell_size(in.shell_size), pair_atom(in.pair_atom), pair_box(in.pair_box), nshell(in.pair_atom.size()),
        nbe(in.pair_atom.size()), npairs((long long)in.pair_atom.size()), nshells((int)in.shell_rad.size()) {}
  LbScreenView view(const unsigned char* want) const {
    LbScreenView v{};
    v.atoms = atoms.p; v.box_lo = box_lo.p; v.box_up = box_up.p; v.shell_xyz = shell_xyz.p; v.shell_rad = shell_rad.p;
    v.shell_size = shell_size.p; v.pair_atom = pair_atom.p; v.pair_box = pair_box.p; v.nshell_out = nshell.p;
    v.nbe_out = nbe.p; v.want = want; v.npairs = npairs; v.nshells = nshells;
    return v;
  }
};

LbScreen::LbScreen(const LbScreenInput& in) : st_(new State(in)) {}
LbScreen::~LbScreen() { delete st_; }

void LbScreen::count(std::vector<int>& nshell, std::vector<int>& nbe) {
  const long long np = st_->npairs;
  nshell.assign((size_t)np, 0);
  nbe.assign((size_t)np, 0);
  if (!np) return;
  const int wpb = 8;
  screen_kernel<false><<<(unsigned)((np + wpb - 1) / wpb), wpb * 32>>>(st_->view(nullptr), nullptr, nullptr);
  LB_CUDA_CHECK(cudaGetLastError());
  LB_CUDA_CHECK(cudaMemcpy(nshell.data(), st_->nshell.p, (size_t)np * sizeof(int), cudaMemcpyDeviceToHost));
  LB_CUDA_CHECK(cudaMemcpy(nbe.data(), st_->nbe.p, (size_t)np * sizeof(int), cudaMemcpyDeviceToHost));
}

void LbScreen::fill(const std::vector<unsigned char>& want, const std::vector<long long>& offsets, long long total,
                    std::vector<int>& lists) {
  const long long np = st_->npairs;
  lists.assign((size_t)total, 0);
  if (!np || !total) return;
  Dev<unsigned char> d_want(want);
  Dev<long long> d_off(offsets);
  Dev<int> d_lists((size_t)total);
  const int wpb = 8;
  screen_kernel<true><<<(unsigned)((np + wpb - 1) / wpb), wpb * 32>>>(st_->view(d_want.p), d_off.p, d_lists.p);
  LB_CUDA_CHECK(cudaGetLastError());
  LB_CUDA_CHECK(cudaMemcpy(lists.data(), d_lists.p, (size_t)total * sizeof(int), cudaMemcpyDeviceToHost));
}

}  // namespace gxb
