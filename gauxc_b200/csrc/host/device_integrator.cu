// Device execution space: runtime, device-resident task plan, SSF weights driver,
// EXC/VXC integrator, NCCL reduction driver.  Host orchestration only -- every numerical
// stage is one of the kernels in ../cuda.  There is deliberately no CPU fallback: without a
// CUDA device every entry point here throws.
#include "xc_integrator.hpp"
#include "../cuda/kernels.cuh"
#include <algorithm>
#include <array>
#include <cstdlib>
#include <cstring>
#include <dlfcn.h>
#include <mutex>
#include <nccl.h>
#include <numeric>
#include <queue>

namespace GauXC {

#define CUDA_CHECK(expr)                                                              \
  do {                                                                                \
    cudaError_t _e = (expr);                                                          \
    if (_e != cudaSuccess)                                                            \
      GAUXC_GENERIC_EXCEPTION(std::string("CUDA Failed: ") + cudaGetErrorString(_e) + \
                              " in " #expr);                                          \
  } while (0)

static void require_device() {
  int n = 0;
  cudaError_t e = cudaGetDeviceCount(&n);
  if (e != cudaSuccess || n == 0) {
    cudaGetLastError();
    GAUXC_GENERIC_EXCEPTION(
        "No CUDA device: the Device execution space has no CPU fallback in this build");
  }
}

// ------------------------------------------------------------------------------------
// runtime
// ------------------------------------------------------------------------------------
DeviceRuntimeEnvironment::DeviceRuntimeEnvironment(double fill_fraction)
    : fill_fraction_(fill_fraction) {
  if (fill_fraction <= 0. || fill_fraction > 1.) GAUXC_GENERIC_EXCEPTION("Invalid Fill Fraction");
}
DeviceRuntimeEnvironment::DeviceRuntimeEnvironment(void* mem, size_t sz)
    : fill_fraction_(0.), user_mem_(mem), user_mem_sz_(sz) {}

template <typename T>
struct DevBuf {
  T* p = nullptr;
  size_t n = 0;
  DevBuf() = default;
  DevBuf(const DevBuf&) = delete;
  DevBuf& operator=(const DevBuf&) = delete;
  ~DevBuf() { release(); }
  void release() {
    if (p) cudaFree(p);
    p = nullptr;
    n = 0;
  }
  void alloc(size_t cnt) {
    release();
    n = cnt;
    if (cnt) CUDA_CHECK(cudaMalloc((void**)&p, cnt * sizeof(T)));
  }
  void upload(const std::vector<T>& v, cudaStream_t s = 0) {
    alloc(v.size());
    if (!v.empty()) CUDA_CHECK(cudaMemcpyAsync(p, v.data(), v.size() * sizeof(T), cudaMemcpyHostToDevice, s));
  }
};

// ------------------------------------------------------------------------------------
// device plan: everything about the local tasks that does not change between calls
// ------------------------------------------------------------------------------------
struct Schedule {
  // per xc-kind batching of the tiles into workspace-sized groups
  std::vector<gxb::DevTile> tiles;  // with ws_off
  struct Batch {
    int tile_begin, tile_end, item_begin, item_end;
  };
  std::vector<Batch> batches;
  std::vector<gxb::VxcItem> items;
  // fused kernel: persistent CTAs pull tiles of the batch from a device-side queue (one counter
  // per batch) in task order, so the tiles in flight at any time belong to the same / neighbouring
  // parent atoms and their P_sub gathers and VXC scatters stay L2-resident
  int ncta = 0;
  DevBuf<gxb::DevTile> d_tiles;
  DevBuf<gxb::VxcItem> d_items;
  DevBuf<int> d_counters;
  size_t ws_doubles = 0;
  int max_batch_tiles = 0;
};

struct DevicePlan {
  int device = -1;  // the CUDA device the buffers live on
  int nbf = 0, natoms = 0;
  size_t npts = 0;
  std::vector<gxb::DevTask> tasks;
  std::vector<gxb::DevTile> tiles;  // ws_off unset; used by the weights kernel
  DevBuf<gxb::DevShell> d_shells;
  DevBuf<double> d_alpha, d_coeff;
  DevBuf<gxb::DevTask> d_tasks;
  DevBuf<gxb::DevTile> d_tiles;
  DevBuf<int> d_task_shells, d_task_shell_bf, d_task_ao;
  DevBuf<int> d_shell_center;  // shell -> atom (BasisSetMap::shell_to_center), EXC gradient
  DevBuf<double> d_px, d_py, d_pz, d_w;
  DevBuf<double> d_atoms, d_rab, d_rab_dist, d_dist_nearest, d_nbr_dist;  // d_rab: 1 / R_AB, d_rab_dist: R_AB
  DevBuf<int> d_nbr_idx;  // per atom: all atoms sorted by distance from it (SSF loop cut-offs)
  double f_dense = 0., sum_nbe_npts = 0.;
  std::map<int, std::shared_ptr<Schedule>> schedules;  // key: nmat
  size_t schedule_ws_bytes = 0;

  gxb::PlanView view() const {
    gxb::PlanView v{};
    v.shells = d_shells.p;
    v.prim_alpha = d_alpha.p;
    v.prim_coeff = d_coeff.p;
    v.tasks = d_tasks.p;
    v.task_shells = d_task_shells.p;
    v.task_shell_bf = d_task_shell_bf.p;
    v.task_ao = d_task_ao.p;
    v.px = d_px.p;
    v.py = d_py.p;
    v.pz = d_pz.p;
    v.w = d_w.p;
    v.nbf = nbf;
    return v;
  }
};

static std::shared_ptr<DevicePlan> build_plan(LoadBalancer& lb) {
  require_device();
  auto plan = std::make_shared<DevicePlan>();
  CUDA_CHECK(cudaGetDevice(&plan->device));
  auto& tasks = lb.get_tasks();
  const auto& basis = lb.basis();
  const auto& bmap = lb.basis_map();
  const auto& mol = lb.molecule();
  const auto& meta = lb.molmeta();

  if (basis.max_l() > 4) GAUXC_GENERIC_EXCEPTION("L > 4 Not Supported on Device");

  // Task order = Morton (Z-curve) order of the task centroids.  The reference device driver sorts by
  // npts*nbe (…exc_vxc.hpp:254-257) to size its per-call batches; here tasks that are close in SPACE
  // are made close in TIME: the ~148 tiles in flight then share most of their shell lists, so their
  // P_sub gathers and VXC scatters hit a small L2-resident region even when P is GBs (measured on
  // ubiquitin: 14 MB distinct P per 148-tile window vs 60 MB in (iParent, shell_list) order).  Load
  // balance comes from the device-side tile queue, not from the order.
  {
    const char* e = std::getenv("GAUXC_B200_TASK_ORDER");
    const std::string mode = e ? e : "morton";
    if (mode == "cost") {
      std::stable_sort(tasks.begin(), tasks.end(), [](const XCTask& a, const XCTask& b) {
        return (a.points.size() * a.bfn_screening.nbe) > (b.points.size() * b.bfn_screening.nbe);
      });
    } else if (mode == "morton" && !tasks.empty()) {
      const size_t nt = tasks.size();
      std::vector<std::array<double, 3>> cen(nt);
      double lo[3] = {1e300, 1e300, 1e300}, hi[3] = {-1e300, -1e300, -1e300};
#pragma omp parallel for schedule(dynamic, 256)
      for (size_t i = 0; i < nt; ++i) {
        double c[3] = {0., 0., 0.};
        for (auto& p : tasks[i].points)
          for (int d = 0; d < 3; ++d) c[d] += p[d];
        const double inv = tasks[i].points.empty() ? 0. : 1. / double(tasks[i].points.size());
        for (int d = 0; d < 3; ++d) cen[i][d] = c[d] * inv;
      }
      for (size_t i = 0; i < nt; ++i)
        for (int d = 0; d < 3; ++d) {
          lo[d] = std::min(lo[d], cen[i][d]);
          hi[d] = std::max(hi[d], cen[i][d]);
        }
      auto spread = [](uint64_t x) {  // 21 bits -> every third bit
        x &= 0x1fffff;
        x = (x | x << 32) & 0x1f00000000ffffull;
        x = (x | x << 16) & 0x1f0000ff0000ffull;
        x = (x | x << 8) & 0x100f00f00f00f00full;
        x = (x | x << 4) & 0x10c30c30c30c30c3ull;
        x = (x | x << 2) & 0x1249249249249249ull;
        return x;
      };
      std::vector<std::pair<uint64_t, size_t>> key(nt);
      for (size_t i = 0; i < nt; ++i) {
        uint64_t code = 0;
        for (int d = 0; d < 3; ++d) {
          const double ext = std::max(hi[d] - lo[d], 1e-12);
          const uint64_t qd = (uint64_t)std::min(2097151., (cen[i][d] - lo[d]) / ext * 2097151.);
          code |= spread(qd) << d;
        }
        key[i] = {code, i};
      }
      std::sort(key.begin(), key.end());
      std::vector<XCTask> sorted(nt);
#pragma omp parallel for schedule(static)
      for (size_t i = 0; i < nt; ++i) sorted[i] = std::move(tasks[key[i].second]);
      tasks.swap(sorted);
    }
  }

  plan->nbf = bmap.nbf;
  plan->natoms = (int)mol.size();

  // shells
  std::vector<gxb::DevShell> shells(basis.size());
  std::vector<double> alpha, coeff;
  for (size_t s = 0; s < basis.size(); ++s) {
    auto& sh = basis[s];
    gxb::DevShell d{};
    d.x = sh.O[0]; d.y = sh.O[1]; d.z = sh.O[2];
    d.l = sh.l; d.pure = sh.pure; d.nprim = sh.nprim;
    d.prim_off = (int)alpha.size();
    d.ao_off = bmap.shell_to_ao_range[s].first;
    d.nfunc = sh.size();
    for (int k = 0; k < sh.nprim; ++k) {
      alpha.push_back(sh.alpha[k]);
      coeff.push_back(sh.coeff[k]);
    }
    shells[s] = d;
  }

  // flat device arrays: offsets by a serial prefix sum over the tasks, contents by all threads (146 M points and 1e6 tasks
  // for the 2499-atom cluster: the serial version of this loop was seconds of every first call)
  const size_t ntask = tasks.size();
  std::vector<size_t> pt_off(ntask + 1, 0), sh_off(ntask + 1, 0), ao_off(ntask + 1, 0), tl_off(ntask + 1, 0);
  for (size_t it = 0; it < ntask; ++it) {
    const auto& t = tasks[it];
    pt_off[it + 1] = pt_off[it] + t.points.size();
    sh_off[it + 1] = sh_off[it] + t.bfn_screening.shell_list.size();
    ao_off[it + 1] = ao_off[it] + (size_t)t.bfn_screening.nbe;
    tl_off[it + 1] = tl_off[it] + (t.points.size() + gxb::TP - 1) / gxb::TP;
  }
  const size_t npts = pt_off[ntask];
  if (npts > (size_t)std::numeric_limits<int>::max()) GAUXC_GENERIC_EXCEPTION("Too Many Local Points");
  if (ao_off[ntask] > (size_t)std::numeric_limits<int>::max() || sh_off[ntask] > (size_t)std::numeric_limits<int>::max())
    GAUXC_GENERIC_EXCEPTION("Too Many Local Tasks");
  plan->npts = npts;
  std::vector<double> px(npts), py(npts), pz(npts), w(npts);
  std::vector<int> task_shells(sh_off[ntask]), task_shell_bf(sh_off[ntask]), task_ao(ao_off[ntask]);
  plan->tasks.resize(ntask);
  plan->tiles.resize(tl_off[ntask]);
  double f_dense = 0., sum_nbe_npts = 0.;
  int bad_nbe = 0;
#pragma omp parallel for schedule(dynamic, 64) reduction(+ : f_dense, sum_nbe_npts, bad_nbe)
  for (size_t it = 0; it < ntask; ++it) {
    const auto& t = tasks[it];
    gxb::DevTask d{};
    d.shell_off = (int)sh_off[it];
    d.nshells = (int)t.bfn_screening.shell_list.size();
    d.ao_off = (int)ao_off[it];
    d.nbe = t.bfn_screening.nbe;
    d.pt_off = (int)pt_off[it];
    d.npts = (int)t.points.size();
    d.iParent = t.iParent;
    int bf = 0;
    size_t q = sh_off[it], a_out = ao_off[it];
    for (int sidx : t.bfn_screening.shell_list) {
      task_shells[q] = sidx;
      task_shell_bf[q] = bf;
      ++q;
      const auto r = bmap.shell_to_ao_range[sidx];
      if (bf + (r.second - r.first) <= d.nbe)
        for (int a = r.first; a < r.second; ++a) task_ao[a_out++] = a;
      bf += r.second - r.first;
    }
    if (bf != d.nbe) ++bad_nbe;
    const size_t off = pt_off[it];
    for (size_t i = 0; i < t.points.size(); ++i) {
      px[off + i] = t.points[i][0];
      py[off + i] = t.points[i][1];
      pz[off + i] = t.points[i][2];
      w[off + i] = t.weights[i];
    }
    size_t tl = tl_off[it];
    for (int p0 = 0; p0 < d.npts; p0 += gxb::TP) {
      gxb::DevTile tile{};
      tile.task = (int)it;
      tile.pt_off = d.pt_off + p0;
      tile.npts = std::min(gxb::TP, d.npts - p0);
      tile.nbe = d.nbe;
      tile.ao_off = d.ao_off;
      plan->tiles[tl++] = tile;
    }
    f_dense += 4. * double(d.nbe) * double(d.nbe) * double(d.npts);
    sum_nbe_npts += double(d.nbe) * double(d.npts);
    plan->tasks[it] = d;
  }
  if (bad_nbe) GAUXC_GENERIC_EXCEPTION("Inconsistent NBE in Task");
  plan->f_dense = f_dense;
  plan->sum_nbe_npts = sum_nbe_npts;

  std::vector<double> atoms(3 * mol.size());
  for (size_t a = 0; a < mol.size(); ++a) {
    atoms[3 * a] = mol[a].x; atoms[3 * a + 1] = mol[a].y; atoms[3 * a + 2] = mol[a].z;
  }

  plan->d_shells.upload(shells);
  plan->d_alpha.upload(alpha);
  plan->d_coeff.upload(coeff);
  plan->d_tasks.upload(plan->tasks);
  plan->d_tiles.upload(plan->tiles);
  plan->d_task_shells.upload(task_shells);
  plan->d_task_shell_bf.upload(task_shell_bf);
  plan->d_task_ao.upload(task_ao);
  {
    std::vector<int> sc(bmap.shell_to_center.begin(), bmap.shell_to_center.end());
    plan->d_shell_center.upload(sc);
  }
  plan->d_px.upload(px);
  plan->d_py.upload(py);
  plan->d_pz.upload(pz);
  plan->d_w.upload(w);
  plan->d_atoms.upload(atoms);
  {
    // the SSF kernel multiplies by 1 / R_AB (an FP64 division costs ~15 FP64 operations per pair)
    std::vector<double> rab_inv(meta.rab.size());
    for (size_t q = 0; q < rab_inv.size(); ++q) rab_inv[q] = meta.rab[q] > 0. ? 1. / meta.rab[q] : 0.;
    plan->d_rab.upload(rab_inv);
    plan->d_rab_dist.upload(meta.rab);
  }
  plan->d_dist_nearest.upload(meta.dist_nearest);
  {
    const size_t na = mol.size();
    std::vector<int> nbr_idx(na * na);
    std::vector<double> nbr_dist(na * na);
#pragma omp parallel for schedule(static)
    for (size_t a = 0; a < na; ++a) {
      int* idx = nbr_idx.data() + a * na;
      const double* r = meta.rab.data() + a * na;
      std::iota(idx, idx + na, 0);
      std::sort(idx, idx + na, [&](int x, int y) {
        if (x == (int)a || y == (int)a) return x == (int)a && y != (int)a;  // the atom itself first
        return r[x] < r[y] || (r[x] == r[y] && x < y);
      });
      for (size_t k = 0; k < na; ++k) nbr_dist[a * na + k] = (idx[k] == (int)a) ? 0. : r[idx[k]];
    }
    plan->d_nbr_idx.upload(nbr_idx);
    plan->d_nbr_dist.upload(nbr_dist);
  }
  CUDA_CHECK(cudaDeviceSynchronize());
  return plan;
}

std::shared_ptr<DevicePlan> get_device_plan(LoadBalancer& lb) {
  lb.get_tasks();
  int dev = -1;
  if (cudaGetDevice(&dev) != cudaSuccess) dev = -1;
  const bool moved = lb.device_cache && std::static_pointer_cast<DevicePlan>(lb.device_cache)->device != dev;
  if (moved || !lb.device_cache || lb.device_cache_version != lb.version()) {
    // a rebuild reads the host task list: weights that so far live only in the old plan come home first
    lb.sync_host_tasks();
    if (moved) lb.device_cache.reset();  // the calling thread moved to another device: rebuild there
  }
  if (!lb.device_cache || lb.device_cache_version != lb.version()) {
    lb.device_cache = build_plan(lb);
    lb.device_cache_version = lb.version();
  }
  return std::static_pointer_cast<DevicePlan>(lb.device_cache);
}

static std::shared_ptr<Schedule> build_schedule(DevicePlan& plan, int nmat, size_t ws_doubles,
                                                int tiles_per_item, int ncta, bool sym) {
  auto sc = std::make_shared<Schedule>();
  sc->tiles = plan.tiles;
  sc->ncta = ncta;
  size_t cur = 0, ws_max = 0;
  Schedule::Batch b{0, 0, 0, 0};
  auto close_batch = [&](int tile_end) {
    b.tile_end = tile_end;
    // VXC items: runs of tiles of one task, cut into chunks, times the 128 x 128 output blocks
    // (lower block triangle only when M = B^T Z is symmetric, i.e. LDA)
    b.item_begin = (int)sc->items.size();
    int q = b.tile_begin;
    while (q < b.tile_end) {
      int e = q;
      while (e < b.tile_end && sc->tiles[e].task == sc->tiles[q].task) ++e;
      const int nbe = plan.tasks[sc->tiles[q].task].nbe;
      const int nbm = (nbe + gxb::VXC_BLK - 1) / gxb::VXC_BLK;
      const int nbn = (nbe + gxb::VXC_BLN - 1) / gxb::VXC_BLN;
      const size_t need = (size_t)gxb::tile_rows(nmat, nbe) * gxb::TP;
      for (int c = q; c < e; c += tiles_per_item) {
        const int ce = std::min(e, c + tiles_per_item);
        for (int t = c; t < ce; ++t) {
          // what the self-contained item record relies on
          if (sc->tiles[t].ws_off != sc->tiles[c].ws_off + (int64_t)((t - c) * need) ||
              (t + 1 < ce && sc->tiles[t].npts != gxb::TP))
            GAUXC_GENERIC_EXCEPTION("Inconsistent Tile Layout In VXC Item");
        }
        for (int im = 0; im < nbm; ++im)
          for (int in = 0; in < nbn; ++in) {
            // symmetric M (LDA): only blocks that reach the lower triangle
            if (sym && in * gxb::VXC_BLN > im * gxb::VXC_BLK + gxb::VXC_BLK - 1) continue;
            gxb::VxcItem it{};
            it.nbe = nbe;
            it.ao_off = plan.tasks[sc->tiles[q].task].ao_off;
            it.mblk = im;
            it.nblk = in;
            it.row0 = (int)(sc->tiles[c].ws_off / gxb::TP);
            it.ntiles = ce - c;
            it.nks_last = (sc->tiles[ce - 1].npts + 15) / 16;
            sc->items.push_back(it);
          }
      }
      q = e;
    }
    b.item_end = (int)sc->items.size();
    // Few items per CTA (small molecules): the persistent CTAs drain the queue in a handful of pops,
    // so one long item popped last would dominate the kernel.  Longest-first order then bounds the
    // tail by the smallest items; L2 locality of the task order is irrelevant at that size.  Large
    // batches keep the task order (VXC regions stay L2-resident) -- their tail is negligible.
    auto cost = [](const gxb::VxcItem& x) {
      const long long rows = std::min(gxb::VXC_BLK, x.nbe - x.mblk * gxb::VXC_BLK);
      const long long cols = std::min(gxb::VXC_BLN, x.nbe - x.nblk * gxb::VXC_BLN);
      const long long nks = (long long)(x.ntiles - 1) * (gxb::TP / 16) + x.nks_last;
      return rows * cols * nks;
    };
    auto by_cost = [&](const gxb::VxcItem& x, const gxb::VxcItem& y) { return cost(x) > cost(y); };
    if (b.item_end - b.item_begin < 64 * ncta) {
      std::stable_sort(sc->items.begin() + b.item_begin, sc->items.begin() + b.item_end, by_cost);
    } else {
      // large batch: only the END of the queue is ordered longest-first, so the last pops of the
      // persistent CTAs are the cheapest items of that stretch
      std::stable_sort(sc->items.begin() + (b.item_end - 8 * ncta), sc->items.begin() + b.item_end, by_cost);
    }
    sc->batches.push_back(b);
    sc->max_batch_tiles = std::max(sc->max_batch_tiles, b.tile_end - b.tile_begin);
    ws_max = std::max(ws_max, cur);
  };
  for (int i = 0; i < (int)sc->tiles.size(); ++i) {
    const size_t need = (size_t)gxb::tile_rows(nmat, plan.tasks[sc->tiles[i].task].nbe) * gxb::TP;
    if (need > ws_doubles) GAUXC_GENERIC_EXCEPTION("Device Workspace Too Small For One Tile");
    if (cur + need > ws_doubles) {
      close_batch(i);
      b.tile_begin = i;
      cur = 0;
    }
    sc->tiles[i].ws_off = (int64_t)cur;
    cur += need;
  }
  if (!sc->tiles.empty()) close_batch((int)sc->tiles.size());
  sc->ws_doubles = ws_max;
  sc->d_tiles.upload(sc->tiles);
  sc->d_items.upload(sc->items);
  sc->d_counters.alloc(std::max<size_t>(1, 12 * sc->batches.size()));  // queue heads: fused, vxc (UKS: several each); EXC gradient: up to 8 X passes + 2
  CUDA_CHECK(cudaDeviceSynchronize());
  return sc;
}

// ------------------------------------------------------------------------------------
// TMA descriptors over the batch workspace viewed as [rows][128 points] FP64
// ------------------------------------------------------------------------------------
static CUtensorMap make_ws_tensor_map(double* base, size_t ndoubles, int box_cols, int box_rows) {
  using encode_t = CUresult (*)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                const cuuint64_t*, const cuuint32_t*, const cuuint32_t*,
                                CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion,
                                CUtensorMapFloatOOBfill);
  static encode_t encode = nullptr;
  if (!encode) {
    void* fn = nullptr;
    cudaDriverEntryPointQueryResult qres;
    CUDA_CHECK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres));
    if (!fn || qres != cudaDriverEntryPointSuccess)
      GAUXC_GENERIC_EXCEPTION("CUDA Failed: cuTensorMapEncodeTiled unavailable");
    encode = (encode_t)fn;
  }
  CUtensorMap m;
  std::memset(&m, 0, sizeof(m));
  const cuuint64_t rows = std::max<size_t>(1, ndoubles / gxb::TP);
  const cuuint64_t gdim[2] = {(cuuint64_t)gxb::TP, rows};
  const cuuint64_t gstride[1] = {(cuuint64_t)gxb::TP * sizeof(double)};
  const cuuint32_t box[2] = {(cuuint32_t)box_cols, (cuuint32_t)box_rows};
  const cuuint32_t estr[2] = {1, 1};
  const CUresult r = encode(&m, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 2, base, gdim, gstride, box, estr,
                            CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
                            CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS)
    GAUXC_GENERIC_EXCEPTION("CUDA Failed: cuTensorMapEncodeTiled returned " + std::to_string((int)r));
  return m;
}

// ------------------------------------------------------------------------------------
// functional names (tests/standalone_driver.cxx:428-433 uses ExchCXX functional_map)
// ------------------------------------------------------------------------------------
XCFunctional functional_from_string(const std::string& spec_in, bool polarized) {
  std::string spec = spec_in;
  std::transform(spec.begin(), spec.end(), spec.begin(), ::toupper);
  XCFunctional f;
  f.name = spec;
  f.polarized = polarized;
  auto set = [&](bool gga, std::initializer_list<std::pair<int, double>> ks) {
    f.desc.is_gga = gga;
    f.desc.nkern = 0;
    for (auto& k : ks) {
      f.desc.kern[f.desc.nkern] = k.first;
      f.desc.coeff[f.desc.nkern] = k.second;
      ++f.desc.nkern;
    }
  };
  using namespace gxb;
  if (spec == "SVWN5") set(false, {{K_SLATER_X, 1.}, {K_VWN5_C, 1.}});
  else if (spec == "LDA" || spec == "SLATER") set(false, {{K_SLATER_X, 1.}});
  else if (spec == "VWN5") set(false, {{K_VWN5_C, 1.}});
  else if (spec == "SPW92") set(false, {{K_SLATER_X, 1.}, {K_PW92_C, 1.}});
  else if (spec == "PBE") set(true, {{K_PBE_X, 1.}, {K_PBE_C, 1.}});
  else if (spec == "PBE0") { set(true, {{K_PBE_X, 0.75}, {K_PBE_C, 1.}}); f.hyb_exx = 0.25; }
  else if (spec == "REVPBE") set(true, {{K_REVPBE_X, 1.}, {K_PBE_C, 1.}});
  else if (spec == "REVPBE0") { set(true, {{K_REVPBE_X, 0.75}, {K_PBE_C, 1.}}); f.hyb_exx = 0.25; }
  else if (spec == "BLYP") set(true, {{K_B88_X, 1.}, {K_LYP_C, 1.}});
  // libxc hyb_gga_xc_b3lyp: 0.08 LDA_X + 0.72 B88 + 0.19 VWN_RPA + 0.81 LYP (+ 0.20 exact exchange, the caller's)
  else if (spec == "B3LYP") { set(true, {{K_SLATER_X, 0.08}, {K_B88_X, 0.72}, {K_VWN5_C, 0.19}, {K_LYP_C, 0.81}}); f.hyb_exx = 0.20; }
  else GAUXC_GENERIC_EXCEPTION("Functional NYI in B200 path: " + spec_in);
  if (polarized) {
    // spin-polarised forms exist for Slater, VWN (RPA set), B88, LYP, PBE exchange (spin scaling) and PBE correlation
    bool ok = true;
    for (int k = 0; k < f.desc.nkern; ++k) ok = ok && f.desc.kern[k] != K_PW92_C && f.desc.kern[k] != K_VWN3_C;
    if (!ok) GAUXC_GENERIC_EXCEPTION("Polarized (UKS) Functional NYI in B200 path: " + spec_in);
  }
  return f;
}

// ------------------------------------------------------------------------------------
// NCCL reduction driver (dlopen: the library must load on boxes without NCCL)
// ------------------------------------------------------------------------------------
namespace {
struct NcclApi {
  void* h = nullptr;
  ncclResult_t (*GetUniqueId)(ncclUniqueId*) = nullptr;
  ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
  ncclResult_t (*AllReduce)(const void*, void*, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t,
                            cudaStream_t) = nullptr;
  ncclResult_t (*AllGather)(const void*, void*, size_t, ncclDataType_t, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
  const char* (*GetErrorString)(ncclResult_t) = nullptr;
  void load() {
    if (h) return;
    h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
    if (!h) h = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
    if (!h) GAUXC_GENERIC_EXCEPTION("NCCL FAILED: cannot load libnccl.so.2");
    GetUniqueId = (decltype(GetUniqueId))dlsym(h, "ncclGetUniqueId");
    CommInitRank = (decltype(CommInitRank))dlsym(h, "ncclCommInitRank");
    AllReduce = (decltype(AllReduce))dlsym(h, "ncclAllReduce");
    AllGather = (decltype(AllGather))dlsym(h, "ncclAllGather");
    CommDestroy = (decltype(CommDestroy))dlsym(h, "ncclCommDestroy");
    GetErrorString = (decltype(GetErrorString))dlsym(h, "ncclGetErrorString");
    if (!GetUniqueId || !CommInitRank || !AllReduce || !CommDestroy)
      GAUXC_GENERIC_EXCEPTION("NCCL FAILED: missing symbols");
  }
};
NcclApi g_nccl;
ncclComm_t g_comm = nullptr;
int g_comm_rank = 0, g_comm_size = 1;

struct NoopReductionDriver : ReductionDriver {
  bool takes_host_memory() const override { return true; }
  bool takes_device_memory() const override { return true; }
  void allreduce_inplace(double*, size_t, ReductionOp, void*) override {}
  int comm_size() const override { return 1; }
};
struct NCCLReductionDriver : ReductionDriver {
  bool takes_host_memory() const override { return false; }
  bool takes_device_memory() const override { return true; }
  int comm_size() const override { return g_comm_size; }
  void allreduce_inplace(double* data, size_t n, ReductionOp, void* stream) override {
    if (!g_comm) GAUXC_GENERIC_EXCEPTION("NCCL FAILED: communicator not initialised");
    auto r = g_nccl.AllReduce(data, data, n, ncclDouble, ncclSum, g_comm, (cudaStream_t)stream);
    if (r != ncclSuccess) GAUXC_GENERIC_EXCEPTION("NCCL FAILED");
  }
  bool can_allgather() const override { return g_nccl.AllGather != nullptr; }
  void allgather_inplace(double* data, size_t count_per_rank, void* stream) override {
    if (!g_comm) GAUXC_GENERIC_EXCEPTION("NCCL FAILED: communicator not initialised");
    auto r = g_nccl.AllGather(data + (size_t)g_comm_rank * count_per_rank, data, count_per_rank, ncclDouble, g_comm,
                              (cudaStream_t)stream);
    if (r != ncclSuccess) GAUXC_GENERIC_EXCEPTION("NCCL FAILED");
  }
};
}  // namespace

void nccl_get_unique_id(char out[128]) {
  g_nccl.load();
  ncclUniqueId id;
  static_assert(sizeof(ncclUniqueId) == 128, "ncclUniqueId size");
  if (g_nccl.GetUniqueId(&id) != ncclSuccess) GAUXC_GENERIC_EXCEPTION("NCCL FAILED");
  std::memcpy(out, &id, 128);
}
void nccl_init_global(const char idb[128], int rank, int size) {
  require_device();
  g_nccl.load();
  if (g_comm) nccl_finalize_global();
  ncclUniqueId id;
  std::memcpy(&id, idb, 128);
  if (g_nccl.CommInitRank(&g_comm, size, id, rank) != ncclSuccess)
    GAUXC_GENERIC_EXCEPTION("NCCL FAILED");
  g_comm_rank = rank;
  g_comm_size = size;
}
void nccl_finalize_global() {
  if (g_comm) g_nccl.CommDestroy(g_comm);
  g_comm = nullptr;
  g_comm_size = 1;
}

std::shared_ptr<ReductionDriver> make_reduction_driver(const RuntimeEnvironment& rt,
                                                       const std::string& name_in) {
  std::string name = name_in;
  std::transform(name.begin(), name.end(), name.begin(), ::toupper);
  if (name == "BASICMPI")
    GAUXC_GENERIC_EXCEPTION("BasicMPI ReductionDriver unavailable: no MPI in this build (use NCCL)");
  if (name != "DEFAULT" && name != "NCCL")
    GAUXC_GENERIC_EXCEPTION("ReductionDriver Not Recognized: " + name_in);
  if (rt.comm_size() == 1) return std::make_shared<NoopReductionDriver>();
  if (!g_comm || g_comm_size != rt.comm_size())
    GAUXC_GENERIC_EXCEPTION("NCCL FAILED: call gauxc_b200_nccl_init before creating the integrator");
  return std::make_shared<NCCLReductionDriver>();
}

// ------------------------------------------------------------------------------------
// molecular weights
// ------------------------------------------------------------------------------------
MolecularWeights::MolecularWeights(ExecutionSpace ex, const std::string&, MolecularWeightsSettings s)
    : ex_(ex), settings_(s) {
  if (ex != ExecutionSpace::Device)
    GAUXC_GENERIC_EXCEPTION("Host MolecularWeights are not part of the B200 build (Device only)");
}

void MolecularWeights::modify_weights(LoadBalancer& lb) {
  // src/molecular_weights/device/device_molecular_weights.cxx:18-88
  if (lb.state().modified_weights_are_stored)
    GAUXC_GENERIC_EXCEPTION("Attempting to Overwrite Modified Weights");
  if (settings_.weight_alg != XCWeightAlg::SSF)
    GAUXC_GENERIC_EXCEPTION("Non-SSF Weights NYI for Device Integration");

  auto plan = get_device_plan(lb);  // sorts tasks, uploads raw quadrature weights
  cudaEvent_t e0, e1;
  CUDA_CHECK(cudaEventCreate(&e0));
  CUDA_CHECK(cudaEventCreate(&e1));
  CUDA_CHECK(cudaEventRecord(e0, 0));
  gxb::launch_ssf_weights(plan->view(), plan->d_tiles.p, (int)plan->tiles.size(), plan->d_atoms.p,
                          plan->d_rab_dist.p, plan->d_rab.p, plan->d_dist_nearest.p, plan->d_nbr_idx.p, plan->d_nbr_dist.p,
                          plan->natoms, 0);
  CUDA_CHECK(cudaGetLastError());
  CUDA_CHECK(cudaEventRecord(e1, 0));
  CUDA_CHECK(cudaEventSynchronize(e1));
  float ms = 0;
  CUDA_CHECK(cudaEventElapsedTime(&ms, e0, e1));
  cudaEventDestroy(e0);
  cudaEventDestroy(e1);
  timer_.add("MolecularWeights", ms);
  // The weights stay on the device, where the integrator reads them.  The host XCTask list is brought up to
  // date only when a consumer asks for it (LoadBalancer::sync_host_tasks, called by the task accessors of the
  // C ABI): the reference copies every weight back task by task (device_molecular_weights.cxx:60-80), 1.2 GB
  // and a host loop over 1e6 tasks for the 2499-atom cluster.
  std::weak_ptr<DevicePlan> wplan = plan;
  lb.set_host_sync([wplan](std::vector<XCTask>& tasks) {
    auto p = wplan.lock();
    if (!p || !p->npts) return;
    int prev = -1;
    cudaGetDevice(&prev);
    if (prev != p->device) cudaSetDevice(p->device);
    std::vector<double> w(p->npts);
    cudaError_t e = cudaMemcpy(w.data(), p->d_w.p, p->npts * sizeof(double), cudaMemcpyDeviceToHost);
    if (prev != p->device && prev >= 0) cudaSetDevice(prev);
    if (e != cudaSuccess) GAUXC_GENERIC_EXCEPTION(std::string("CUDA Failed: ") + cudaGetErrorString(e));
    size_t off = 0;
    for (auto& t : tasks) {
      std::copy(w.begin() + off, w.begin() + off + t.weights.size(), t.weights.begin());
      off += t.weights.size();
      t.max_weight = t.weights.empty() ? 0. : *std::max_element(t.weights.begin(), t.weights.end());
    }
  });
  lb.state().modified_weights_are_stored = true;
  lb.state().weight_alg = XCWeightAlg::SSF;
  // device copy is already current: no version bump
}

// ------------------------------------------------------------------------------------
// integrator
// ------------------------------------------------------------------------------------
struct XCIntegrator::Impl {
  cudaStream_t stream = nullptr;
  cudaStream_t copy_stream = nullptr;  // H2D of P overlaps the collocation of the first batch
  cudaEvent_t e_p_ready{};
  bool p_pending = false;  // set by the host-buffer entry points: wait for e_p_ready before P is read
  DevBuf<double> dP, dPtri, dVXC, d_ws, d_exc_part, d_nel_part, d_out2, d_pack;
  DevBuf<double> dPz, dPtri_z, dVXCz, d_uks_den;  // UKS: z density / potential, densities per point
  DevBuf<double> d_grad, d_wf;  // EXC gradient: 3 natoms sums, w eps rho per point (weight derivatives)
  gxb::TmapSet tmapA{};  // TMA views of d_ws: 16 rows x (32..128) points
  CUtensorMap tmapV{}, tmapZ{};  // 128 / 64 rows x 16 points
  int ncta = 0;
  std::shared_ptr<DevicePlan> plan;
  std::shared_ptr<Schedule> sched;
  int sched_nmat = 0;
  bool profile = false;
  std::vector<cudaEvent_t> ev;  // profile mode: 5 events per batch, read after the final sync
  cudaEvent_t e_begin{}, e_lw0{}, e_lw1{}, e_end{};
  double* h_out2 = nullptr;  // pinned
  ~Impl() {
    for (auto e : ev) cudaEventDestroy(e);
    if (stream) cudaStreamDestroy(stream);
    if (copy_stream) cudaStreamDestroy(copy_stream);
    if (e_p_ready) cudaEventDestroy(e_p_ready);
    if (h_out2) cudaFreeHost(h_out2);
  }

  // plan + schedule + workspace + TMA views for tiles of `nmat` matrices (B [, dBx, dBy, dBz], Z [, Z_z])
  void prepare(LoadBalancer& lb, int nmat, bool sym);
  // P (nbf x nbf, ld = nbf) and VXC buffers of the host-buffer entry points; P is padded so that an
  // all-gather of per-rank column slabs fits
  void ensure_matrices(size_t nbf, int nranks, bool uks);
};

static size_t workspace_bytes(const LoadBalancer& lb) {
  if (const char* e = std::getenv("GAUXC_B200_WORKSPACE_MB")) return (size_t)std::atoll(e) << 20;
  size_t free_b = 0, tot_b = 0;
  cudaMemGetInfo(&free_b, &tot_b);
  double frac = 0.9;
  if (auto* d = dynamic_cast<const DeviceRuntimeEnvironment*>(&lb.runtime()))
    if (d->fill_fraction() > 0.) frac = d->fill_fraction();
  size_t cap = (size_t)16 << 30;  // default: 16 GiB of B / dB workspace per batch
  if (const char* e = std::getenv("GAUXC_DEVICE_MEMORY_CAP")) cap = (size_t)std::atoll(e);
  return std::min<size_t>(cap, (size_t)(frac * free_b * 0.8));
}

void XCIntegrator::Impl::prepare(LoadBalancer& lb, int nmat, bool sym) {
  {
    auto np = get_device_plan(lb);
    if (np != plan) {
      plan = np;
      sched.reset();
    }
  }
  if (plan->nbf && (size_t)plan->nbf * plan->nbf > (size_t)std::numeric_limits<int>::max() * 64)
    GAUXC_GENERIC_EXCEPTION("Basis Too Large");
  if (sched && sched_nmat == nmat) return;
  auto it = plan->schedules.find(nmat);
  if (it == plan->schedules.end()) {
    const size_t wsb = workspace_bytes(lb);
    int tpi = 16;
    if (const char* e = std::getenv("GAUXC_B200_TILES_PER_ITEM")) tpi = std::max(1, std::atoi(e));
    if (!ncta) {
      int dev = 0;
      CUDA_CHECK(cudaGetDevice(&dev));
      CUDA_CHECK(cudaDeviceGetAttribute(&ncta, cudaDevAttrMultiProcessorCount, dev));
    }
    plan->schedules[nmat] = build_schedule(*plan, nmat, wsb / sizeof(double), tpi, ncta, sym);
    plan->schedule_ws_bytes = wsb;
    it = plan->schedules.find(nmat);
  }
  sched = it->second;
  sched_nmat = nmat;
  d_ws.alloc(sched->ws_doubles);
  if (sched->ws_doubles) {
    // TMA may read (never use) rows of a neighbouring matrix: keep every byte finite
    CUDA_CHECK(cudaMemsetAsync(d_ws.p, 0, sched->ws_doubles * sizeof(double), stream));
    for (int w = 0; w < 4; ++w) tmapA.m[w] = make_ws_tensor_map(d_ws.p, sched->ws_doubles, 32 * (w + 1), 16);
    tmapV = make_ws_tensor_map(d_ws.p, sched->ws_doubles, 16, gxb::VXC_BLK);
    tmapZ = make_ws_tensor_map(d_ws.p, sched->ws_doubles, 16, gxb::VXC_BLN);
  }
  d_exc_part.alloc(std::max<size_t>(1, sched->tiles.size()));
  d_nel_part.alloc(std::max<size_t>(1, sched->tiles.size()));
}

void XCIntegrator::Impl::ensure_matrices(size_t nbf, int nranks, bool uks) {
  const size_t cpr = (nbf + nranks - 1) / std::max(1, nranks);  // columns per rank slab
  const size_t np = std::max(nbf * nbf, cpr * nranks * nbf);
  if (dP.n != np) {
    dP.alloc(np);
    dVXC.alloc(nbf * nbf);
    d_out2.alloc(2);
  }
  if (uks && dPz.n != np) {
    dPz.alloc(np);
    dVXCz.alloc(nbf * nbf);
  }
}

XCIntegrator::XCIntegrator(ExecutionSpace ex, const std::string& input_type,
                           const std::string& integrator_kernel, const std::string& lwd_kernel,
                           const std::string& reduction_kernel, std::shared_ptr<XCFunctional> func,
                           std::shared_ptr<LoadBalancer> lb)
    : func_(std::move(func)), lb_(std::move(lb)) {
  auto up = [](std::string s) {
    std::transform(s.begin(), s.end(), s.begin(), ::toupper);
    return s;
  };
  if (ex != ExecutionSpace::Device)
    GAUXC_GENERIC_EXCEPTION("Host XCIntegrator is not part of the B200 build (Device only)");
  if (up(input_type) != "REPLICATED") GAUXC_GENERIC_EXCEPTION("INTEGRATOR TYPE NOT RECOGNIZED");
  const auto ik = up(integrator_kernel);
  if (ik != "DEFAULT" && ik != "INCORE")
    GAUXC_GENERIC_EXCEPTION("Integrator Kernel Not Recognized: " + integrator_kernel);
  const auto lk = up(lwd_kernel);
  if (lk != "DEFAULT" && lk != "SCHEME1" && lk != "B200")
    GAUXC_GENERIC_EXCEPTION("LWD Not Recognized: " + lwd_kernel);
  if (!func_ || !lb_) GAUXC_GENERIC_EXCEPTION("Null Functional / LoadBalancer");
  red_ = make_reduction_driver(lb_->runtime(), reduction_kernel);
  require_device();
  impl_ = std::make_shared<Impl>();
  CUDA_CHECK(cudaStreamCreateWithFlags(&impl_->stream, cudaStreamNonBlocking));
  CUDA_CHECK(cudaStreamCreateWithFlags(&impl_->copy_stream, cudaStreamNonBlocking));
  CUDA_CHECK(cudaEventCreateWithFlags(&impl_->e_p_ready, cudaEventDisableTiming));
  CUDA_CHECK(cudaEventCreate(&impl_->e_begin));
  CUDA_CHECK(cudaEventCreate(&impl_->e_lw0));
  CUDA_CHECK(cudaEventCreate(&impl_->e_lw1));
  CUDA_CHECK(cudaEventCreate(&impl_->e_end));
  CUDA_CHECK(cudaMallocHost((void**)&impl_->h_out2, 2 * sizeof(double)));
}

XCIntegrator::~XCIntegrator() = default;
void XCIntegrator::set_profile(bool on) { impl_->profile = on; }
void* XCIntegrator::stream() const { return impl_->stream; }

// Sum over the ranks of the lower triangle(s) of VXC and of {EXC, N_EL} in ONE collective
// (incore_replicated_xc_device_integrator_exc_vxc.hpp:158-162 reduces VXC, EXC and N_EL one by one, full
// matrices): pack [tril(VXC) | (tril(VXCz)) | EXC | N_EL] -> allreduce -> unpack + mirror.  Single rank:
// symmetrise in place.
void XCIntegrator::reduce_and_symmetrize_(double* dV, double* dVz, double* d_out2, int nbf, bool do_vxc) {
  auto& I = *impl_;
  cudaStream_t s = I.stream;
  const int nmatv = do_vxc ? (dVz ? 2 : 1) : 0;
  if (red_->comm_size() <= 1) {
    if (nmatv >= 1) gxb::launch_symmetrize(dV, nbf, nbf, s);
    if (nmatv == 2) gxb::launch_symmetrize(dVz, nbf, nbf, s);
    return;
  }
  const size_t ntri = (size_t)nbf * (nbf + 1) / 2;
  const size_t total = nmatv * ntri + 2;
  if (I.d_pack.n < total) I.d_pack.alloc(total);
  if (nmatv >= 1) gxb::launch_pack_tril(dV, nbf, nbf, I.d_pack.p, s);
  if (nmatv == 2) gxb::launch_pack_tril(dVz, nbf, nbf, I.d_pack.p + ntri, s);
  CUDA_CHECK(cudaMemcpyAsync(I.d_pack.p + nmatv * ntri, d_out2, 2 * sizeof(double), cudaMemcpyDeviceToDevice, s));
  red_->allreduce_inplace(I.d_pack.p, total, ReductionOp::Sum, s);
  if (nmatv >= 1) gxb::launch_unpack_tril_sym(I.d_pack.p, dV, nbf, nbf, s);
  if (nmatv == 2) gxb::launch_unpack_tril_sym(I.d_pack.p + ntri, dVz, nbf, nbf, s);
  CUDA_CHECK(cudaMemcpyAsync(d_out2, I.d_pack.p + nmatv * ntri, 2 * sizeof(double), cudaMemcpyDeviceToDevice, s));
}

void XCIntegrator::eval_exc_vxc_device(const double* dP, double* dVXC, double* d_out2, bool do_vxc) {
  auto& I = *impl_;
  if (func_->polarized) GAUXC_GENERIC_EXCEPTION("RKS Evaluation Requires An Unpolarized Functional");
  if (!lb_->state().modified_weights_are_stored)
    GAUXC_GENERIC_EXCEPTION("Weights Have Not Been Modified");
  const bool gga = func_->is_gga();
  const int nmat = gga ? 5 : 2;
  I.prepare(*lb_, nmat, !gga);
  auto& plan = *I.plan;
  cudaStream_t s = I.stream;
  auto& sc = *I.sched;
  const gxb::PlanView pv = plan.view();
  const int nbf = plan.nbf;

  if (do_vxc) CUDA_CHECK(cudaMemsetAsync(dVXC, 0, sizeof(double) * (size_t)nbf * nbf, s));
  CUDA_CHECK(cudaMemsetAsync(sc.d_counters.p, 0, sizeof(int) * sc.d_counters.n, s));
  CUDA_CHECK(cudaEventRecord(I.e_lw0, s));
  long long launches = 0;
  const double* dPin = dP;
  if (!gga && I.dPtri.n != (size_t)nbf * nbf) {
    I.dPtri.alloc((size_t)nbf * nbf);
    CUDA_CHECK(cudaMemsetAsync(I.dPtri.p, 0, sizeof(double) * (size_t)nbf * nbf, s));
  }
  // First use of P on the stream: wait for a pending H2D (host-buffer entry points), then, for LDA,
  // prepare P' -- the fused kernel walks the lower triangle of the quadratic form (half the DMMA
  // work of X = P_sub B).  Called after the first collocation launch so the upload hides behind it.
  auto first_use_of_P = [&]() {
    if (I.p_pending) {
      CUDA_CHECK(cudaStreamWaitEvent(s, I.e_p_ready, 0));
      I.p_pending = false;
    }
    if (!gga) {
      gxb::launch_sym_half(dPin, nbf, I.dPtri.p, nbf, s);
      ++launches;
    }
  };
  if (!gga) dP = I.dPtri.p;
  double kms[4] = {0, 0, 0, 0};
  // profile mode brackets every kernel with events on the launching stream; they are read
  // after the final synchronise, so the pipeline is not stalled by the measurement
  if (I.profile)
    while (I.ev.size() < 5 * sc.batches.size()) {
      cudaEvent_t e;
      CUDA_CHECK(cudaEventCreate(&e));
      I.ev.push_back(e);
    }
  size_t ib = 0;
  for (auto& b : sc.batches) {
    const int nt = b.tile_end - b.tile_begin;
    const gxb::DevTile* tl = sc.d_tiles.p + b.tile_begin;
    cudaEvent_t* ev = I.profile ? &I.ev[5 * ib] : nullptr;
    if (ev) CUDA_CHECK(cudaEventRecord(ev[0], s));
    gxb::launch_collocation(pv, tl, nt, I.d_ws.p, gga, s);
    if (ev) CUDA_CHECK(cudaEventRecord(ev[1], s));
    if (ib == 0) first_use_of_P();
    CUDA_CHECK(gxb::launch_fused(I.tmapA, pv, tl, nt, sc.d_counters.p + ib, sc.ncta, I.d_ws.p, dP, nbf,
                                 func_->desc, I.d_exc_part.p, I.d_nel_part.p, b.tile_begin, s));
    if (ev) CUDA_CHECK(cudaEventRecord(ev[2], s));
    if (ev) CUDA_CHECK(cudaEventRecord(ev[3], s));
    launches += 2;
    if (do_vxc) {
      CUDA_CHECK(gxb::launch_vxc(I.tmapV, I.tmapZ, pv, sc.d_items.p + b.item_begin, b.item_end - b.item_begin,
                                 sc.d_counters.p + sc.batches.size() + ib, sc.ncta, gga ? 4 : 1, nmat, !gga,
                                 dVXC, nbf, s));
      ++launches;
    }
    if (ev) CUDA_CHECK(cudaEventRecord(ev[4], s));
    ++ib;
  }
  if (sc.batches.empty()) first_use_of_P();  // nothing consumed P: still retire the pending upload
  gxb::launch_reduce_partials(I.d_exc_part.p, I.d_nel_part.p, (int)sc.tiles.size(), d_out2, s);
  ++launches;
  CUDA_CHECK(cudaGetLastError());
  CUDA_CHECK(cudaEventRecord(I.e_lw1, s));

  // replicated-data reduction over the ranks + upper <- lower
  reduce_and_symmetrize_(dVXC, nullptr, d_out2, nbf, do_vxc);
  launches += do_vxc ? (red_->comm_size() > 1 ? 2 : 1) : 0;
  CUDA_CHECK(cudaGetLastError());
  CUDA_CHECK(cudaStreamSynchronize(s));
  if (I.profile)
    for (size_t q = 0; q < sc.batches.size(); ++q)
      for (int k = 0; k < 4; ++k) {
        float ems = 0;
        cudaEventElapsedTime(&ems, I.ev[5 * q + k], I.ev[5 * q + k + 1]);
        kms[k] += ems;
      }
  float ms = 0;
  cudaEventElapsedTime(&ms, I.e_lw0, I.e_lw1);
  stats_.last_local_work_ms = ms;
  timer_.add("XCIntegrator.LocalWork_EXC_VXC", ms);
  for (int k = 0; k < 4; ++k) stats_.kernel_ms[k] = kms[k];
  stats_.kernel_launches = launches;
  stats_.f_dense = plan.f_dense;
  stats_.sum_nbe_npts = plan.sum_nbe_npts;
  stats_.npts = (long long)plan.npts;
  stats_.ntiles = (long long)sc.tiles.size();
  stats_.nbatches = (long long)sc.batches.size();
  stats_.nitems = (long long)sc.items.size();
}

static void check_dims(const LoadBalancer& lb, int64_t m, int64_t n, int64_t ldp, int64_t ldv,
                       bool vxc) {
  // incore_replicated_xc_device_integrator_exc_vxc.hpp:66-92
  const int64_t nbf = lb.basis_map().nbf;
  if (m != n) GAUXC_GENERIC_EXCEPTION("P/VXC Must Be Square");
  if (m != nbf) GAUXC_GENERIC_EXCEPTION("P/VXC Must Have Same Dimension as Basis");
  if (ldp < nbf) GAUXC_GENERIC_EXCEPTION("Invalid LDP");
  if (vxc && ldv < nbf) GAUXC_GENERIC_EXCEPTION("Invalid LDVXC");
}

// Host P -> device, on the copy stream.  One rank: one copy.  Several ranks of one box (the contract
// replicates P in every rank's host memory): every rank uploads only its slab of columns over its own
// PCIe link and the slabs are all-gathered over NVLink -- the host links carry nbf^2 * 8 bytes per box
// instead of per GPU (the reference uploads the full matrix on every rank,
// xc_device_stack_data.cxx:350-384).
void XCIntegrator::upload_density_(const double* P, int64_t ldp, double* dP, size_t nbf) {
  auto& I = *impl_;
  const int nr = red_->comm_size();
  auto copy_cols = [&](size_t c0, size_t c1) {
    if (c1 <= c0) return;
    if ((size_t)ldp == nbf)
      CUDA_CHECK(cudaMemcpyAsync(dP + c0 * nbf, P + c0 * nbf, (c1 - c0) * nbf * sizeof(double),
                                 cudaMemcpyHostToDevice, I.copy_stream));
    else
      CUDA_CHECK(cudaMemcpy2DAsync(dP + c0 * nbf, nbf * sizeof(double), P + c0 * ldp, ldp * sizeof(double),
                                   nbf * sizeof(double), c1 - c0, cudaMemcpyHostToDevice, I.copy_stream));
  };
  size_t min_slab = (size_t)4 << 20;  // below 4 MB per slab the collective costs more than it saves
  if (const char* e = std::getenv("GAUXC_B200_SLAB_UPLOAD_MIN_BYTES")) min_slab = (size_t)std::atoll(e);
  const size_t cpr = (nbf + nr - 1) / nr;
  if (nr > 1 && red_->can_allgather() && cpr * nbf * sizeof(double) >= min_slab) {
    const size_t r = (size_t)lb_->runtime().comm_rank();
    copy_cols(std::min(nbf, r * cpr), std::min(nbf, (r + 1) * cpr));
    red_->allgather_inplace(dP, cpr * nbf, I.copy_stream);
  } else {
    copy_cols(0, nbf);
  }
}

static void download_matrix(double* H, int64_t ldh, const double* D, size_t nbf, cudaStream_t s) {
  if ((size_t)ldh == nbf)
    CUDA_CHECK(cudaMemcpyAsync(H, D, nbf * nbf * sizeof(double), cudaMemcpyDeviceToHost, s));
  else
    CUDA_CHECK(cudaMemcpy2DAsync(H, ldh * sizeof(double), D, nbf * sizeof(double), nbf * sizeof(double), nbf,
                                 cudaMemcpyDeviceToHost, s));
}

void XCIntegrator::eval_exc_vxc(int64_t m, int64_t n, const double* P, int64_t ldp, double* VXC,
                                int64_t ldvxc, double* EXC) {
  check_dims(*lb_, m, n, ldp, ldvxc, true);
  if (!lb_->state().modified_weights_are_stored)
    GAUXC_GENERIC_EXCEPTION("Weights Have Not Been Modified");
  auto& I = *impl_;
  const size_t nbf = (size_t)m;
  cudaStream_t s = I.stream;
  I.ensure_matrices(nbf, red_->comm_size(), false);
  CUDA_CHECK(cudaEventRecord(I.e_begin, s));
  CUDA_CHECK(cudaStreamWaitEvent(I.copy_stream, I.e_begin, 0));
  upload_density_(P, ldp, I.dP.p, nbf);
  CUDA_CHECK(cudaEventRecord(I.e_p_ready, I.copy_stream));
  I.p_pending = true;
  eval_exc_vxc_device(I.dP.p, I.dVXC.p, I.d_out2.p, true);
  if (!vxc_root_only_ || lb_->runtime().comm_rank() == 0) download_matrix(VXC, ldvxc, I.dVXC.p, nbf, s);
  CUDA_CHECK(cudaMemcpyAsync(I.h_out2, I.d_out2.p, 2 * sizeof(double), cudaMemcpyDeviceToHost, s));
  CUDA_CHECK(cudaEventRecord(I.e_end, s));
  CUDA_CHECK(cudaStreamSynchronize(s));
  float ms = 0;
  cudaEventElapsedTime(&ms, I.e_begin, I.e_end);
  stats_.last_total_ms = ms;
  *EXC = I.h_out2[0];
  stats_.n_el = I.h_out2[1];
}

void XCIntegrator::eval_exc(int64_t m, int64_t n, const double* P, int64_t ldp, double* EXC) {
  check_dims(*lb_, m, n, ldp, 0, false);
  auto& I = *impl_;
  const size_t nbf = (size_t)m;
  cudaStream_t s = I.stream;
  I.ensure_matrices(nbf, red_->comm_size(), false);
  CUDA_CHECK(cudaEventRecord(I.e_begin, s));
  CUDA_CHECK(cudaStreamWaitEvent(I.copy_stream, I.e_begin, 0));
  upload_density_(P, ldp, I.dP.p, nbf);
  CUDA_CHECK(cudaEventRecord(I.e_p_ready, I.copy_stream));
  I.p_pending = true;
  eval_exc_vxc_device(I.dP.p, I.dVXC.p, I.d_out2.p, false);  // reduces {EXC, N_EL} over the ranks
  CUDA_CHECK(cudaMemcpyAsync(I.h_out2, I.d_out2.p, 2 * sizeof(double), cudaMemcpyDeviceToHost, s));
  CUDA_CHECK(cudaStreamSynchronize(s));
  *EXC = I.h_out2[0];
  stats_.n_el = I.h_out2[1];
}

// integrate_den (incore_replicated_xc_device_integrator_integrate_den.hpp): N = sum_i w_i B_i^T P B_i with
// X = 1.0 * P * B -- collocation + the density half of the fused kernel only; no functional, no Z factors
// that anything reads, no VXC kernel.  The fused kernel carries the RKS factor 2, so halve.
void XCIntegrator::integrate_den(int64_t m, int64_t n, const double* P, int64_t ldp, double* N_EL) {
  check_dims(*lb_, m, n, ldp, 0, false);
  auto& I = *impl_;
  const size_t nbf = (size_t)m;
  cudaStream_t s = I.stream;
  I.ensure_matrices(nbf, red_->comm_size(), false);
  CUDA_CHECK(cudaEventRecord(I.e_begin, s));
  CUDA_CHECK(cudaStreamWaitEvent(I.copy_stream, I.e_begin, 0));
  upload_density_(P, ldp, I.dP.p, nbf);
  CUDA_CHECK(cudaEventRecord(I.e_p_ready, I.copy_stream));
  CUDA_CHECK(cudaStreamWaitEvent(s, I.e_p_ready, 0));
  I.prepare(*lb_, 2, true);
  auto& plan = *I.plan;
  auto& sc = *I.sched;
  const gxb::PlanView pv = plan.view();
  const int inbf = plan.nbf;
  if (I.dPtri.n != nbf * nbf) {
    I.dPtri.alloc(nbf * nbf);
    CUDA_CHECK(cudaMemsetAsync(I.dPtri.p, 0, sizeof(double) * nbf * nbf, s));
  }
  CUDA_CHECK(cudaMemsetAsync(sc.d_counters.p, 0, sizeof(int) * sc.d_counters.n, s));
  gxb::launch_sym_half(I.dP.p, inbf, I.dPtri.p, inbf, s);
  gxb::FunctionalDesc none{};  // nkern = 0: eps = vrho = 0, only N_EL is accumulated
  long long launches = 1;
  size_t ib = 0;
  for (auto& b : sc.batches) {
    const int nt = b.tile_end - b.tile_begin;
    const gxb::DevTile* tl = sc.d_tiles.p + b.tile_begin;
    gxb::launch_collocation(pv, tl, nt, I.d_ws.p, false, s);
    CUDA_CHECK(gxb::launch_fused(I.tmapA, pv, tl, nt, sc.d_counters.p + ib, sc.ncta, I.d_ws.p, I.dPtri.p, inbf,
                                 none, I.d_exc_part.p, I.d_nel_part.p, b.tile_begin, s));
    launches += 2;
    ++ib;
  }
  gxb::launch_reduce_partials(I.d_exc_part.p, I.d_nel_part.p, (int)sc.tiles.size(), I.d_out2.p, s);
  ++launches;
  CUDA_CHECK(cudaGetLastError());
  if (red_->comm_size() > 1) red_->allreduce_inplace(I.d_out2.p, 2, ReductionOp::Sum, s);
  CUDA_CHECK(cudaMemcpyAsync(I.h_out2, I.d_out2.p, 2 * sizeof(double), cudaMemcpyDeviceToHost, s));
  CUDA_CHECK(cudaStreamSynchronize(s));
  stats_.kernel_launches = launches;
  stats_.n_el = I.h_out2[1];
  *N_EL = 0.5 * I.h_out2[1];
}

// UKS: Ps = P_alpha + P_beta, Pz = P_alpha - P_beta
// (reference_replicated_xc_host_integrator_exc_vxc.hpp:107-601 with is_uks; device counterpart
// incore_replicated_xc_device_integrator_exc_vxc.hpp:46-386).  Per batch: collocation, the fused kernel over Ps
// (rho_s [, grad n] per point), the fused kernel over Pz (rho_z [, grad M_z] -> rho_+- [, gamma_++,+-,--] ->
// polarised functional -> factors of Z_s and Z_z), the VXC rank update twice (factor rows 0.. and 4..).
void XCIntegrator::eval_uks_(int64_t m, int64_t n, const double* Ps, int64_t ldps, const double* Pz, int64_t ldpz,
                             double* VXCs, int64_t ldvxcs, double* VXCz, int64_t ldvxcz, double* EXC,
                             bool do_vxc) {
  check_dims(*lb_, m, n, ldps, ldvxcs, do_vxc);
  check_dims(*lb_, m, n, ldpz, ldvxcz, do_vxc);
  if (!lb_->state().modified_weights_are_stored) GAUXC_GENERIC_EXCEPTION("Weights Have Not Been Modified");
  if (!func_->polarized) GAUXC_GENERIC_EXCEPTION("UKS Evaluation Requires A Polarized Functional");
  const bool gga = func_->is_gga();
  auto& I = *impl_;
  cudaStream_t s = I.stream;
  const size_t nbf = (size_t)m;
  const size_t nn = nbf * nbf;
  I.ensure_matrices(nbf, red_->comm_size(), true);
  if (!gga && I.dPtri.n != nn) {
    I.dPtri.alloc(nn);
    CUDA_CHECK(cudaMemsetAsync(I.dPtri.p, 0, sizeof(double) * nn, s));
  }
  if (!gga && I.dPtri_z.n != nn) {
    I.dPtri_z.alloc(nn);
    CUDA_CHECK(cudaMemsetAsync(I.dPtri_z.p, 0, sizeof(double) * nn, s));
  }
  CUDA_CHECK(cudaEventRecord(I.e_begin, s));
  CUDA_CHECK(cudaStreamWaitEvent(I.copy_stream, I.e_begin, 0));
  upload_density_(Ps, ldps, I.dP.p, nbf);
  upload_density_(Pz, ldpz, I.dPz.p, nbf);
  CUDA_CHECK(cudaEventRecord(I.e_p_ready, I.copy_stream));
  CUDA_CHECK(cudaStreamWaitEvent(s, I.e_p_ready, 0));
  const int nmat = gga ? 6 : 3;  // B [, dBx, dBy, dBz], Z_s, Z_z
  I.prepare(*lb_, nmat, !gga);
  auto& plan = *I.plan;
  const size_t nden = gga ? 4 : 1;  // rho_s (+ grad n) carried from the pass over Ps to the pass over Pz
  if (I.d_uks_den.n < nden * plan.npts) I.d_uks_den.alloc(std::max<size_t>(1, nden * plan.npts));
  auto& sc = *I.sched;
  const gxb::PlanView pv = plan.view();
  const int inbf = plan.nbf;
  const size_t nb = sc.batches.size();

  if (do_vxc) {
    CUDA_CHECK(cudaMemsetAsync(I.dVXC.p, 0, sizeof(double) * nn, s));
    CUDA_CHECK(cudaMemsetAsync(I.dVXCz.p, 0, sizeof(double) * nn, s));
  }
  CUDA_CHECK(cudaMemsetAsync(sc.d_counters.p, 0, sizeof(int) * sc.d_counters.n, s));
  CUDA_CHECK(cudaEventRecord(I.e_lw0, s));
  long long launches = 0;
  const double *dPs = I.dP.p, *dPzz = I.dPz.p;
  if (!gga) {  // LDA: the triangular quadratic form, as in the RKS path
    gxb::launch_sym_half(I.dP.p, inbf, I.dPtri.p, inbf, s);
    gxb::launch_sym_half(I.dPz.p, inbf, I.dPtri_z.p, inbf, s);
    dPs = I.dPtri.p;
    dPzz = I.dPtri_z.p;
    launches = 2;
  }
  size_t ib = 0;
  for (auto& b : sc.batches) {
    const int nt = b.tile_end - b.tile_begin;
    const gxb::DevTile* tl = sc.d_tiles.p + b.tile_begin;
    gxb::launch_collocation(pv, tl, nt, I.d_ws.p, gga, s);
    CUDA_CHECK(gxb::launch_fused(I.tmapA, pv, tl, nt, sc.d_counters.p + ib, sc.ncta, I.d_ws.p, dPs, inbf,
                                 func_->desc, I.d_exc_part.p, I.d_nel_part.p, b.tile_begin, s, 1, I.d_uks_den.p,
                                 plan.npts));
    CUDA_CHECK(gxb::launch_fused(I.tmapA, pv, tl, nt, sc.d_counters.p + nb + ib, sc.ncta, I.d_ws.p, dPzz, inbf,
                                 func_->desc, I.d_exc_part.p, I.d_nel_part.p, b.tile_begin, s, 2, I.d_uks_den.p,
                                 plan.npts));
    launches += 3;
    if (do_vxc) {
      CUDA_CHECK(gxb::launch_vxc(I.tmapV, I.tmapZ, pv, sc.d_items.p + b.item_begin, b.item_end - b.item_begin,
                                 sc.d_counters.p + 2 * nb + ib, sc.ncta, gga ? 4 : 1, nmat, !gga, I.dVXC.p, inbf,
                                 s));
      CUDA_CHECK(gxb::launch_vxc(I.tmapV, I.tmapZ, pv, sc.d_items.p + b.item_begin, b.item_end - b.item_begin,
                                 sc.d_counters.p + 3 * nb + ib, sc.ncta, gga ? 5 : 2, nmat, !gga, I.dVXCz.p,
                                 inbf, s));
      launches += 2;
    }
    ++ib;
  }
  gxb::launch_reduce_partials(I.d_exc_part.p, I.d_nel_part.p, (int)sc.tiles.size(), I.d_out2.p, s);
  ++launches;
  CUDA_CHECK(cudaGetLastError());
  CUDA_CHECK(cudaEventRecord(I.e_lw1, s));
  if (do_vxc) {
    reduce_and_symmetrize_(I.dVXC.p, I.dVXCz.p, I.d_out2.p, inbf, true);
    launches += 2;
    if (!vxc_root_only_ || lb_->runtime().comm_rank() == 0) {
      download_matrix(VXCs, ldvxcs, I.dVXC.p, nbf, s);
      download_matrix(VXCz, ldvxcz, I.dVXCz.p, nbf, s);
    }
  } else if (red_->comm_size() > 1) {
    red_->allreduce_inplace(I.d_out2.p, 2, ReductionOp::Sum, s);
  }
  CUDA_CHECK(cudaMemcpyAsync(I.h_out2, I.d_out2.p, 2 * sizeof(double), cudaMemcpyDeviceToHost, s));
  CUDA_CHECK(cudaEventRecord(I.e_end, s));
  CUDA_CHECK(cudaStreamSynchronize(s));
  float ms = 0;
  cudaEventElapsedTime(&ms, I.e_lw0, I.e_lw1);
  stats_.last_local_work_ms = ms;
  cudaEventElapsedTime(&ms, I.e_begin, I.e_end);
  stats_.last_total_ms = ms;
  timer_.add("XCIntegrator.LocalWork_EXC_VXC_UKS", stats_.last_local_work_ms);
  stats_.kernel_launches = launches;
  stats_.f_dense = 2. * plan.f_dense;
  stats_.sum_nbe_npts = plan.sum_nbe_npts;
  stats_.npts = (long long)plan.npts;
  stats_.ntiles = (long long)sc.tiles.size();
  stats_.nbatches = (long long)sc.batches.size();
  stats_.nitems = (long long)sc.items.size();
  *EXC = I.h_out2[0];
  stats_.n_el = I.h_out2[1];
}

void XCIntegrator::eval_exc_vxc_uks(int64_t m, int64_t n, const double* Ps, int64_t ldps, const double* Pz,
                                    int64_t ldpz, double* VXCs, int64_t ldvxcs, double* VXCz,
                                    int64_t ldvxcz, double* EXC) {
  eval_uks_(m, n, Ps, ldps, Pz, ldpz, VXCs, ldvxcs, VXCz, ldvxcz, EXC, true);
}
void XCIntegrator::eval_exc_uks(int64_t m, int64_t n, const double* Ps, int64_t ldps, const double* Pz,
                                int64_t ldpz, double* EXC) {
  eval_uks_(m, n, Ps, ldps, Pz, ldpz, nullptr, 0, nullptr, 0, EXC, false);
}

// EXC gradient, RKS LDA / GGA (incore_replicated_xc_device_integrator_exc_grad.hpp:21-72, 134-270; host semantics
// reference_replicated_xc_host_integrator_exc_grad.hpp:107-601).  Per batch: collocation gradient (LDA) / Hessian
// (GGA) -> X = 2 P_sub B on the DMMA pipe -> densities, functional [GGA: -> U = per-point combination of the dB ->
// Y = 2 P_sub U on the DMMA pipe] -> per-atom sums (exc_grad.cu: two contractions where the reference runs four);
// with weight derivatives (the default, IntegratorSettingsEXC_GRAD) one more kernel
// contracts the SSF weight derivatives with w eps rho over all local points.  The 3 natoms sums are reduced over the
// ranks on the device (the reference refuses a device reduction here: "Device Reduction + EXC Grad NYI").
void XCIntegrator::eval_exc_grad(int64_t m, int64_t n, const double* P, int64_t ldp, double* EXC_GRAD,
                                 bool include_weight_derivatives) {
  if (func_->polarized) GAUXC_GENERIC_EXCEPTION("RKS Evaluation Requires An Unpolarized Functional");
  eval_exc_grad_(m, n, P, ldp, nullptr, 0, EXC_GRAD, include_weight_derivatives);
}
// UKS: Ps = P_alpha + P_beta, Pz = P_alpha - P_beta (the (Ps, Pz) overload of eval_exc_grad_, :75-131); X of both
// densities with factor 1, the polarised functional, the vrho_n / vrho_z and four vgamma combinations of the host loop
void XCIntegrator::eval_exc_grad_uks(int64_t m, int64_t n, const double* Ps, int64_t ldps, const double* Pz,
                                     int64_t ldpz, double* EXC_GRAD, bool include_weight_derivatives) {
  if (!func_->polarized) GAUXC_GENERIC_EXCEPTION("UKS Evaluation Requires A Polarized Functional");
  if (!Pz) GAUXC_GENERIC_EXCEPTION("Invalid LDPZ");
  check_dims(*lb_, m, n, ldpz, 0, false);
  eval_exc_grad_(m, n, Ps, ldps, Pz, ldpz, EXC_GRAD, include_weight_derivatives);
}
void XCIntegrator::eval_exc_grad_(int64_t m, int64_t n, const double* P, int64_t ldp, const double* Pz, int64_t ldpz,
                                  double* EXC_GRAD, bool include_weight_derivatives) {
  check_dims(*lb_, m, n, ldp, 0, false);
  if (!lb_->state().modified_weights_are_stored) GAUXC_GENERIC_EXCEPTION("Weights Have Not Been Modified");
  if (include_weight_derivatives && lb_->state().weight_alg != XCWeightAlg::SSF)
    GAUXC_GENERIC_EXCEPTION("Weight Alg Not Supported");
  auto& I = *impl_;
  const size_t nbf = (size_t)m;
  cudaStream_t s = I.stream;
  const bool gga = func_->is_gga();
  const bool uks = Pz != nullptr;
  // tile matrices: LDA [B dB | X(s)], GGA [B dB ddB | X(s) | U(s) | Y(s) | F] (exc_grad.cu)
  const int nb = gga ? 10 : 4, nden = uks ? 2 : 1, nmat = gga ? nb + 3 * nden + 1 : nb + nden;
  I.ensure_matrices(nbf, red_->comm_size(), uks);
  CUDA_CHECK(cudaEventRecord(I.e_begin, s));
  CUDA_CHECK(cudaStreamWaitEvent(I.copy_stream, I.e_begin, 0));
  upload_density_(P, ldp, I.dP.p, nbf);
  if (uks) upload_density_(Pz, ldpz, I.dPz.p, nbf);
  CUDA_CHECK(cudaEventRecord(I.e_p_ready, I.copy_stream));
  I.prepare(*lb_, nmat, false);
  auto& plan = *I.plan;
  auto& sc = *I.sched;
  const gxb::PlanView pv = plan.view();
  const int inbf = plan.nbf, natoms = plan.natoms;
  if (I.d_grad.n != (size_t)3 * natoms) I.d_grad.alloc((size_t)3 * natoms);
  if (include_weight_derivatives && I.d_wf.n < plan.npts) I.d_wf.alloc(std::max<size_t>(1, plan.npts));
  CUDA_CHECK(cudaMemsetAsync(I.d_grad.p, 0, sizeof(double) * 3 * natoms, s));
  CUDA_CHECK(cudaMemsetAsync(sc.d_counters.p, 0, sizeof(int) * sc.d_counters.n, s));
  CUDA_CHECK(cudaEventRecord(I.e_lw0, s));
  const size_t nbatch = sc.batches.size();
  long long launches = 0;
  size_t ib = 0;
  for (auto& b : sc.batches) {
    const int nt = b.tile_end - b.tile_begin;
    const gxb::DevTile* tl = sc.d_tiles.p + b.tile_begin;
    if (gga) gxb::launch_collocation_hessian(pv, tl, nt, I.d_ws.p, s);
    else gxb::launch_collocation(pv, tl, nt, I.d_ws.p, true, s);
    if (ib == 0) CUDA_CHECK(cudaStreamWaitEvent(s, I.e_p_ready, 0));  // the upload of P hides behind the collocation
    // X = fac A P_sub on the DMMA pipe: matrix a_slot of the tile times the density d, written to x_slot; q: queue head
    auto xpass = [&](int q, int d, int a_slot, int x_slot) {
      CUDA_CHECK(gxb::launch_fused(I.tmapA, pv, tl, nt, sc.d_counters.p + q * nbatch + ib, sc.ncta, I.d_ws.p,
                                   d == 0 ? I.dP.p : I.dPz.p, inbf, func_->desc, I.d_exc_part.p, I.d_nel_part.p,
                                   b.tile_begin, s, 3, nullptr,
                                   (size_t)a_slot | ((size_t)x_slot << 16) | ((size_t)(uks ? 1 : 0) << 32)));
      ++launches;
    };
    auto gpass = [&](int q, int phase) {
      CUDA_CHECK(gxb::launch_exc_grad(pv, tl, nt, sc.d_counters.p + q * nbatch + ib, sc.ncta, I.d_ws.p, func_->desc, gga,
                                      uks, phase, plan.d_shell_center.p, natoms, include_weight_derivatives,
                                      include_weight_derivatives ? I.d_wf.p : nullptr, I.d_grad.p, s));
      ++launches;
    };
    for (int d = 0; d < nden; ++d) xpass(d, d, 0, nb + d);  // X(s) from B
    if (!gga) {
      gpass(8, 2);
    } else {
      gpass(8, 0);                                                               // densities, functional, U(s), F
      for (int d = 0; d < nden; ++d) xpass(2 + d, d, nb + nden + d, nb + 2 * nden + d);  // Y(s) from U(s)
      gpass(10, 1);                                                              // assembly
    }
    ++launches;  // collocation
    ++ib;
  }
  if (sc.batches.empty()) CUDA_CHECK(cudaStreamWaitEvent(s, I.e_p_ready, 0));
  if (include_weight_derivatives && !plan.tiles.empty()) {
    const cudaError_t e = gxb::launch_ssf_weight_grad(pv, plan.d_tiles.p, (int)plan.tiles.size(),
                                                      sc.d_counters.p + 9 * nbatch, sc.ncta, plan.d_atoms.p,
                                                      plan.d_rab.p, plan.d_dist_nearest.p, natoms, I.d_wf.p,
                                                      I.d_grad.p, s);
    if (e == cudaErrorInvalidConfiguration)
      GAUXC_GENERIC_EXCEPTION("SSF Weight Derivatives NYI in B200 path for this many atoms");
    CUDA_CHECK(e);
    ++launches;
  }
  CUDA_CHECK(cudaGetLastError());
  CUDA_CHECK(cudaEventRecord(I.e_lw1, s));
  if (red_->comm_size() > 1) red_->allreduce_inplace(I.d_grad.p, (size_t)3 * natoms, ReductionOp::Sum, s);
  CUDA_CHECK(cudaMemcpyAsync(EXC_GRAD, I.d_grad.p, sizeof(double) * 3 * natoms, cudaMemcpyDeviceToHost, s));
  CUDA_CHECK(cudaStreamSynchronize(s));
  float ms = 0;
  cudaEventElapsedTime(&ms, I.e_lw0, I.e_lw1);
  stats_.last_local_work_ms = ms;
  timer_.add("XCIntegrator.LocalWork_EXC_GRAD", ms);
  stats_.kernel_launches = launches;
  stats_.nbatches = (long long)sc.batches.size();
  stats_.ntiles = (long long)sc.tiles.size();
  stats_.npts = (long long)plan.npts;
}

}  // namespace GauXC

// ------------------------------------------------------------------------------------
// helpers behind the C ABI extensions
// ------------------------------------------------------------------------------------
namespace GauXC {

void device_set(int dev) {
  require_device();
  CUDA_CHECK(cudaSetDevice(dev));
}

int device_count() {
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess) {
    cudaGetLastError();
    return 0;
  }
  return n;
}

double device_probe_peak(int which) {
  require_device();
  switch (which) {
    case 0: return gxb::probe_dmma_tflops(20000);
    case 1: return gxb::probe_dfma_tflops(20000);
    case 2: return gxb::probe_copy_gbs((size_t)1 << 30, 10);
    default: GAUXC_GENERIC_EXCEPTION("Unknown probe");
  }
}

void device_allreduce(double* dptr, size_t n) {
  if (g_comm_size <= 1) return;
  NCCLReductionDriver d;
  d.allreduce_inplace(dptr, n, ReductionOp::Sum, nullptr);
  CUDA_CHECK(cudaStreamSynchronize(0));
}

// Collocation of one ad-hoc task on the device (unit tests against the golden collocation
// fixture, tests/collocation.cxx:45-91 in the reference).  Output in the host layout
// eval[ipt*nbe + mu] of gau2grid_collocation.cxx.
void device_eval_collocation(const BasisSet& basis, const std::vector<int32_t>& shell_list,
                             int64_t npts, const double* points, double* eval, double* dx,
                             double* dy, double* dz, double* hess6) {
  require_device();
  const bool grad = dx != nullptr;
  const bool hess = hess6 != nullptr;  // six more [npts][nbe] arrays: xx, xy, xz, yy, yz, zz
  Molecule mol;  // centres are irrelevant here
  BasisSetMap bmap(basis, mol);
  std::vector<gxb::DevShell> shells(basis.size());
  std::vector<double> alpha, coeff;
  for (size_t s = 0; s < basis.size(); ++s) {
    auto& sh = basis[s];
    gxb::DevShell d{};
    d.x = sh.O[0]; d.y = sh.O[1]; d.z = sh.O[2];
    d.l = sh.l; d.pure = sh.pure; d.nprim = sh.nprim;
    d.prim_off = (int)alpha.size();
    d.ao_off = bmap.shell_to_ao_range[s].first;
    d.nfunc = sh.size();
    for (int k = 0; k < sh.nprim; ++k) { alpha.push_back(sh.alpha[k]); coeff.push_back(sh.coeff[k]); }
    shells[s] = d;
  }
  gxb::DevTask task{};
  std::vector<int> tsh, tbf;
  int nbe = 0;
  for (int s : shell_list) {
    tsh.push_back(s);
    tbf.push_back(nbe);
    nbe += basis.at(s).size();
  }
  task.nshells = (int)shell_list.size();
  task.nbe = nbe;
  task.npts = (int)npts;
  std::vector<gxb::DevTile> tiles;
  const int nmat = hess ? 10 : (grad ? 4 : 1);
  for (int p0 = 0; p0 < npts; p0 += gxb::TP) {
    gxb::DevTile t{};
    t.task = 0; t.pt_off = p0; t.npts = (int)std::min<int64_t>(gxb::TP, npts - p0);
    t.nbe = nbe; t.ao_off = 0;
    t.ws_off = (int64_t)tiles.size() * nmat * gxb::pad16(nbe) * gxb::TP;
    tiles.push_back(t);
  }
  std::vector<double> px(npts), py(npts), pz(npts);
  for (int64_t i = 0; i < npts; ++i) { px[i] = points[3 * i]; py[i] = points[3 * i + 1]; pz[i] = points[3 * i + 2]; }
  DevBuf<gxb::DevShell> d_sh; DevBuf<double> d_a, d_c, d_px, d_py, d_pz, d_ws;
  DevBuf<gxb::DevTask> d_task; DevBuf<gxb::DevTile> d_tiles; DevBuf<int> d_tsh, d_tbf;
  d_sh.upload(shells); d_a.upload(alpha); d_c.upload(coeff);
  d_px.upload(px); d_py.upload(py); d_pz.upload(pz);
  d_task.upload(std::vector<gxb::DevTask>{task}); d_tiles.upload(tiles);
  d_tsh.upload(tsh); d_tbf.upload(tbf);
  const size_t wsn = tiles.size() * (size_t)nmat * gxb::pad16(nbe) * gxb::TP;
  d_ws.alloc(wsn);
  gxb::PlanView pv{};
  pv.shells = d_sh.p; pv.prim_alpha = d_a.p; pv.prim_coeff = d_c.p; pv.tasks = d_task.p;
  pv.task_shells = d_tsh.p; pv.task_shell_bf = d_tbf.p; pv.px = d_px.p; pv.py = d_py.p; pv.pz = d_pz.p;
  if (hess) gxb::launch_collocation_hessian(pv, d_tiles.p, (int)tiles.size(), d_ws.p, 0);
  else gxb::launch_collocation(pv, d_tiles.p, (int)tiles.size(), d_ws.p, grad, 0);
  CUDA_CHECK(cudaGetLastError());
  std::vector<double> h(wsn);
  CUDA_CHECK(cudaMemcpy(h.data(), d_ws.p, wsn * sizeof(double), cudaMemcpyDeviceToHost));
  const size_t ms = (size_t)gxb::pad16(nbe) * gxb::TP;
  for (size_t t = 0; t < tiles.size(); ++t)
    for (int i = 0; i < tiles[t].npts; ++i)
      for (int mu = 0; mu < nbe; ++mu) {
        const size_t src = (size_t)tiles[t].ws_off + (size_t)mu * gxb::TP + gxb::swz(mu, i);
        const size_t dst = (size_t)(tiles[t].pt_off + i) * nbe + mu;
        eval[dst] = h[src];
        if (grad) { dx[dst] = h[src + ms]; dy[dst] = h[src + 2 * ms]; dz[dst] = h[src + 3 * ms]; }
        if (hess)
          for (int q = 0; q < 6; ++q) hess6[(size_t)q * npts * nbe + dst] = h[src + (4 + q) * ms];
      }
}

}  // namespace GauXC
