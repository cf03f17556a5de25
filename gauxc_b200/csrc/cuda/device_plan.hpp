// Device-resident work description shared by the host integrator and the sm_100a kernels.
//
// Layout in HBM (all FP64 unless noted), replacing XCDeviceData / XCDeviceTask
// (src/xc_integrator/xc_data/device/xc_device_task.hpp:16-242):
//   * points/weights: SoA px,py,pz,w over ALL local points, ordered task by task
//     (tasks sorted by npts*nbe descending like the reference,
//      incore_replicated_xc_device_integrator_exc_vxc.hpp:254-257).  Resident across calls.
//   * a task is cut into TILES of <= TP consecutive points that share the task's shell list.
//   * per batch of tiles a workspace holds, per tile, NMAT matrices [nbp][TP] (nbp = nbe
//     rounded up to 16 rows, pad rows are zero): B, (dBx,dBy,dBz for GGA), Z (UKS: Z_s, Z_z).  Rows are
//     1 KB (point index fastest) and XOR-swizzled: element (row, i) lives at column
//     i ^ ((row & 3) << 2).  With that one permutation a TMA box of 16 rows x 128 points (the
//     A operand of X = B P) and TMA boxes of 128 / 64 rows x 16 points (the operands of B^T Z)
//     land in shared memory dense AND conflict-free for the m8n8k4 DMMA fragment loads,
//     while every global row access stays fully coalesced.
#pragma once
#include <cstdint>

#if defined(__CUDACC__)
#define GXB_HOST_DEVICE __host__ __device__
#else
#define GXB_HOST_DEVICE
#endif

namespace gxb {

constexpr int TP = 128;  // points per tile

GXB_HOST_DEVICE inline int pad16(int nbe) { return (nbe + 15) & ~15; }
// columns of a tile that are ever written / read: npts rounded up to 32
GXB_HOST_DEVICE inline int tile_width(int npts) { return (npts + 31) & ~31; }
GXB_HOST_DEVICE inline int swz(int row, int i) { return i ^ ((row & 3) << 2); }
// workspace rows of one tile: nmat matrices of pad16(nbe) rows
GXB_HOST_DEVICE inline int tile_rows(int nmat, int nbe) { return nmat * pad16(nbe); }

struct DevShell {
  double x, y, z;
  int l, pure, nprim, prim_off;  // prim_off: offset into the alpha/coeff arrays
  int ao_off, nfunc;             // first global AO, functions in this shell
  int pad0, pad1;
};

struct DevTask {
  int shell_off, nshells;  // into task_shells / task_shell_bf
  int ao_off, nbe;         // into task_ao (local mu -> global AO)
  int pt_off, npts;        // into the point arrays
  int iParent, pad;
};

struct DevTile {
  int task;
  int pt_off;  // global point offset of the tile's first point
  int npts;    // <= TP
  int nbe;     // copy of the task's nbe / ao_off: the kernels need no dependent DevTask load
  int64_t ws_off;  // offset (in doubles) of the tile's matrices inside the batch workspace
  int ao_off;
  int pad;
};

// one unit of the VXC rank update: output block (mblk, nblk) of VXC_BLK x VXC_BLN of one task,
// accumulated over a run of `ntiles` consecutive tiles of that task in the current batch.
// Self-contained (32 bytes, one load): the tiles of a task lie back to back in the workspace (row
// stride tile_rows(nmat, nbe)), all full (TP points = TP/16 K steps) except possibly the last one
// (nks_last K steps).
struct VxcItem {
  int nbe, ao_off, mblk, nblk;
  int row0, ntiles, nks_last, pad;  // row0: workspace row (ws_off / TP) of the first tile
};
constexpr int VXC_BLK = 128;  // output block rows (mu) of the VXC rank update
constexpr int VXC_BLN = 64;   // output block columns (nu)

struct PlanView {
  // static
  const DevShell* shells;
  const double* prim_alpha;
  const double* prim_coeff;
  const DevTask* tasks;
  const int* task_shells;    // global shell index
  const int* task_shell_bf;  // local AO offset of each listed shell
  const int* task_ao;        // local mu -> global AO index
  const double *px, *py, *pz;
  double* w;
  int nbf;
};

enum XcKind : int { XC_LDA = 0, XC_GGA = 1 };

// functional = sum_k coeff_k * kernel_k  (ExchCXX XCFunctional semantics)
enum KernelId : int {
  K_SLATER_X = 0, K_VWN5_C = 1, K_PBE_X = 2, K_PBE_C = 3, K_VWN3_C = 4, K_PW92_C = 5,
  K_B88_X = 6, K_LYP_C = 7, K_REVPBE_X = 8
};
struct FunctionalDesc {
  int nkern;
  int is_gga;
  int kern[4];
  double coeff[4];
};

}  // namespace gxb
