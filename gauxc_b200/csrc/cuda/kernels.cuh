// Launch wrappers of the sm_100a kernels (definitions in *.cu next to this file).
#pragma once
#include "device_plan.hpp"
#include <cuda.h>
#include <cuda_runtime.h>

namespace gxb {

// TMA descriptors of the B^T operand of the fused kernel: boxes of 16 rows x (32, 64, 96, 128) points
struct TmapSet {
  CUtensorMap m[4];
};

// K_A  basis collocation (+gradient): writes B (dBx,dBy,dBz) [nbe][TP] per tile
void launch_collocation(const PlanView& pv, const DevTile* tiles, int ntiles, double* ws,
                        bool gradient, cudaStream_t s);

// K_F  fused persistent kernel: X = B P_sub (DMMA) -> rho / grad rho -> functional, weights,
//      EXC / N_EL tile partials -> Z.  tmapA: boxes of 16 rows x W points over the workspace;
//      the ncta persistent CTAs pull tile indices [0, ntiles) from *counter (zeroed by the caller).
//      func.nkern == 0: density only (integrate_den).  spin 1 / 2: the two UKS passes (Ps then Pz),
//      uks_den: 1 (LDA) or 4 (GGA) arrays of uks_stride doubles carried between them.
//      spin 3 (EXC gradient): X = 2 A P_sub only, A = matrix (uks_stride & 0xffff) of the tile, written to matrix
//      ((uks_stride >> 16) & 0xffff) of the tile (bit 32 of uks_stride: factor 1 instead of 2, UKS); no density,
//      functional or Z.
cudaError_t launch_fused(const TmapSet& tmapA, const PlanView& pv, const DevTile* tiles, int ntiles,
                         int* counter, int ncta, double* ws, const double* P, int ldp,
                         FunctionalDesc func, double* exc_part, double* nel_part, int part_off,
                         cudaStream_t s, int spin = 0, double* uks_den = nullptr, size_t uks_stride = 0);

// K_D  VXC_sub += B^T Z (+ transpose) on the DMMA pipe, scatter-added (lower triangle) into VXC.
//      tmapA / tmapZ: boxes of 128 / 64 rows x 16 points over the workspace; Z is matrix `zmat` of the `nmat`
//      matrices of a tile (RKS: LDA 1 of 2, GGA 4 of 5; UKS: Z_s / Z_z = 1 / 2 of 3 (LDA), 4 / 5 of 6 (GGA));
//      sym: M = B^T Z symmetric (LDA).  Two persistent CTAs per SM (nsm SMs) pull items [0, nitems) from
//      *counter (zeroed by the caller).
cudaError_t launch_vxc(const CUtensorMap& tmapA, const CUtensorMap& tmapZ, const PlanView& pv, const VxcItem* items,
                       int nitems, int* counter, int nsm, int zmat, int nmat, bool sym, double* VXC, int ldv,
                       cudaStream_t s);

// EXC gradient (exc_grad.cu)
//   Hessian collocation: ten matrices B, dx, dy, dz, xx, xy, xz, yy, yz, zz per tile
void launch_collocation_hessian(const PlanView& pv, const DevTile* tiles, int ntiles, double* ws, cudaStream_t s);
//   gradient kernel over tiles holding [B dx dy dz | X] (LDA; UKS: XN XZ; phase 2 = everything) or
//   [B dx dy dz xx xy xz yy yz zz | X | U | Y | F] (GGA; UKS: two of each X, U, Y): phase 0 evaluates densities +
//   functional and writes the factor rows F and U, phase 1 assembles with Y = fac U P_sub (exc_grad.cu);
//   shell_atom: shell -> atom; include_wd: skip the parent atom's shells, give the parent the opposite sum and
//   write wf_out[point] = w eps rho for the weight-derivative kernel; grad: 3 natoms, accumulated
cudaError_t launch_exc_grad(const PlanView& pv, const DevTile* tiles, int ntiles, int* counter, int nsm, double* ws,
                            FunctionalDesc func, bool gga, bool uks, int phase, const int* shell_atom, int natoms,
                            bool include_wd, double* wf_out, double* grad, cudaStream_t s);
//   SSF weight derivatives contracted with wf, accumulated into grad (cudaErrorInvalidConfiguration: too many atoms
//   for the shared-memory lists)
cudaError_t launch_ssf_weight_grad(const PlanView& pv, const DevTile* tiles, int ntiles, int* counter, int nsm,
                                   const double* atoms, const double* rab_inv, const double* dist_nearest, int natoms,
                                   const double* wf, double* grad, cudaStream_t s);

// finalisation
void launch_reduce_partials(const double* exc_part, const double* nel_part, int n, double* out2,
                            cudaStream_t s);
void launch_symmetrize(double* VXC, int nbf, int ldv, cudaStream_t s);
// lower triangle of A (n x n, column-major) -> packed n(n+1)/2 vector, and back with upper <- lower
void launch_pack_tril(const double* A, int n, int ld, double* out, cudaStream_t s);
void launch_unpack_tril_sym(const double* in, double* A, int n, int ld, cudaStream_t s);
// LDA density operand P' (lower triangle of (P + P^T)/2, diagonal halved), ld = nbf
void launch_sym_half(const double* P, int ldp, double* out, int nbf, cudaStream_t s);

// SSF weights (in place on pv.w)
//   rab / rab_inv [natoms][natoms]: R_AB and 1 / R_AB (0 on the diagonal);
//   nbr_idx / nbr_dist [natoms][natoms]: per atom, all atoms sorted by distance from it (itself first)
void launch_ssf_weights(const PlanView& pv, const DevTile* tiles, int ntiles, const double* atoms,
                        const double* rab, const double* rab_inv, const double* dist_nearest, const int* nbr_idx,
                        const double* nbr_dist, int natoms, cudaStream_t s);

// FP64 peak probes (DMMA m8n8k4 and DFMA), return achieved TFLOP/s
double probe_dmma_tflops(int iters);
double probe_dfma_tflops(int iters);
double probe_copy_gbs(size_t bytes, int iters);

}  // namespace gxb
