// Launch wrappers of the sm_100a kernels (definitions in *.cu next to this file).
#pragma once
#include "device_plan.hpp"
#include <cuda_runtime.h>

namespace gxb {

// K_A  basis collocation (+gradient): writes B (dBx,dBy,dBz) [nbe][TP] per tile
void launch_collocation(const PlanView& pv, const DevTile* tiles, int ntiles, double* ws,
                        bool gradient, cudaStream_t s);

// K_B  X = P_sub * B on the FP64 tensor pipe fused with rho / grad rho
void launch_xmat_density(const PlanView& pv, const DevTile* tiles, int ntiles, double* ws,
                         const double* P, int ldp, double* den, bool gga, cudaStream_t s);

// K_C  functional evaluation, weight scaling, EXC/N_EL partials, Z formation
void launch_func_zmat(const PlanView& pv, const DevTile* tiles, int ntiles, double* ws,
                      const double* den, FunctionalDesc func, double* exc_part,
                      double* nel_part, int part_off, cudaStream_t s);

// K_D  VXC_sub += B^T Z on the FP64 tensor pipe, scatter-added (lower triangle) into VXC
void launch_vxc(const PlanView& pv, const DevTile* tiles, const VxcItem* items, int nitems,
                const double* ws, bool gga, double* VXC, int ldv, cudaStream_t s);

// finalisation
void launch_reduce_partials(const double* exc_part, const double* nel_part, int n, double* out2,
                            cudaStream_t s);
void launch_symmetrize(double* VXC, int nbf, int ldv, cudaStream_t s);

// SSF weights (in place on pv.w)
void launch_ssf_weights(const PlanView& pv, const DevTile* tiles, int ntiles, const double* atoms,
                        const double* rab, const double* dist_nearest, int natoms,
                        cudaStream_t s);

// FP64 peak probes (DMMA m8n8k4 and DFMA), return achieved TFLOP/s
double probe_dmma_tflops(int iters);
double probe_dfma_tflops(int iters);
double probe_copy_gbs(size_t bytes, int iters);

}  // namespace gxb
