/* Shim of the reference's generated <gauxc/c/gauxc_config.h>: what this build provides. */
#pragma once
#define GAUXC_HAS_C 1
#define GAUXC_HAS_DEVICE 1
#define GAUXC_HAS_CUDA 1
#define GAUXC_HAS_NCCL 1
#define GAUXC_HAS_HDF5 1
/* no MPI in this build: rank / size come from the launcher (gauxc_b200_runtime_environment_set_comm) */
