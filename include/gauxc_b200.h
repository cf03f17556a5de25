/* C ABI of libgauxc_b200.so -- the drop-in boundary for GauXC's Device execution space.
 *
 * Part 1 re-declares, name for name and argument for argument, the entry points of the
 * reference's C API (include/gauxc/c/ in the reference tree) that lie on the EXC/VXC hot
 * path, so that a program written against <gauxc/c/...> links against this library
 * unchanged.  Each block cites the reference header it replaces.
 * Part 2 (gauxc_b200_*) are extensions: rank/size and NCCL bootstrap without MPI,
 * device-resident evaluation, introspection used by the parity tests and the bench.
 *
 * Every call takes a GauXCStatus* first: code 0 = ok, 1 = error with a malloc'd message
 * (freed by the next call or gauxc_status_delete); with status == NULL an error is thrown
 * as a C++ exception, exactly like src/c-api/c_status.hpp:23-43.
 */
#ifndef GAUXC_B200_H
#define GAUXC_B200_H

#ifdef __cplusplus
#include <cstddef>
#include <cstdint>
extern "C" {
#else
#include <stdbool.h>
#include <stddef.h>
#include <stdint.h>
#endif

/* ---- include/gauxc/c/status.h:24-30 ------------------------------------------------ */
typedef struct GauXCStatus {
  int code;
  char* message;
} GauXCStatus;
void gauxc_status_delete(GauXCStatus* status);

/* ---- include/gauxc/c/types.h:26-78 -------------------------------------------------- */
enum GauXC_Type {
  GauXC_Type_Molecule = 1,
  GauXC_Type_BasisSet = 2,
  GauXC_Type_MolGrid = 3,
  GauXC_Type_RuntimeEnvironment = 4,
  GauXC_Type_LoadBalancer = 5,
  GauXC_Type_LoadBalancerFactory = 6,
  GauXC_Type_MolecularWeights = 7,
  GauXC_Type_MolecularWeightsFactory = 8,
  GauXC_Type_Functional = 9,
  GauXC_Type_Integrator = 10
};
typedef struct GauXCHeader {
  enum GauXC_Type type;
} GauXCHeader;
void gauxc_object_delete(GauXCStatus* status, void** handle);
void gauxc_objects_delete(GauXCStatus* status, void** handles, size_t nhandles);

/* ---- include/gauxc/c/enums.h ---------------------------------------------------------- */
enum GauXC_RadialQuad {
  GauXC_RadialQuad_Becke,
  GauXC_RadialQuad_MuraKnowles,
  GauXC_RadialQuad_MurrayHandyLaming,
  GauXC_RadialQuad_TreutlerAhlrichs
};
enum GauXC_AtomicGridSizeDefault {
  GauXC_AtomicGridSizeDefault_FineGrid,
  GauXC_AtomicGridSizeDefault_UltraFineGrid,
  GauXC_AtomicGridSizeDefault_SuperFineGrid,
  GauXC_AtomicGridSizeDefault_GM3,
  GauXC_AtomicGridSizeDefault_GM5,
  GauXC_AtomicGridSizeDefault_PySCF0,
  GauXC_AtomicGridSizeDefault_PySCF1,
  GauXC_AtomicGridSizeDefault_PySCF2,
  GauXC_AtomicGridSizeDefault_PySCF3,
  GauXC_AtomicGridSizeDefault_PySCF4,
  GauXC_AtomicGridSizeDefault_PySCF5,
  GauXC_AtomicGridSizeDefault_PySCF6,
  GauXC_AtomicGridSizeDefault_PySCF7,
  GauXC_AtomicGridSizeDefault_PySCF8,
  GauXC_AtomicGridSizeDefault_PySCF9
};
enum GauXC_XCWeightAlg {
  GauXC_XCWeightAlg_NOTPARTITIONED,
  GauXC_XCWeightAlg_Becke,
  GauXC_XCWeightAlg_SSF,
  GauXC_XCWeightAlg_LKO
};
enum GauXC_ExecutionSpace { GauXC_ExecutionSpace_Host, GauXC_ExecutionSpace_Device };
enum GauXC_PruningScheme {
  GauXC_PruningScheme_Unpruned,
  GauXC_PruningScheme_Robust,
  GauXC_PruningScheme_Treutler
};

/* ---- include/gauxc/c/atom.h, molecule.h ------------------------------------------------ */
typedef struct GauXCAtom {
  int64_t Z;
  double x, y, z; /* bohr */
} GauXCAtom;
typedef struct GauXCMolecule {
  GauXCHeader hdr;
  void* ptr;
} GauXCMolecule;
GauXCMolecule gauxc_molecule_new(GauXCStatus* status);
GauXCMolecule gauxc_molecule_new_from_atoms(GauXCStatus* status, const GauXCAtom* atoms, size_t natoms);
void gauxc_molecule_delete(GauXCStatus* status, GauXCMolecule* mol);
size_t gauxc_molecule_natoms(GauXCStatus* status, const GauXCMolecule mol);
bool gauxc_molecule_equal(GauXCStatus* status, const GauXCMolecule mol1, const GauXCMolecule mol2);

/* ---- include/gauxc/c/shell.h, basisset.h ------------------------------------------------ */
typedef struct GauXCShell {
  int32_t l;
  bool pure;
  int32_t nprim;
  double exponents[32];
  double coefficients[32];
  double origin[3];
  double shell_tolerance;
} GauXCShell;
typedef struct GauXCBasisSet {
  GauXCHeader hdr;
  void* ptr;
} GauXCBasisSet;
GauXCBasisSet gauxc_basisset_new(GauXCStatus* status);
GauXCBasisSet gauxc_basisset_new_from_shells(GauXCStatus* status, const GauXCShell* shells,
                                             size_t nshells, bool normalize);
void gauxc_basisset_delete(GauXCStatus* status, GauXCBasisSet* basis);

/* ---- include/gauxc/c/molgrid.h ------------------------------------------------------------ */
typedef struct GauXCMolGrid {
  GauXCHeader hdr;
  void* ptr;
} GauXCMolGrid;
GauXCMolGrid gauxc_molgrid_new_default(GauXCStatus* status, const GauXCMolecule mol,
                                       enum GauXC_PruningScheme pruning_scheme, int64_t batchsize,
                                       enum GauXC_RadialQuad radial_quad,
                                       enum GauXC_AtomicGridSizeDefault grid_size);
void gauxc_molgrid_delete(GauXCStatus* status, GauXCMolGrid* molgrid);

/* ---- include/gauxc/c/runtime_environment.h (GAUXC_HAS_DEVICE, no MPI) ---------------------- */
typedef struct GauXCRuntimeEnvironment {
  GauXCHeader hdr;
  void* ptr;
  void* device_ptr;
} GauXCRuntimeEnvironment;
GauXCRuntimeEnvironment gauxc_runtime_environment_new(GauXCStatus* status);
void gauxc_runtime_environment_delete(GauXCStatus* status, GauXCRuntimeEnvironment* env);
int gauxc_runtime_environment_comm_rank(GauXCStatus* status, const GauXCRuntimeEnvironment env);
int gauxc_runtime_environment_comm_size(GauXCStatus* status, const GauXCRuntimeEnvironment env);
GauXCRuntimeEnvironment gauxc_device_runtime_environment_new(GauXCStatus* status, double fill_fraction);
GauXCRuntimeEnvironment gauxc_device_runtime_environment_new_mem(GauXCStatus* status, void* mem, size_t mem_sz);

/* ---- include/gauxc/c/load_balancer.h ------------------------------------------------------- */
typedef struct GauXCLoadBalancer {
  GauXCHeader hdr;
  void* ptr;
} GauXCLoadBalancer;
typedef struct GauXCLoadBalancerFactory {
  GauXCHeader hdr;
  void* ptr;
} GauXCLoadBalancerFactory;
void gauxc_load_balancer_delete(GauXCStatus* status, GauXCLoadBalancer* lb);
GauXCLoadBalancerFactory gauxc_load_balancer_factory_new(GauXCStatus* status, enum GauXC_ExecutionSpace ex,
                                                         const char* kernel_name);
void gauxc_load_balancer_factory_delete(GauXCStatus* status, GauXCLoadBalancerFactory* factory);
GauXCLoadBalancer gauxc_load_balancer_factory_get_instance(GauXCStatus* status,
                                                           const GauXCLoadBalancerFactory factory,
                                                           const GauXCRuntimeEnvironment env,
                                                           const GauXCMolecule mol, const GauXCMolGrid mg,
                                                           const GauXCBasisSet basis);

/* ---- include/gauxc/c/molecular_weights.h ---------------------------------------------------- */
typedef struct GauXCMolecularWeightsSettings {
  enum GauXC_XCWeightAlg weight_alg;
  bool becke_size_adjustment;
} GauXCMolecularWeightsSettings;
typedef struct GauXCMolecularWeights {
  GauXCHeader hdr;
  void* ptr;
} GauXCMolecularWeights;
typedef struct GauXCMolecularWeightsFactory {
  GauXCHeader hdr;
  void* ptr;
} GauXCMolecularWeightsFactory;
void gauxc_molecular_weights_delete(GauXCStatus* status, GauXCMolecularWeights* mw);
void gauxc_molecular_weights_modify_weights(GauXCStatus* status, const GauXCMolecularWeights mw,
                                            const GauXCLoadBalancer lb);
GauXCMolecularWeightsFactory gauxc_molecular_weights_factory_new(GauXCStatus* status,
                                                                 enum GauXC_ExecutionSpace ex,
                                                                 const char* local_work_kernel_name,
                                                                 GauXCMolecularWeightsSettings settings);
void gauxc_molecular_weights_factory_delete(GauXCStatus* status, GauXCMolecularWeightsFactory* factory);
GauXCMolecularWeights gauxc_molecular_weights_factory_get_instance(GauXCStatus* status,
                                                                   const GauXCMolecularWeightsFactory factory);

/* ---- include/gauxc/c/functional.h (string constructor; LDA/GGA subset) ------- */
typedef struct GauXCFunctional {
  GauXCHeader hdr;
  void* ptr;
} GauXCFunctional;
GauXCFunctional gauxc_functional_from_string(GauXCStatus* status, const char* functional_spec, bool polarized);
/* include/gauxc/c/functional.h:22-330 (same enumerators, same order as ExchCXX::Functional); LDA/GGA members
 * SVWN5, BLYP, B3LYP, PBE, revPBE, PBE0, LDA, SPW92, VWN5, revPBE0 are built, the rest -> status 1 "NYI" */
enum GauXC_Functional {
  GauXC_Functional_SVWN3, GauXC_Functional_SVWN5, GauXC_Functional_BLYP, GauXC_Functional_B3LYP,
  GauXC_Functional_PBE, GauXC_Functional_revPBE, GauXC_Functional_PBE0, GauXC_Functional_SCAN,
  GauXC_Functional_R2SCAN, GauXC_Functional_R2SCANL, GauXC_Functional_M062X, GauXC_Functional_PKZB,
  GauXC_Functional_EPC17_1, GauXC_Functional_EPC17_2, GauXC_Functional_EPC18_1, GauXC_Functional_EPC18_2,
  GauXC_Functional_B97D, GauXC_Functional_B97D3ZERO, GauXC_Functional_CAMB3LYP, GauXC_Functional_LDA,
  GauXC_Functional_M06L, GauXC_Functional_SCAN0, GauXC_Functional_SPW92, GauXC_Functional_TPSS,
  GauXC_Functional_TPSSh, GauXC_Functional_TPSS0, GauXC_Functional_VWN3, GauXC_Functional_VWN5,
  GauXC_Functional_LRCwPBE, GauXC_Functional_LRCwPBEh, GauXC_Functional_BP86, GauXC_Functional_HSE03,
  GauXC_Functional_HSE06, GauXC_Functional_revB3LYP, GauXC_Functional_revPBE0, GauXC_Functional_revTPSS,
  GauXC_Functional_revTPSSh, GauXC_Functional_PW91, GauXC_Functional_mBEEF, GauXC_Functional_B3PW91,
  GauXC_Functional_O3LYP, GauXC_Functional_OLYP, GauXC_Functional_OPBE, GauXC_Functional_MPW1K,
  GauXC_Functional_RPBE, GauXC_Functional_B88, GauXC_Functional_MPW91, GauXC_Functional_RSCAN,
  GauXC_Functional_TUNEDCAMB3LYP, GauXC_Functional_wB97, GauXC_Functional_wB97X, GauXC_Functional_wB97XD,
  GauXC_Functional_wB97XD3, GauXC_Functional_LCwPBE, GauXC_Functional_X3LYP, GauXC_Functional_XLYP,
  GauXC_Functional_BHANDH, GauXC_Functional_BMK, GauXC_Functional_BP86VWN, GauXC_Functional_PW86B95,
  GauXC_Functional_PW86PBE, GauXC_Functional_R2SCAN0, GauXC_Functional_R2SCANh, GauXC_Functional_R2SCAN50,
  GauXC_Functional_M05, GauXC_Functional_M06, GauXC_Functional_M08HX, GauXC_Functional_M08SO,
  GauXC_Functional_M052X, GauXC_Functional_M06SX, GauXC_Functional_CF22D, GauXC_Functional_SOGGA11X,
  GauXC_Functional_M06HF, GauXC_Functional_M11, GauXC_Functional_MN12L, GauXC_Functional_MN12SX,
  GauXC_Functional_MN15, GauXC_Functional_MN15L, GauXC_Functional_revM06L
};
GauXCFunctional gauxc_functional_from_enum(GauXCStatus* status, enum GauXC_Functional functional_type,
                                           bool polarized);
void gauxc_functional_delete(GauXCStatus* status, GauXCFunctional* functional);

/* ---- include/gauxc/c/xc_integrator.h:34-190 -------------------------------------------------- */
typedef struct GauXCIntegrator {
  GauXCHeader hdr;
  void* ptr;
} GauXCIntegrator;
void gauxc_integrator_delete(GauXCStatus* status, GauXCIntegrator* integrator);
GauXCIntegrator gauxc_integrator_new(GauXCStatus* status, const GauXCFunctional functional,
                                     const GauXCLoadBalancer lb, enum GauXC_ExecutionSpace execution_space,
                                     const char* integrator_input_type, const char* integrator_kernel_name,
                                     const char* local_work_kernel_name, const char* reduction_kernel_name);
void gauxc_integrator_integrate_den(GauXCStatus* status, const GauXCIntegrator integrator, const int64_t m,
                                    const int64_t n, const double* density_matrix, const int64_t ldp,
                                    double* den);
void gauxc_integrator_eval_exc_rks(GauXCStatus* status, const GauXCIntegrator integrator, const int64_t m,
                                   const int64_t n, const double* density_matrix, const int64_t ldp,
                                   double* exc);
void gauxc_integrator_eval_exc_vxc_rks(GauXCStatus* status, const GauXCIntegrator integrator,
                                       const int64_t m, const int64_t n, const double* density_matrix,
                                       const int64_t ldp, double* exc, double* vxc_matrix,
                                       const int64_t vxc_ld);
/* UKS: functionals built with polarized = true; density_matrix_s = P_alpha + P_beta,
 * density_matrix_z = P_alpha - P_beta. */
void gauxc_integrator_eval_exc_vxc_uks(GauXCStatus* status, const GauXCIntegrator integrator,
                                       const int64_t m, const int64_t n, const double* density_matrix_s,
                                       const int64_t ldp_s, const double* density_matrix_z,
                                       const int64_t ldp_z, double* exc, double* vxc_matrix_s,
                                       const int64_t vxc_ld_s, double* vxc_matrix_z, const int64_t vxc_ld_z);
void gauxc_integrator_eval_exc_grad_rks(GauXCStatus* status, const GauXCIntegrator integrator,
                                        const int64_t m, const int64_t n, const double* density_matrix,
                                        const int64_t ldp, double* exc_grad);
void gauxc_integrator_eval_exc_uks(GauXCStatus* status, const GauXCIntegrator integrator, const int64_t m,
                                   const int64_t n, const double* density_matrix_s, const int64_t ldp_s,
                                   const double* density_matrix_z, const int64_t ldp_z, double* exc);
/* include/gauxc/c/xc_integrator.h:304-314: UKS EXC gradient (default settings: weight derivatives included) */
void gauxc_integrator_eval_exc_grad_uks(GauXCStatus* status, const GauXCIntegrator integrator, const int64_t m,
                                        const int64_t n, const double* density_matrix_s, const int64_t ldp_s,
                                        const double* density_matrix_z, const int64_t ldp_z, double* exc_grad);
/* include/gauxc/c/xc_integrator.h:124-365, outside the LDA/GGA RKS/UKS path: exported for link
 * compatibility, every call returns status code 1 with a "... NYI in B200 path" message. */
void gauxc_integrator_eval_exc_gks(GauXCStatus* status, const GauXCIntegrator integrator, const int64_t m,
                                   const int64_t n, const double* density_matrix_s, const int64_t ldp_s,
                                   const double* density_matrix_z, const int64_t ldp_z,
                                   const double* density_matrix_y, const int64_t ldp_y,
                                   const double* density_matrix_x, const int64_t ldp_x, double* exc);
void gauxc_integrator_eval_exc_vxc_gks(GauXCStatus* status, const GauXCIntegrator integrator, const int64_t m,
                                       const int64_t n, const double* density_matrix_s, const int64_t ldp_s,
                                       const double* density_matrix_z, const int64_t ldp_z,
                                       const double* density_matrix_y, const int64_t ldp_y,
                                       const double* density_matrix_x, const int64_t ldp_x, double* exc,
                                       double* vxc_matrix_s, const int64_t vxc_ld_s, double* vxc_matrix_z,
                                       const int64_t vxc_ld_z, double* vxc_matrix_y, const int64_t vxc_ld_y,
                                       double* vxc_matrix_x, const int64_t vxc_ld_x);
void gauxc_integrator_eval_exx_rks(GauXCStatus* status, const GauXCIntegrator integrator, const int64_t m,
                                   const int64_t n, const double* density_matrix, const int64_t ldp, double* K,
                                   const int64_t ldk);
void gauxc_integrator_eval_fxc_contraction_rks(GauXCStatus* status, const GauXCIntegrator integrator,
                                               const int64_t m, const int64_t n, const double* density_matrix,
                                               const int64_t ldp, const double* t_density_matrix,
                                               const int64_t ltdp, double* fxc, const int64_t ldfxc);
void gauxc_integrator_eval_fxc_contraction_uks(GauXCStatus* status, const GauXCIntegrator integrator,
                                               const int64_t m, const int64_t n, const double* density_matrix_s,
                                               const int64_t ldp_s, const double* density_matrix_z,
                                               const int64_t ldp_z, const double* t_density_matrix_s,
                                               const int64_t ldtp_s, const double* t_density_matrix_z,
                                               const int64_t ldtp_z, double* fxc_s, const int64_t ldfxc_s,
                                               double* fxc_z, const int64_t ldfxc_z);

/* ---- include/gauxc/c/hdf5.h:24-52 (HDF5 records of the reference's fixtures; self-contained reader /
 * writer, no libhdf5) -------------------------------------------------------------------------- */
void gauxc_molecule_write_hdf5_record(GauXCStatus* status, GauXCMolecule mol, const char* fname, const char* dset);
void gauxc_basisset_write_hdf5_record(GauXCStatus* status, GauXCBasisSet basis, const char* fname,
                                      const char* dset);
void gauxc_molecule_read_hdf5_record(GauXCStatus* status, GauXCMolecule mol, const char* fname, const char* dset);
void gauxc_basisset_read_hdf5_record(GauXCStatus* status, GauXCBasisSet basis, const char* fname, const char* dset);

/* =============================== Part 2: extensions ======================================== */

/* Rank / size of this process in the job (replaces the MPI_Comm argument of the reference's
 * runtime constructors; there is no MPI here). */
void gauxc_b200_runtime_environment_set_comm(GauXCStatus* status, GauXCRuntimeEnvironment env, int rank,
                                             int size);
/* NCCL bootstrap without MPI (the reference broadcasts the id with MPI_Bcast,
 * src/reduction_driver/device/nccl_reduction_driver.hpp:24-34): rank 0 obtains 128 id bytes,
 * the launcher broadcasts them, every rank calls init. */
void gauxc_b200_nccl_get_unique_id(GauXCStatus* status, char id[128]);
void gauxc_b200_nccl_init(GauXCStatus* status, const char id[128], int rank, int size);
void gauxc_b200_nccl_finalize(GauXCStatus* status);
/* in-place sum over ranks of a device buffer through the integrator's reduction driver
 * (ReductionDriver::allreduce_inplace, include/gauxc/reduction_driver.hpp:47-61) */
void gauxc_b200_allreduce_device(GauXCStatus* status, double* dptr, size_t n);

/* Device-resident EXC/VXC: dP (nbf x nbf, ld = nbf) and dVXC are device pointers, d_out2 a
 * device array of 2 doubles {EXC, N_EL}; no host<->device copies inside. */
void gauxc_b200_integrator_eval_exc_vxc_rks_device(GauXCStatus* status, const GauXCIntegrator integrator,
                                                   const double* dP, double* dVXC, double* d_out2);

/* Dense FP64 datasets of the reference's HDF5 fixtures (/DENSITY, /VXC, /EXC, ...): size query (returns the element
 * count, fills up to 4 dims in file order), read, and append-to-file write (root group, contiguous). */
int64_t gauxc_b200_hdf5_dataset_size(GauXCStatus* status, const char* fname, const char* dset, int64_t* dims4,
                                     int* rank);
void gauxc_b200_hdf5_read_dataset(GauXCStatus* status, const char* fname, const char* dset, double* out, int64_t n);
void gauxc_b200_hdf5_write_dataset(GauXCStatus* status, const char* fname, const char* dset, const double* data,
                                   const int64_t* dims, int rank);
void gauxc_b200_molecule_get_atoms(GauXCStatus* status, const GauXCMolecule mol, GauXCAtom* atoms);

/* Introspection (tests / bench). */
int64_t gauxc_b200_basisset_nbf(GauXCStatus* status, const GauXCBasisSet basis);
int64_t gauxc_b200_basisset_nshells(GauXCStatus* status, const GauXCBasisSet basis);
void gauxc_b200_basisset_set_shell_tolerance(GauXCStatus* status, GauXCBasisSet basis, double tol);
/* shell s -> {l, pure, nprim, cutoff_radius, origin[3], alpha[32], coeff[32]} */
void gauxc_b200_basisset_get_shell(GauXCStatus* status, const GauXCBasisSet basis, int64_t s, int32_t* l,
                                   int32_t* pure, int32_t* nprim, double* cutoff, double* origin,
                                   double* alpha, double* coeff);
int64_t gauxc_b200_load_balancer_ntasks(GauXCStatus* status, const GauXCLoadBalancer lb);
int64_t gauxc_b200_load_balancer_total_npts(GauXCStatus* status, const GauXCLoadBalancer lb);
/* per task: iParent, npts, nbe, nshells, dist_nearest (arrays of length ntasks) */
void gauxc_b200_load_balancer_task_info(GauXCStatus* status, const GauXCLoadBalancer lb, int32_t* iParent,
                                        int32_t* npts, int32_t* nbe, int32_t* nshells, double* dist_nearest);
/* LoadBalancer::state() (include/gauxc/load_balancer.hpp:37-48): weights already partitioned? by which scheme? */
void gauxc_b200_load_balancer_state(GauXCStatus* status, const GauXCLoadBalancer lb, int* modified_weights_are_stored,
                                    int* weight_alg);
/* copy one task out (points npts x 3 row-major, weights npts, shell_list nshells) */
void gauxc_b200_load_balancer_get_task(GauXCStatus* status, const GauXCLoadBalancer lb, int64_t itask,
                                       double* points, double* weights, int32_t* shell_list);
/* overwrite a task's weights (marks device copies stale) */
void gauxc_b200_load_balancer_set_task_weights(GauXCStatus* status, GauXCLoadBalancer lb, int64_t itask,
                                               const double* weights);
/* replace the whole task list by user supplied tasks (parity tests on golden fixtures):
 * concatenated points/weights, per task npts/iParent/dist_nearest and shell lists */
void gauxc_b200_load_balancer_set_tasks(GauXCStatus* status, GauXCLoadBalancer lb, int64_t ntasks,
                                        const int32_t* npts, const int32_t* iParent,
                                        const double* dist_nearest, const double* points,
                                        const double* weights, const int32_t* nshells,
                                        const int32_t* shell_lists, int weights_are_modified);
/* stats of the last eval call: out[0..15] = {local_work_ms, total_ms, k_colloc_ms, k_xmat_ms,
 * k_zmat_ms, k_vxc_ms, launches, f_dense, sum_nbe_npts, npts, ntiles, nbatches, nitems, n_el,
 * 0, 0}; per-kernel ms are only filled in profile mode */
/* EXC gradient with the reference's IntegratorSettingsEXC_GRAD::include_weight_derivatives
 * (include/gauxc/xc_integrator_settings.hpp:28-30; the C entry point gauxc_integrator_eval_exc_grad_rks has no
 * settings argument and uses the default, true).  exc_grad: 3 * natoms doubles, atom-major. */
void gauxc_b200_integrator_eval_exc_grad_rks(GauXCStatus* status, const GauXCIntegrator integrator, const int64_t m,
                                             const int64_t n, const double* density_matrix, const int64_t ldp,
                                             double* exc_grad, int include_weight_derivatives);
void gauxc_b200_integrator_eval_exc_grad_uks(GauXCStatus* status, const GauXCIntegrator integrator, const int64_t m,
                                             const int64_t n, const double* density_matrix_s, const int64_t ldp_s,
                                             const double* density_matrix_z, const int64_t ldp_z, double* exc_grad,
                                             int include_weight_derivatives);
void gauxc_b200_integrator_stats(GauXCStatus* status, const GauXCIntegrator integrator, double* out16);
void gauxc_b200_integrator_set_profile(GauXCStatus* status, const GauXCIntegrator integrator, int on);
/* extension: only rank 0 copies VXC back to host memory (default off = the reference's replicated result) */
void gauxc_b200_integrator_set_vxc_root_only(GauXCStatus* status, const GauXCIntegrator integrator, int on);
double gauxc_b200_molecular_weights_last_ms(GauXCStatus* status, const GauXCMolecularWeights mw);
/* Lebedev / radial tables of the grid generator (tests) */
int64_t gauxc_b200_lebedev(GauXCStatus* status, int npts, double* xyz, double* w);
void gauxc_b200_radial(GauXCStatus* status, enum GauXC_RadialQuad rq, int n, double R, double* r, double* w);
/* device collocation of arbitrary points (tests): eval[nbf_list x npts] (+ gradients if
 * deval_x != NULL), function index fastest like the host layout of the golden file */
void gauxc_b200_eval_collocation(GauXCStatus* status, const GauXCBasisSet basis, int64_t nshells,
                                 const int32_t* shell_list, int64_t npts, const double* points,
                                 double* eval, double* deval_x, double* deval_y, double* deval_z);
/* the same with second derivatives (the EXC gradient's collocation): hess6 = six more [npts][nbe] arrays
 * xx, xy, xz, yy, yz, zz (gau2grid_collocation_hessian, .../host/reference/gau2grid_collocation.cxx:153-216) */
void gauxc_b200_eval_collocation_hessian(GauXCStatus* status, const GauXCBasisSet basis, int64_t nshells,
                                         const int32_t* shell_list, int64_t npts, const double* points, double* eval,
                                         double* dx, double* dy, double* dz, double* hess6);
/* product functional evaluated on the HOST for unit tests of the formulas only (never used
 * by the integrator): out = {eps, vrho, vsigma} per point */
void gauxc_b200_functional_eval_host(GauXCStatus* status, const GauXCFunctional functional, int64_t npts,
                                     const double* rho, const double* sigma, double* eps, double* vrho,
                                     double* vsigma);
/* same for the spin-polarised LDA kernels of the UKS path: out = {eps, vrho_a, vrho_b} per point */
void gauxc_b200_functional_eval_host_pol(GauXCStatus* status, const GauXCFunctional functional, int64_t npts,
                                         const double* rho_a, const double* rho_b, double* eps, double* vrho_a,
                                         double* vrho_b);
/* the functional's spin-polarised evaluation as the UKS GGA path of the fused kernel performs it (host evaluation
 * for unit tests): rho2 = {rho_a, rho_b}, gamma3 = {sigma_aa, sigma_ab, sigma_bb} per point */
void gauxc_b200_functional_eval_host_pol_full(GauXCStatus* status, const GauXCFunctional functional, int64_t npts,
                                              const double* rho2, const double* gamma3, double* eps, double* vrho2,
                                              double* vgamma3);
/* spin-polarised GGA kernels prepared for the UKS GGA path (host evaluation for unit tests; kern: 0 = B88
 * exchange, 1 = LYP correlation; rho2 = {rho_a, rho_b}, gamma3 = {sigma_aa, sigma_ab, sigma_bb} per point) */
void gauxc_b200_functional_eval_host_pol_gga(GauXCStatus* status, int nkern, const int* kern, const double* coeff,
                                             int64_t npts, const double* rho2, const double* gamma3, double* eps,
                                             double* vrho2, double* vgamma3);
/* FP64 machine-peak probes (roofline denominators): which = 0 DMMA TF/s, 1 DFMA TF/s, 2 HBM copy GB/s */
double gauxc_b200_probe_peak(GauXCStatus* status, int which);
int gauxc_b200_device_count(void);
/* select the CUDA device of the calling thread (one process per GPU: LOCAL_RANK) */
void gauxc_b200_set_device(GauXCStatus* status, int device);
/* the cudaStream_t every kernel of this integrator is launched on (for CUDA-event timing) */
void* gauxc_b200_integrator_stream(GauXCStatus* status, const GauXCIntegrator integrator);
const char* gauxc_b200_version(void);

#ifdef __cplusplus
}
#endif
#endif
