/* A plain C program written against the C ABI only (include/gauxc_b200.h), the way a host code uses
 * the reference's <gauxc/c/...> headers: water, a small s/p basis,
 * UltraFine MuraKnowles grid, SSF weights on the Device, PBE EXC + VXC with host buffers.
 *
 * prints   NO_DEVICE <message>            when the library reports that there is no CUDA device
 *          EXC <exc>  NEL <2 * integrate_den>  NBF <nbf>  VXC <nbf*nbf values>     otherwise
 * exit code 0 in both cases, 1 on any other error.  Built and run by tests/test_c_client.py. */
#include <gauxc_b200.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

static int failed(GauXCStatus* st, const char* what) {
  if (st->code == 0) return 0;
  if (st->message && strstr(st->message, "No CUDA device")) {
    printf("NO_DEVICE %s\n", st->message);
    exit(0);
  }
  fprintf(stderr, "%s failed: %s\n", what, st->message ? st->message : "?");
  return 1;
}

static GauXCShell shell(int l, int nprim, const double* a, const double* c, const GauXCAtom* at) {
  GauXCShell s;
  memset(&s, 0, sizeof(s));
  s.l = l; s.pure = true; s.nprim = nprim;
  for (int k = 0; k < nprim; ++k) { s.exponents[k] = a[k]; s.coefficients[k] = c[k]; }
  s.origin[0] = at->x; s.origin[1] = at->y; s.origin[2] = at->z;
  s.shell_tolerance = 1e-10;
  return s;
}

int main(void) {
  GauXCStatus st = {0, NULL};
  const GauXCAtom atoms[3] = {{8, 0., -0.07579, 0.}, {1, 0.86681, 0.60144, 0.}, {1, -0.86681, 0.60144, 0.}};
  /* STO-3G-like contractions (any normalisable set will do: the test compares two bindings of the
   * same library on the same input) */
  const double a_o1[3] = {130.70932, 23.808861, 6.4436083}, c_s[3] = {0.15432897, 0.53532814, 0.44463454};
  const double a_o2[3] = {5.0331513, 1.1695961, 0.3803890}, c_2s[3] = {-0.09996723, 0.39951283, 0.70011547};
  const double c_2p[3] = {0.15591627, 0.60768372, 0.39195739};
  const double a_h[3] = {3.42525091, 0.62391373, 0.16885540};
  GauXCShell shells[5];
  shells[0] = shell(0, 3, a_o1, c_s, &atoms[0]);
  shells[1] = shell(0, 3, a_o2, c_2s, &atoms[0]);
  shells[2] = shell(1, 3, a_o2, c_2p, &atoms[0]);
  shells[3] = shell(0, 3, a_h, c_s, &atoms[1]);
  shells[4] = shell(0, 3, a_h, c_s, &atoms[2]);

  GauXCMolecule mol = gauxc_molecule_new_from_atoms(&st, atoms, 3);
  if (failed(&st, "molecule")) return 1;
  GauXCBasisSet basis = gauxc_basisset_new_from_shells(&st, shells, 5, true);
  if (failed(&st, "basisset")) return 1;
  GauXCMolGrid mg = gauxc_molgrid_new_default(&st, mol, GauXC_PruningScheme_Unpruned, 512,
                                              GauXC_RadialQuad_MuraKnowles,
                                              GauXC_AtomicGridSizeDefault_UltraFineGrid);
  if (failed(&st, "molgrid")) return 1;
  GauXCRuntimeEnvironment rt = gauxc_device_runtime_environment_new(&st, 0.5);
  if (failed(&st, "runtime")) return 1;
  GauXCLoadBalancerFactory lbf = gauxc_load_balancer_factory_new(&st, GauXC_ExecutionSpace_Host, "Default");
  if (failed(&st, "lb factory")) return 1;
  GauXCLoadBalancer lb = gauxc_load_balancer_factory_get_instance(&st, lbf, rt, mol, mg, basis);
  if (failed(&st, "load balancer")) return 1;
  const long long nbf = (long long)gauxc_b200_basisset_nbf(&st, basis);
  const long long npts = (long long)gauxc_b200_load_balancer_total_npts(&st, lb);
  if (failed(&st, "introspection")) return 1;
  fprintf(stderr, "nbf %lld, %lld grid points, %lld tasks\n", nbf, npts,
          (long long)gauxc_b200_load_balancer_ntasks(&st, lb));

  GauXCMolecularWeightsSettings ws = {GauXC_XCWeightAlg_SSF, false};
  GauXCMolecularWeightsFactory mwf =
      gauxc_molecular_weights_factory_new(&st, GauXC_ExecutionSpace_Device, "Default", ws);
  if (failed(&st, "weights factory")) return 1;
  GauXCMolecularWeights mw = gauxc_molecular_weights_factory_get_instance(&st, mwf);
  if (failed(&st, "weights")) return 1;
  gauxc_molecular_weights_modify_weights(&st, mw, lb);
  if (failed(&st, "modify_weights")) return 1; /* CPU-only box: NO_DEVICE, exit 0 */

  GauXCFunctional f = gauxc_functional_from_string(&st, "PBE", false);
  if (failed(&st, "functional")) return 1;
  GauXCIntegrator integ = gauxc_integrator_new(&st, f, lb, GauXC_ExecutionSpace_Device, "Replicated",
                                               "Default", "Default", "Default");
  if (failed(&st, "integrator")) return 1;

  double* P = (double*)calloc((size_t)(nbf * nbf), sizeof(double));
  double* V = (double*)calloc((size_t)(nbf * nbf), sizeof(double));
  /* P_alpha: occupied-looking diagonal plus a small symmetric coupling */
  const double occ[7] = {1.0, 0.9, 0.7, 0.7, 0.7, 0.3, 0.3};
  for (long long i = 0; i < nbf; ++i) {
    P[i * nbf + i] = occ[i % 7];
    for (long long j = 0; j < i; ++j) P[i * nbf + j] = P[j * nbf + i] = 0.01 / (double)(1 + i + j);
  }
  double exc = 0., nel = 0.;
  gauxc_integrator_eval_exc_vxc_rks(&st, integ, nbf, nbf, P, nbf, &exc, V, nbf);
  if (failed(&st, "eval_exc_vxc_rks")) return 1;
  gauxc_integrator_integrate_den(&st, integ, nbf, nbf, P, nbf, &nel);
  if (failed(&st, "integrate_den")) return 1;

  /* error behaviour of the reference: wrong dimension -> status 1, message, no abort */
  gauxc_integrator_eval_exc_vxc_rks(&st, integ, nbf + 1, nbf, P, nbf, &exc, V, nbf);
  if (st.code != 1 || !st.message || !strstr(st.message, "Must Be Square")) {
    fprintf(stderr, "expected the reference's dimension error\n");
    return 1;
  }
  gauxc_integrator_eval_exc_vxc_rks(&st, integ, nbf, nbf, P, nbf, &exc, V, nbf); /* clears the status */
  if (failed(&st, "eval_exc_vxc_rks (2)")) return 1;

  printf("EXC %.17g NEL %.17g NBF %lld VXC", exc, 2. * nel, nbf);
  for (long long i = 0; i < nbf * nbf; ++i) printf(" %.17g", V[i]);
  printf("\n");

  gauxc_integrator_delete(&st, &integ);
  gauxc_functional_delete(&st, &f);
  gauxc_molecular_weights_delete(&st, &mw);
  gauxc_molecular_weights_factory_delete(&st, &mwf);
  gauxc_load_balancer_delete(&st, &lb);
  gauxc_load_balancer_factory_delete(&st, &lbf);
  gauxc_runtime_environment_delete(&st, &rt);
  gauxc_molgrid_delete(&st, &mg);
  gauxc_basisset_delete(&st, &basis);
  gauxc_molecule_delete(&st, &mol);
  gauxc_status_delete(&st);
  free(P);
  free(V);
  return 0;
}
