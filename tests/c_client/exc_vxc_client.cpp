// The same calculation as exc_vxc_client.c through the C++ facade include/gauxc_b200.hpp, written the
// way the reference's tests/standalone_driver.cxx:150-480 drives GauXC (factories selected by
// ExecutionSpace + strings, std::tie(EXC, VXC) = integrator.eval_exc_vxc(P)).  Same output format.
#include <gauxc_b200.hpp>

#include <cstdio>
#include <cstring>
#include <vector>

using namespace GauXC;

// minimal column-major matrix with the interface XCIntegrator<MatrixType> needs (the reference's
// drivers use Eigen::MatrixXd, include/gauxc/xc_integrator/replicated/impl.hpp:108-120)
struct matrix_type {
  using value_type = double;
  matrix_type() = default;
  matrix_type(long r, long c) : r_(r), c_(c), v_((size_t)(r * c), 0.) {}
  long rows() const { return r_; }
  long cols() const { return c_; }
  double* data() { return v_.data(); }
  const double* data() const { return v_.data(); }
  double& operator()(long i, long j) { return v_[(size_t)(i + j * r_)]; }
  long r_ = 0, c_ = 0;
  std::vector<double> v_;
};

static Shell<double> make_shell(int l, const double (&a)[3], const double (&c)[3], const Atom& at) {
  Shell<double>::prim_array alpha{}, coeff{};
  for (int k = 0; k < 3; ++k) { alpha[k] = a[k]; coeff[k] = c[k]; }
  Shell<double> s(PrimSize(3), AngularMomentum(l), SphericalType(1), alpha, coeff, {at.x, at.y, at.z});
  s.set_shell_tolerance(1e-10);
  return s;
}

int main() {
  try {
    Molecule mol;
    mol.emplace_back(AtomicNumber(8), 0., -0.07579, 0.);
    mol.emplace_back(AtomicNumber(1), 0.86681, 0.60144, 0.);
    mol.emplace_back(AtomicNumber(1), -0.86681, 0.60144, 0.);
    const double a_o1[3] = {130.70932, 23.808861, 6.4436083}, c_s[3] = {0.15432897, 0.53532814, 0.44463454};
    const double a_o2[3] = {5.0331513, 1.1695961, 0.3803890}, c_2s[3] = {-0.09996723, 0.39951283, 0.70011547};
    const double c_2p[3] = {0.15591627, 0.60768372, 0.39195739};
    const double a_h[3] = {3.42525091, 0.62391373, 0.16885540};
    BasisSet<double> basis;
    basis.push_back(make_shell(0, a_o1, c_s, mol[0]));
    basis.push_back(make_shell(0, a_o2, c_2s, mol[0]));
    basis.push_back(make_shell(1, a_o2, c_2p, mol[0]));
    basis.push_back(make_shell(0, a_h, c_s, mol[1]));
    basis.push_back(make_shell(0, a_h, c_s, mol[2]));

    auto mg = MolGridFactory::create_default_molgrid(mol, PruningScheme::Unpruned, BatchSize(512),
                                                     RadialQuad::MuraKnowles, AtomicGridSizeDefault::UltraFineGrid);
    auto rt = DeviceRuntimeEnvironment(0.5);
    LoadBalancerFactory lb_factory(ExecutionSpace::Host, "Default");
    auto lb = lb_factory.get_shared_instance(rt, mol, mg, basis);
    const long nbf = basis.nbf();
    std::fprintf(stderr, "nbf %ld, %zu grid points, %zu tasks\n", nbf, lb->total_npts(), lb->ntasks());

    MolecularWeightsFactory mw_factory(ExecutionSpace::Device, "Default", MolecularWeightsSettings{});
    auto mw = mw_factory.get_instance();
    mw.modify_weights(*lb);

    functional_type func("PBE");
    XCIntegratorFactory<matrix_type> integrator_factory(ExecutionSpace::Device, "Replicated", "Default", "Default",
                                                        "Default");
    auto integrator = integrator_factory.get_instance(func, lb);

    matrix_type P(nbf, nbf), VXC;
    const double occ[7] = {1.0, 0.9, 0.7, 0.7, 0.7, 0.3, 0.3};
    for (long i = 0; i < nbf; ++i) {
      P(i, i) = occ[i % 7];
      for (long j = 0; j < i; ++j) P(i, j) = P(j, i) = 0.01 / (double)(1 + i + j);
    }
    double EXC = 0.;
    std::tie(EXC, VXC) = integrator.eval_exc_vxc(P);
    const double N_EL = integrator.integrate_den(P);

    // the reference's error behaviour surfaces as an exception with the reference's message
    bool threw = false;
    try {
      matrix_type bad(nbf + 1, nbf);
      integrator.eval_exc_vxc(bad);
    } catch (const generic_gauxc_exception& e) {
      threw = std::strstr(e.what(), "Must Be Square") != nullptr;
    }
    if (!threw) {
      std::fprintf(stderr, "expected the reference's dimension error\n");
      return 1;
    }

    std::printf("EXC %.17g NEL %.17g NBF %ld VXC", EXC, 2. * N_EL, nbf);
    for (long i = 0; i < nbf * nbf; ++i) std::printf(" %.17g", VXC.data()[i]);
    std::printf("\n");
  } catch (const generic_gauxc_exception& e) {
    if (std::strstr(e.what(), "No CUDA device")) {
      std::printf("NO_DEVICE %s\n", e.what());
      return 0;
    }
    std::fprintf(stderr, "GauXC exception: %s\n", e.what());
    return 1;
  }
  return 0;
}
