"""The C ABI from plain C: tests/c_client/exc_vxc_client.c is compiled with gcc against
include/gauxc_b200.h and linked to libgauxc_b200.so -- the binding INTEGRATION.md section 1 describes
for programs written against the reference's <gauxc/c/...> headers."""
import os
import subprocess

import numpy as np
import pytest

from conftest import ROOT
from gauxc_b200 import capi

SRC = os.path.join(ROOT, "tests", "c_client", "exc_vxc_client.c")
SRC_CPP = os.path.join(ROOT, "tests", "c_client", "exc_vxc_client.cpp")
LIBDIR = os.path.join(ROOT, "gauxc_b200")


@pytest.fixture(scope="module")
def client(tmp_path_factory):
    exe = str(tmp_path_factory.mktemp("c_client") / "exc_vxc_client")
    cmd = ["gcc", "-std=c11", "-O1", "-Wall", "-Werror", "-I", os.path.join(ROOT, "include"), SRC, "-o", exe,
           "-L", LIBDIR, "-lgauxc_b200", "-Wl,-rpath," + LIBDIR]
    r = subprocess.run(cmd, capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    return exe


@pytest.fixture(scope="module")
def client_cpp(tmp_path_factory):
    exe = str(tmp_path_factory.mktemp("cpp_client") / "exc_vxc_client_cpp")
    cmd = ["g++", "-std=c++17", "-O1", "-Wall", "-Wextra", "-Werror", "-I", os.path.join(ROOT, "include"), SRC_CPP,
           "-o", exe, "-L", LIBDIR, "-lgauxc_b200", "-Wl,-rpath," + LIBDIR]
    r = subprocess.run(cmd, capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    return exe


def test_cpp_facade_client_compiles_links_and_fails_loudly_without_a_gpu(client_cpp):
    """include/gauxc_b200.hpp: the reference's C++ API surface for this path (factories, MatrixType
    facade, exceptions) header-only over the C ABI; same contract as the C client."""
    r = subprocess.run([client_cpp], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stderr
    assert r.stderr.startswith("nbf 7, ") and 170000 < int(r.stderr.split()[2]) <= 175230
    if capi.device_count() == 0:
        assert r.stdout.startswith("NO_DEVICE"), r.stdout
    else:
        assert r.stdout.startswith("EXC "), r.stdout


@pytest.mark.gpu
def test_cpp_facade_client_matches_the_c_client(client, client_cpp):
    a = subprocess.run([client], capture_output=True, text=True, timeout=300)
    b = subprocess.run([client_cpp], capture_output=True, text=True, timeout=300)
    assert a.returncode == 0 and b.returncode == 0, a.stderr + b.stderr
    ta, tb = a.stdout.split(), b.stdout.split()
    assert ta[0] == tb[0] == "EXC" and ta[5] == tb[5]
    va = np.array([float(x) for x in ta[7:]])
    vb = np.array([float(x) for x in tb[7:]])
    assert abs(float(ta[1]) - float(tb[1])) <= 1e-12 and np.abs(va - vb).max() <= 1e-12
    assert abs(float(ta[3]) - float(tb[3])) <= 1e-10


def test_c_client_compiles_links_and_fails_loudly_without_a_gpu(client):
    """C11 + -Werror against the header; host-side objects (grid, load balancer) work from C; without a
    CUDA device the Device entry points return status 1 "No CUDA device" -- there is no CPU fallback."""
    r = subprocess.run([client], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stderr
    # 3 atoms x 99 x 590 = 175 230 UltraFine points minus the far batches no shell reaches
    assert r.stderr.startswith("nbf 7, ") and 170000 < int(r.stderr.split()[2]) <= 175230
    if capi.device_count() == 0:
        assert r.stdout.startswith("NO_DEVICE"), r.stdout
    else:
        assert r.stdout.startswith("EXC "), r.stdout


@pytest.mark.gpu
def test_c_client_matches_the_python_binding(client):
    import gauxc_b200 as gx
    r = subprocess.run([client], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stderr
    tok = r.stdout.split()
    assert tok[0] == "EXC" and tok[2] == "NEL" and tok[4] == "NBF" and tok[6] == "VXC"
    exc_c, nel_c, nbf = float(tok[1]), float(tok[3]), int(tok[5])
    vxc_c = np.array([float(x) for x in tok[7:]]).reshape(nbf, nbf)

    atoms = [(8, 0., -0.07579, 0.), (1, 0.86681, 0.60144, 0.), (1, -0.86681, 0.60144, 0.)]
    a_o1, c_s = [130.70932, 23.808861, 6.4436083], [0.15432897, 0.53532814, 0.44463454]
    a_o2, c_2s = [5.0331513, 1.1695961, 0.3803890], [-0.09996723, 0.39951283, 0.70011547]
    c_2p = [0.15591627, 0.60768372, 0.39195739]
    a_h = [3.42525091, 0.62391373, 0.16885540]

    def sh(l, a, c, at):
        return dict(l=l, pure=True, exps=a, coefs=c, origin=list(at[1:]), tol=1e-10)

    shells = [sh(0, a_o1, c_s, atoms[0]), sh(0, a_o2, c_2s, atoms[0]), sh(1, a_o2, c_2p, atoms[0]),
              sh(0, a_h, c_s, atoms[1]), sh(0, a_h, c_s, atoms[2])]
    mol = gx.Molecule(atoms)
    basis = gx.BasisSet(shells, normalize=True)
    assert basis.nbf() == nbf
    mg = gx.MolGrid(mol, "Unpruned", 512, "MuraKnowles", "UltraFineGrid")
    rt = gx.RuntimeEnvironment(device=True)
    lb = gx.LoadBalancerFactory("Host", "Replicated").get_instance(rt, mol, mg, basis)
    gx.MolecularWeightsFactory("Device", "Default", "SSF").get_instance().modify_weights(lb)
    integ = gx.XCIntegratorFactory("Device").get_instance(gx.Functional("PBE"), lb)
    occ = [1.0, 0.9, 0.7, 0.7, 0.7, 0.3, 0.3]
    P = np.zeros((nbf, nbf), order="F")
    for i in range(nbf):
        P[i, i] = occ[i % 7]
        for j in range(i):
            P[i, j] = P[j, i] = 0.01 / (1 + i + j)
    exc, vxc = integ.eval_exc_vxc(P)
    assert abs(exc - exc_c) <= 1e-12 * max(1.0, abs(exc))
    assert np.abs(vxc - vxc_c).max() <= 1e-12
    assert abs(nel_c - integ.stats()["n_el"]) <= 1e-10


SRC_ADAPTOR = os.path.join(ROOT, "tests", "c_client", "b200_integrator_adaptor.cpp")


@pytest.fixture(scope="module")
def adaptor(tmp_path_factory):
    exe = str(tmp_path_factory.mktemp("adaptor") / "b200_integrator_adaptor")
    cmd = ["g++", "-std=c++17", "-O1", "-Wall", "-Wextra", "-Werror", "-I", os.path.join(ROOT, "include"), SRC_ADAPTOR,
           "-o", exe, "-L", LIBDIR, "-lgauxc_b200", "-Wl,-rpath," + LIBDIR]
    r = subprocess.run(cmd, capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    return exe


def test_reference_side_adaptor_compiles_and_fails_loudly_without_a_gpu(adaptor):
    """the ReplicatedXCDeviceIntegrator<double>::eval_exc_vxc_ hook implemented over the
    C ABI (the binding a GauXC maintainer would add), compiled -Werror; no CPU fallback behind it."""
    if capi.device_count() > 0:
        pytest.skip("GPU present: covered by the gpu test")
    r = subprocess.run([adaptor], capture_output=True, text=True, timeout=300)
    assert r.returncode == 1 and "No CUDA device" in r.stderr, r.stdout + r.stderr


@pytest.mark.gpu
def test_reference_side_adaptor_matches_the_direct_call(adaptor):
    """The adaptor hands the 'reference' task list (points, SSF weights, shell lists) to the device path through
    gauxc_b200_load_balancer_set_tasks and must reproduce the direct integrator to 1e-12."""
    r = subprocess.run([adaptor], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stdout + r.stderr
    assert r.stdout.startswith("EXC direct"), r.stdout
