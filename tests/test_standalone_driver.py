"""The INI front-end (tools/standalone_driver.cxx; reference tests/standalone_driver.cxx:31-851 + ini_input.cxx):
the reference's input-deck keys over an HDF5 file with the reference's record layout.  CPU: the deck is parsed, the
records are read, and the Device path refuses to run without a GPU; GPU: the benzene SVWN5 / PBE0 fixtures
(rewritten from the committed golden .npz with the product's HDF5 writer) reproduce their EXC / VXC."""
import os
import subprocess

import numpy as np
import pytest

from gauxc_b200 import capi, systems
import gauxc_b200 as gx

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
EXE = os.path.join(ROOT, "gauxc_b200", "standalone_driver")


def write_fixture(name, path):
    atoms, shells, P, VXC, EXC = systems.golden_system(name)
    gx.Molecule(atoms).write_hdf5(path, "/MOLECULE")
    gx.BasisSet(shells, normalize=False).write_hdf5(path, "/BASIS")
    capi.hdf5_write_dataset(path, "/DENSITY", P)
    capi.hdf5_write_dataset(path, "/VXC", VXC)
    capi.hdf5_write_dataset(path, "/EXC", np.array([EXC]))
    gold = np.load(os.path.join(ROOT, "tests", "golden", "benzene_exc_grad.npz"))
    if f"{name}:EXC_GRAD_FULL" in gold.files:
        capi.hdf5_write_dataset(path, "/EXC_GRAD_FULL", gold[f"{name}:EXC_GRAD_FULL"])


def deck(tmp_path, ref, func, extra=""):
    inp = tmp_path / "input.inp"
    inp.write_text(f"""# same keys as the reference's tests/ref_data/ut_input.inp
[GAUXC]
ref_file = {ref}
grid = UltraFine
pruning_scheme = Unpruned
batch_size = 512
basis_tol = 2.22e-16
func = {func}
integrate_vxc = TRUE
integrate_den = TRUE
integrate_exx = FALSE
outfile = {tmp_path / 'out.hdf5'}
{extra}""")
    return str(inp)


def test_driver_parses_deck_and_fails_loudly_without_gpu(tmp_path):
    ref = str(tmp_path / "ref.hdf5")
    write_fixture("benzene_svwn5_cc-pvdz_ufg_ssf", ref)
    r = subprocess.run([EXE, deck(tmp_path, ref, "svwn5")], capture_output=True, text=True)
    assert "REF_FILE" in r.stdout and "FUNCTIONAL        = SVWN5" in r.stdout
    assert "12 atoms" in r.stdout and "114 functions" in r.stdout
    if capi.device_count() == 0:
        assert r.returncode == 1 and "No CUDA device" in r.stderr
    r = subprocess.run([EXE, deck(tmp_path, ref, "svwn5", "int_exec_space = Host")], capture_output=True, text=True)
    assert r.returncode == 1 and "Host" in r.stderr
    r = subprocess.run([EXE, deck(tmp_path, str(tmp_path / "nope.hdf5"), "svwn5")], capture_output=True, text=True)
    assert r.returncode == 1 and "cannot open" in r.stderr


@pytest.mark.gpu
@pytest.mark.parametrize("name,func", [("benzene_svwn5_cc-pvdz_ufg_ssf", "svwn5"), ("benzene_pbe0_cc-pvdz_ufg_ssf", "pbe0")])
def test_driver_reproduces_the_reference_fixture(tmp_path, name, func):
    ref = str(tmp_path / "ref.hdf5")
    write_fixture(name, ref)
    r = subprocess.run([EXE, deck(tmp_path, ref, func)], capture_output=True, text=True)
    assert r.returncode == 0, r.stdout + r.stderr
    exc = float(capi.hdf5_read_dataset(str(tmp_path / "out.hdf5"), "/EXC")[0])
    vxc = capi.hdf5_read_dataset(str(tmp_path / "out.hdf5"), "/VXC")
    g = systems.golden(name)
    assert abs(exc - float(g["EXC"][0])) < 1e-10
    assert np.abs(vxc - g["VXC"]).max() < 1e-10
    nel = float(capi.hdf5_read_dataset(str(tmp_path / "out.hdf5"), "/N_EL")[0])
    assert abs(nel - 21.0) < 1e-3  # integrate_den of the alpha density: 42 electrons / 2
    m = gx.Molecule.from_hdf5(str(tmp_path / "out.hdf5"))
    assert m.natoms() == 12


@pytest.mark.gpu
def test_driver_exc_gradient_matches_the_reference_fixture(tmp_path):
    """integrate_exc_grad = TRUE (reference tests/standalone_driver.cxx:495-514, 715-737): the default settings
    include the weight derivatives, compared with the fixture's /EXC_GRAD_FULL."""
    name = "benzene_pbe0_cc-pvdz_ufg_ssf"
    ref = str(tmp_path / "ref.hdf5")
    write_fixture(name, ref)
    r = subprocess.run([EXE, deck(tmp_path, ref, "pbe0", "integrate_exc_grad = TRUE")], capture_output=True, text=True)
    assert r.returncode == 0, r.stdout + r.stderr
    assert "| EXC_GRAD (diff) |" in r.stdout
    g = capi.hdf5_read_dataset(str(tmp_path / "out.hdf5"), "/EXC_GRAD")
    gold = np.load(os.path.join(ROOT, "tests", "golden", "benzene_exc_grad.npz"))[f"{name}:EXC_GRAD_FULL"]
    assert np.abs(np.asarray(g).reshape(-1, 3) - gold).max() < 1e-10
