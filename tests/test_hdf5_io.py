"""CPU: the self-contained HDF5 reader / writer behind gauxc_{molecule,basisset}_{read,write}_hdf5_record
(reference: src/external/hdf5_read.cxx:47-156, hdf5_write.cxx:22-86 over HighFive / libhdf5, absent here).
Round trips are checked with the product's C++ reader AND with the independent Python parser tools/h5mini.py;
when the reference tree is present (build container) its own fixture files are read and compared with the
committed golden .npz conversions."""
import os
import sys

import numpy as np
import pytest

from gauxc_b200 import capi, systems
import gauxc_b200 as gx

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tools"))


def test_round_trip_molecule_basis_datasets(tmp_path):
    from h5mini import H5File
    atoms = systems.geometry("benzene")
    shells = systems.make_basis_shells(atoms, "cc-pvdz")
    mol, basis = gx.Molecule(atoms), gx.BasisSet(shells)
    fn = str(tmp_path / "rt.hdf5")
    mol.write_hdf5(fn, "/MOLECULE")
    basis.write_hdf5(fn, "/BASIS")
    rng = np.random.default_rng(1)
    P = rng.standard_normal((7, 5))
    capi.hdf5_write_dataset(fn, "/DENSITY", P)
    capi.hdf5_write_dataset(fn, "/EXC", np.array([-1.25]))
    for k in range(12):  # more than one symbol-table node
        capi.hdf5_write_dataset(fn, f"/X{k:02d}", np.arange(k + 1.0))
    # product reader
    m2 = gx.Molecule.from_hdf5(fn)
    assert m2.atoms() == [tuple(a) for a in atoms]
    b2 = gx.BasisSet.from_hdf5(fn)
    assert b2.nshells() == basis.nshells() and b2.nbf() == basis.nbf()
    for s in range(basis.nshells()):
        a, b = basis.get_shell(s), b2.get_shell(s)
        assert a["l"] == b["l"] and a["pure"] == b["pure"] and a["nprim"] == b["nprim"]
        assert np.array_equal(a["alpha"], b["alpha"]) and np.array_equal(a["coeff"], b["coeff"])
        assert np.array_equal(a["origin"], b["origin"])
    assert np.array_equal(capi.hdf5_read_dataset(fn, "/DENSITY"), P)
    assert capi.hdf5_read_dataset(fn, "/EXC")[0] == -1.25
    assert np.array_equal(capi.hdf5_read_dataset(fn, "/X11"), np.arange(12.0))
    # independent parser
    f = H5File(fn)
    assert f.keys("/") == sorted(["MOLECULE", "BASIS", "DENSITY", "EXC"] + [f"X{k:02d}" for k in range(12)])
    assert np.array_equal(f.array("/DENSITY"), P)
    dims, esize, tclass, raw = f.raw("/MOLECULE")
    assert dims == (12,) and esize == 32 and tclass == 6
    dims, esize, tclass, raw = f.raw("/BASIS")
    assert dims == (basis.nshells(),) and esize == 552 and tclass == 6
    with pytest.raises(gx.GauXCError, match="Dataset Creation Failed"):
        capi.hdf5_write_dataset(fn, "/EXC", np.array([0.0]))
    with pytest.raises(gx.GauXCError, match="no such object"):
        capi.hdf5_read_dataset(fn, "/NOPE")
    with pytest.raises(gx.GauXCError, match="cannot open"):
        gx.Molecule.from_hdf5(str(tmp_path / "missing.hdf5"))


@pytest.mark.skipif(not os.path.isdir("/root/reference/tests/ref_data"), reason="reference fixtures not present")
@pytest.mark.parametrize("name", ["benzene_svwn5_cc-pvdz_ufg_ssf", "benzene_pbe0_cc-pvdz_ufg_ssf"])
def test_reads_the_reference_fixture_files(name):
    fn = f"/root/reference/tests/ref_data/{name}.hdf5"
    g = systems.golden(name)
    mol = gx.Molecule.from_hdf5(fn, "/MOLECULE")
    at = mol.atoms()
    assert [a[0] for a in at] == list(g["mol_Z"]) and np.array_equal(np.array([a[1:] for a in at]), g["mol_xyz"])
    basis = gx.BasisSet.from_hdf5(fn, "/BASIS")
    assert basis.nshells() == len(g["sh_l"]) and basis.nbf() == g["DENSITY"].shape[0]
    for s in range(basis.nshells()):
        sh = basis.get_shell(s)
        n = int(g["sh_nprim"][s])
        assert sh["l"] == g["sh_l"][s] and sh["nprim"] == n and bool(sh["pure"]) == bool(g["sh_pure"][s])
        assert np.array_equal(sh["alpha"][:n], g["sh_alpha"][s, :n]) and np.array_equal(sh["coeff"][:n], g["sh_coeff"][s, :n])
    assert np.array_equal(capi.hdf5_read_dataset(fn, "/DENSITY"), g["DENSITY"])
    assert np.array_equal(capi.hdf5_read_dataset(fn, "/VXC"), g["VXC"])
    assert capi.hdf5_read_dataset(fn, "/EXC")[0] == float(g["EXC"][0])
