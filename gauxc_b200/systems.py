"""Benchmark / test systems of BASELINE.json: geometries, basis sets, synthetic densities.

Geometries are the reference's own (tests/standards.cxx:18-1405, converted once by
tools/make_golden.py); cc-pVDZ is the reference's tests/basis/old/cc-pvdz.g94 (EMSL data);
def2-SVP H/C/N/O is written from the published tables (C/H exponents cross-checked against the
reference's benzene def2-SVP fixture).  Densities for systems without an SCF solution are the
deterministic synthetic matrices described in SURVEY.md 8(d).
"""
import json
import os
import numpy as np

_DATA = os.path.join(os.path.dirname(os.path.abspath(__file__)), "data")
_SYM = {1: "H", 6: "C", 7: "N", 8: "O", 16: "S"}


def geometry(name):
    g = json.load(open(os.path.join(_DATA, "geometries.json")))
    return [tuple(a) for a in g[name]]


def water_cluster(n=833, spacing=5.8):
    """n copies of make_water() on a cubic lattice (deterministic order, x fastest)."""
    w = geometry("water")
    side = int(np.ceil(n ** (1. / 3.)))
    atoms = []
    k = 0
    for iz in range(side):
        for iy in range(side):
            for ix in range(side):
                if k >= n:
                    break
                for (Z, x, y, z) in w:
                    atoms.append((Z, x + ix * spacing, y + iy * spacing, z + iz * spacing))
                k += 1
    return atoms


def basis_library(name):
    return json.load(open(os.path.join(_DATA, name.lower() + ".json")))


def make_basis_shells(atoms, name, spherical=True, tol=1e-10):
    """Shell dicts for capi.BasisSet (to be normalised by the library), atom-major order like
    the reference's parse_basis (tests/basis/parse_basis.cxx:149-240): spherical applies to
    l > 1 only."""
    lib = basis_library(name)
    shells = []
    for (Z, x, y, z) in atoms:
        for sh in lib[_SYM[Z]]:
            shells.append(dict(l=sh["l"], pure=bool(spherical and sh["l"] > 1), exps=sh["exps"],
                               coefs=sh["coefs"], origin=(x, y, z), tol=tol))
    return shells


def shell_size(sh):
    l = sh["l"]
    return 2 * l + 1 if sh["pure"] else (l + 1) * (l + 2) // 2


def synthetic_density(atoms, shells, seed=20261017, amp=1e-3, rcut=6.0, floor=0.02):
    """Deterministic symmetric P (alpha density): per-atom diagonal occupations carrying Z/2
    electrons, a small floor on every function, plus a seeded symmetric perturbation between
    functions whose centres are closer than rcut."""
    from scipy.spatial import cKDTree
    sizes = np.array([shell_size(s) for s in shells])
    first = np.concatenate([[0], np.cumsum(sizes)])
    nbf = int(first[-1])
    xyz = np.array([a[1:] for a in atoms])
    # shells per atom (shell origins coincide with atoms)
    tree = cKDTree(xyz)
    sh_atom = tree.query(np.array([s["origin"] for s in shells]))[1]
    diag = np.full(nbf, floor)
    for ia, (Z, *_r) in enumerate(atoms):
        idx = np.where(sh_atom == ia)[0]
        remaining = Z / 2.0
        for l, cap in ((0, 1.0), (1, 3.0), (2, 5.0)):
            ls = [i for i in idx if shells[i]["l"] == l]
            use = ls[:-1] if len(ls) > 1 else ls
            for i in use:
                occ = min(cap, remaining)
                if occ <= 0:
                    break
                diag[first[i]:first[i + 1]] += occ / sizes[i]
                remaining -= occ
    P = np.diag(diag)
    # perturbation between near atoms (including the atom with itself)
    rng = np.random.default_rng(seed)
    ao_atom = np.repeat(sh_atom, sizes)
    atom_aos = [np.where(ao_atom == ia)[0] for ia in range(len(atoms))]
    pairs = tree.query_pairs(rcut, output_type="ndarray")
    for ia in range(len(atoms)):
        a = atom_aos[ia]
        blk = rng.standard_normal((len(a), len(a))) * amp
        blk = np.triu(blk, 1)
        P[np.ix_(a, a)] += blk + blk.T
    for (ia, ja) in pairs:
        a, b = atom_aos[ia], atom_aos[ja]
        blk = rng.standard_normal((len(a), len(b))) * amp
        P[np.ix_(a, b)] += blk
        P[np.ix_(b, a)] += blk.T
    return np.asfortranarray(P)


# name -> (geometry, basis, functional, grid, P source)
CONFIGS = {
    "water": dict(geom="water", basis="cc-pvdz", func="SVWN5", grid="UltraFineGrid"),
    "benzene": dict(geom="benzene", basis="cc-pvdz", func="PBE", grid="UltraFineGrid"),
    "taxol": dict(geom="taxol", basis="def2-svp", func="PBE", grid="SuperFineGrid"),
    "ubiquitin": dict(geom="ubiquitin", basis="cc-pvdz", func="SVWN5", grid="FineGrid"),
    "water833": dict(geom="water_cluster:833", basis="cc-pvdz", func="PBE", grid="UltraFineGrid"),
}


def config_atoms(cfg):
    g = CONFIGS[cfg]["geom"]
    if g.startswith("water_cluster:"):
        return water_cluster(int(g.split(":")[1]))
    return geometry(g)


def golden(name):
    """One of the golden fixtures converted from the reference's tests/ref_data."""
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    return np.load(os.path.join(root, "tests", "golden", name + ".npz"))


def golden_system(name):
    """(atoms, shells[already normalised], P, VXC, EXC) of a golden EXC/VXC fixture."""
    d = golden(name)
    atoms = [(int(Z), *xyz) for Z, xyz in zip(d["mol_Z"], d["mol_xyz"])]
    shells = []
    for i in range(len(d["sh_l"])):
        n = int(d["sh_nprim"][i])
        shells.append(dict(l=int(d["sh_l"][i]), pure=bool(d["sh_pure"][i]), exps=list(d["sh_alpha"][i, :n]),
                           coefs=list(d["sh_coeff"][i, :n]), origin=tuple(d["sh_O"][i])))
    return atoms, shells, d["DENSITY"], d["VXC"], float(d["EXC"][0])
