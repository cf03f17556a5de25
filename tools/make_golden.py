#!/usr/bin/env python
"""Convert the reference's golden HDF5 fixtures into small flat .npz files.

Run ONCE in the build container (needs /root/reference); the outputs under
tests/golden/ and gauxc_b200/data/ are committed so that nothing at test /
bench time reads /root/reference (it does not exist on the GPU box).

Sources (all under /root/reference/tests/):
  ref_data/benzene_{svwn5,pbe0}_cc-pvdz_ufg_ssf[_robust_prune|_treutler_prune].hdf5
      consumed by tests/xc_integrator.cxx:405-426       -> benzene_*.npz
  ref_data/cytosine_svwn5_cc-pvdz_ufg_ssf_robust_uks.hdf5
      consumed by tests/xc_integrator.cxx:455-459 (UKS LDA)  -> cytosine_svwn5_..._uks.npz
  ref_data/benzene_*.hdf5 /EXC_GRAD_HELLFEY, /EXC_GRAD_FULL
      consumed by tests/xc_integrator.cxx:276-297       -> benzene_exc_grad.npz
  ref_data/water_cc-pVDZ_collocation.hdf5
      consumed by tests/collocation.cxx:45-91           -> water_collocation.npz
  ref_data/benzene_weights_ssf.hdf5
      consumed by tests/weights.cxx:58-77               -> benzene_weights_ssf.npz
  ref_data/benzene_cc-pvdz_ufg_tasks_1mpi_rank0_pv1.hdf5
      consumed by tests/load_balancer_test.cxx:40-71    -> benzene_lb_tasks.npz
  basis/old/cc-pvdz.g94 (EMSL data; tests/standards.cxx:1421-1425)
                                                        -> gauxc_b200/data/cc-pvdz.json
  standards.cxx:18-1405 geometries (bohr)               -> gauxc_b200/data/geometries.json
"""
import json
import os
import re
import sys
import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, HERE)
from h5mini import H5File  # noqa: E402

REF = "/root/reference/tests"
OUT = os.path.join(ROOT, "tests", "golden")
DATA = os.path.join(ROOT, "gauxc_b200", "data")

MOL_DT = np.dtype([("Z", "<i4"), ("pad", "<i4"), ("x", "<f8"), ("y", "<f8"), ("z", "<f8")])
SH_DT = np.dtype([("nprim", "<i4"), ("l", "<i4"), ("pure", "<i4"), ("pad", "<i4"),
                  ("alpha", "<f8", 32), ("coeff", "<f8", 32), ("O", "<f8", 3)])
assert MOL_DT.itemsize == 32 and SH_DT.itemsize == 552


def mol_basis(f):
    d, es, tc, raw = f.raw("/MOLECULE")
    m = np.frombuffer(raw, dtype=MOL_DT)
    out = dict(mol_Z=m["Z"].astype(np.int64),
               mol_xyz=np.stack([m["x"], m["y"], m["z"]], 1))
    if "BASIS" in f.keys("/"):
        d, es, tc, raw = f.raw("/BASIS")
        b = np.frombuffer(raw, dtype=SH_DT)
        out.update(sh_nprim=b["nprim"].copy(), sh_l=b["l"].copy(), sh_pure=b["pure"].copy(),
                   sh_alpha=b["alpha"].copy(), sh_coeff=b["coeff"].copy(), sh_O=b["O"].copy())
        # zero the unused tails (they hold uninitialised bytes in the fixtures)
        for i, n in enumerate(out["sh_nprim"]):
            out["sh_alpha"][i, n:] = 0
            out["sh_coeff"][i, n:] = 0
    return out


def conv_xc(name):
    f = H5File(f"{REF}/ref_data/{name}.hdf5")
    o = mol_basis(f)
    o["DENSITY"] = f.array("/DENSITY")
    o["VXC"] = f.array("/VXC")
    o["EXC"] = f.array("/EXC")
    np.savez_compressed(f"{OUT}/{name}.npz", **o)
    print(name, "EXC", o["EXC"], "nbf", o["DENSITY"].shape)


def conv_xc_uks(name):
    """UKS fixtures (tests/xc_integrator.cxx:448-472): scalar and z densities / potentials."""
    f = H5File(f"{REF}/ref_data/{name}.hdf5")
    o = mol_basis(f)
    for k in ("DENSITY_SCALAR", "DENSITY_Z", "VXC_SCALAR", "VXC_Z", "EXC"):
        o[k] = f.array("/" + k)
    np.savez_compressed(f"{OUT}/{name}.npz", **o)
    print(name, "EXC", o["EXC"], "nbf", o["DENSITY_SCALAR"].shape)


def conv_grad():
    """EXC gradients (tests/xc_integrator.cxx:117-150, 276-297): /EXC_GRAD_HELLFEY (include_weight_derivatives
    = false) and /EXC_GRAD_FULL (true), natoms x 3 row-major, for the RKS benzene fixtures."""
    o = {}
    for name in ("benzene_svwn5_cc-pvdz_ufg_ssf", "benzene_pbe0_cc-pvdz_ufg_ssf",
                 "benzene_svwn5_cc-pvdz_ufg_ssf_robust_prune", "benzene_svwn5_cc-pvdz_ufg_ssf_treutler_prune"):
        f = H5File(f"{REF}/ref_data/{name}.hdf5")
        for k in ("EXC_GRAD_HELLFEY", "EXC_GRAD_FULL"):
            o[f"{name}:{k}"] = np.asarray(f.array("/" + k), dtype=np.float64).reshape(-1, 3)
        print(name, "grad", o[f"{name}:EXC_GRAD_FULL"].shape, np.abs(o[f"{name}:EXC_GRAD_FULL"]).max())
    np.savez_compressed(f"{OUT}/benzene_exc_grad.npz", **o)
    o = {}
    for name in ("cytosine_svwn5_cc-pvdz_ufg_ssf_robust_uks", "cytosine_blyp_cc-pvdz_ufg_ssf_robust_uks"):
        f = H5File(f"{REF}/ref_data/{name}.hdf5")
        for k in ("EXC_GRAD_HELLFEY", "EXC_GRAD_FULL"):
            o[f"{name}:{k}"] = np.asarray(f.array("/" + k), dtype=np.float64).reshape(-1, 3)
        print(name, "grad", o[f"{name}:EXC_GRAD_FULL"].shape, np.abs(o[f"{name}:EXC_GRAD_FULL"]).max())
    np.savez_compressed(f"{OUT}/cytosine_uks_exc_grad.npz", **o)


def conv_basis_only(name, out):
    f = H5File(f"{REF}/ref_data/{name}.hdf5")
    o = mol_basis(f)
    np.savez_compressed(f"{OUT}/{out}.npz", **o)


def conv_collocation():
    f = H5File(f"{REF}/ref_data/water_cc-pVDZ_collocation.hdf5")
    o = {}
    n = 0
    while f"entry_{n}" in f.keys("/"):
        for k in ("mask", "pts", "eval", "deval_x", "deval_y", "deval_z"):
            o[f"e{n}_{k}"] = f.array(f"/entry_{n}/{k}")
        n += 1
    o["nentries"] = np.array([n])
    np.savez_compressed(f"{OUT}/water_collocation.npz", **o)
    print("collocation entries", n)


def conv_weights():
    f = H5File(f"{REF}/ref_data/benzene_weights_ssf.hdf5")
    o = mol_basis(f)
    nt = int(f.array("/tasks_unm/ntasks")[0])
    o["ntasks"] = np.array([nt])
    for i in range(nt):
        g = f"/tasks_unm/task_{i}"
        o[f"t{i}_points"] = f.array(g + "/points")
        o[f"t{i}_weights"] = f.array(g + "/weights")
        o[f"t{i}_iParent"] = f.array(g + "/iParent")
        o[f"t{i}_dist_nearest"] = f.array(g + "/dist_nearest")
        o[f"t{i}_weights_mod"] = f.array(f"/tasks_mod/task_{i}/weights")
    np.savez_compressed(f"{OUT}/benzene_weights_ssf.npz", **o)
    print("weights tasks", nt, [len(o[f"t{i}_weights"]) for i in range(nt)])


def conv_lb():
    f = H5File(f"{REF}/ref_data/benzene_cc-pvdz_ufg_tasks_1mpi_rank0_pv1.hdf5")
    nt = int(f.array("/tasks/ntasks")[0])
    ip, npts, nbe, dn, sl, slo = [], [], [], [], [], [0]
    for i in range(nt):
        g = f"/tasks/task_{i}"
        ip.append(int(f.array(g + "/iParent")[0]))
        npts.append(int(f.array(g + "/npts")[0]))
        nbe.append(int(f.array(g + "/bfn_screening_nbe")[0]))
        dn.append(float(f.array(g + "/dist_nearest")[0]))
        s = f.array(g + "/shell_list")
        sl.append(s)
        slo.append(slo[-1] + len(s))
    np.savez_compressed(f"{OUT}/benzene_lb_tasks.npz", iParent=np.array(ip), npts=np.array(npts),
                        nbe=np.array(nbe), dist_nearest=np.array(dn),
                        shell_list=np.concatenate(sl).astype(np.int32), shell_off=np.array(slo))
    print("lb tasks", nt, "npts", sum(npts))


def parse_g94(path, want):
    """EMSL Gaussian94 format -> {symbol: [ {l, exps, coefs} ]} (SP shells split)."""
    AM = dict(S=0, P=1, D=2, F=3, G=4, H=5, I=6)
    txt = open(path).read().replace("D+", "E+").replace("D-", "E-")
    blocks = txt.split("****")
    out = {}
    for blk in blocks:
        lines = [l for l in blk.strip().splitlines() if l.strip() and not l.startswith("!")]
        if not lines:
            continue
        sym = lines[0].split()[0]
        if sym not in want:
            continue
        shells, i = [], 1
        while i < len(lines):
            t, n = lines[i].split()[0], int(lines[i].split()[1])
            rows = [[float(x) for x in lines[i + 1 + k].split()] for k in range(n)]
            i += 1 + n
            if t == "SP":
                shells.append(dict(l=0, exps=[r[0] for r in rows], coefs=[r[1] for r in rows]))
                shells.append(dict(l=1, exps=[r[0] for r in rows], coefs=[r[2] for r in rows]))
            else:
                shells.append(dict(l=AM[t], exps=[r[0] for r in rows], coefs=[r[1] for r in rows]))
        out[sym] = shells
    return out


def conv_basis_lib():
    os.makedirs(DATA, exist_ok=True)
    want = {"H", "C", "N", "O", "S"}
    b = parse_g94(f"{REF}/basis/old/cc-pvdz.g94", want)
    json.dump(b, open(f"{DATA}/cc-pvdz.json", "w"), indent=0)
    print("cc-pvdz", {k: len(v) for k, v in b.items()})


def conv_geometries():
    src = open(f"{REF}/standards.cxx").read()
    out = {}
    for name in ("water", "benzene", "taxol", "ubiquitin"):
        m = re.search(r"Molecule make_%s\(\) \{(.*?)return mol;" % name, src, re.S)
        atoms = []
        for line in m.group(1).splitlines():
            line = line.strip()
            if line.startswith("//"):
                continue
            mm = re.match(r"mol\.emplace_back\(AtomicNumber\((\d+)\),\s*([^,]+),\s*([^,]+),\s*([^)]+)\);", line)
            if mm:
                atoms.append([int(mm.group(1)), float(mm.group(2)), float(mm.group(3)), float(mm.group(4))])
        out[name] = atoms
        print(name, len(atoms))
    json.dump(out, open(f"{DATA}/geometries.json", "w"))


if __name__ == "__main__":
    os.makedirs(OUT, exist_ok=True)
    for n in ("benzene_svwn5_cc-pvdz_ufg_ssf", "benzene_pbe0_cc-pvdz_ufg_ssf",
              "benzene_svwn5_cc-pvdz_ufg_ssf_robust_prune", "benzene_svwn5_cc-pvdz_ufg_ssf_treutler_prune"):
        conv_xc(n)
    conv_xc_uks("cytosine_svwn5_cc-pvdz_ufg_ssf_robust_uks")
    conv_xc_uks("cytosine_blyp_cc-pvdz_ufg_ssf_robust_uks")
    conv_grad()
    conv_basis_only("benzene_m062x_def2-svp_ufg_ssf", "benzene_def2-svp_basis")
    conv_collocation()
    conv_weights()
    conv_lb()
    conv_basis_lib()
    conv_geometries()
