#!/usr/bin/env python
"""Developer check on a B200: peaks, kernel parity against the oracle/goldens, timings."""
import os
import sys
import time
import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import gauxc_b200 as gx  # noqa: E402
from gauxc_b200 import capi, systems  # noqa: E402
import pyoracle as orc  # noqa: E402
from __graft_entry__ import _match_raw_weights  # noqa: E402


def main():
    print("devices", capi.device_count(), "blas", orc.init_blas(), "threads", orc.num_threads())
    if "--noprobe" not in sys.argv:
        for w in ("dmma", "dfma", "copy"):
            print("peak", w, capi.probe_peak(w))

    # collocation vs golden
    atoms = systems.geometry("water")
    shells = systems.make_basis_shells(atoms, "cc-pvdz", spherical=True)
    basis = gx.BasisSet(shells, normalize=True)
    col = systems.golden("water_collocation")
    err = 0
    for e in range(int(col["nentries"][0])):
        mask, pts = col[f"e{e}_mask"], col[f"e{e}_pts"]
        ev, dx, dy, dz = capi.eval_collocation(basis, mask, pts, gradient=True)
        for a, k in ((ev, "eval"), (dx, "deval_x"), (dy, "deval_y"), (dz, "deval_z")):
            err = max(err, np.abs(a - col[f"e{e}_{k}"].reshape(a.shape)).max())
    print("device collocation vs golden max err", err)

    for name, fn in (("benzene_svwn5_cc-pvdz_ufg_ssf", "SVWN5"), ("benzene_pbe0_cc-pvdz_ufg_ssf", "PBE0"),
                     ("benzene_pbe0_cc-pvdz_ufg_ssf", "PBE")):
        atoms, shells, P, VXC, EXC = systems.golden_system(name)
        for s in shells:
            s["tol"] = np.finfo(float).eps
        mol = gx.Molecule(atoms)
        basis = gx.BasisSet(shells, normalize=False)
        mg = gx.MolGrid(mol, "Unpruned", 512, "MuraKnowles", "UltraFineGrid")
        rt = gx.RuntimeEnvironment(device=True)
        lb = gx.LoadBalancerFactory("Host", "Replicated").get_instance(rt, mol, mg, basis)
        raw = lb.export_tasks()
        mw = gx.MolecularWeightsFactory("Device", "Default", "SSF").get_instance()
        t = time.time()
        mw.modify_weights(lb)
        print(name, fn, "ssf wall", time.time() - t, "kernel ms", mw.last_ms())
        tasks = lb.export_tasks()
        coords = np.array([a[1:] for a in atoms])
        w_or = orc.ssf_weights(coords, tasks["npts"], tasks["iParent"], tasks["dist_nearest"], tasks["points"],
                               _match_raw_weights(raw, tasks))
        print("  ssf max abs diff vs oracle", np.abs(w_or - tasks["weights"]).max(), "wmax", w_or.max())
        integ = gx.XCIntegratorFactory("Device").get_instance(gx.Functional(fn), lb)
        exc, vxc = integ.eval_exc_vxc(P)
        ref = orc.exc_vxc(basis.flat(), basis.nbf(), P, tasks, fn)
        st = integ.stats()
        print("  EXC", exc, "oracle", ref["exc"], "d", exc - ref["exc"], "golden d", exc - EXC if fn != "PBE" else None)
        print("  VXC max diff oracle", np.abs(vxc - ref["vxc"]).max(), "golden",
              np.abs(vxc - VXC).max() if fn != "PBE" else None, "asym", np.abs(vxc - vxc.T).max())
        print("  NEL", st["n_el"], "oracle", ref["nel"], "d", st["n_el"] - ref["nel"])
        integ.set_profile(True)
        exc2, vxc2 = integ.eval_exc_vxc(P)
        st = integ.stats()
        print("  rerun dEXC", exc2 - exc, "dVXC", np.abs(vxc2 - vxc).max())
        print("  profile:", {k: round(v, 4) for k, v in st.items()})
        integ.set_profile(False)
        ts = []
        for _ in range(5):
            integ.eval_exc_vxc(P)
            ts.append(integ.stats()["local_work_ms"])
        st = integ.stats()
        ms = min(ts)
        print("  local work ms", ts, "total", st["total_ms"], "pts/s", st["npts"] / ms * 1e3, "TF/s",
              st["f_dense"] / ms / 1e9)


if __name__ == "__main__":
    main()
