#!/usr/bin/env python
"""LoadBalancer task creation, Host vs Device execution space.  usage: python tools/lb_time.py taxol water833"""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    import gauxc_b200 as gx
    from gauxc_b200 import systems
    os.environ["GAUXC_B200_LB_TIMING"] = "1"
    for w in sys.argv[1:]:
        cfg = systems.CONFIGS[w]
        atoms = systems.config_atoms(w)
        shells = systems.make_basis_shells(atoms, cfg["basis"], tol=1e-10)
        mol, basis = gx.Molecule(atoms), gx.BasisSet(shells)
        t0 = time.time()
        mg = gx.MolGrid(mol, "Unpruned", 512, "MuraKnowles", cfg["grid"])
        t_grid = time.time() - t0
        rt = gx.RuntimeEnvironment(device=True)
        out = {}
        for ex in ("Device", "Host", "Device"):
            t0 = time.time()
            lb = gx.LoadBalancerFactory(ex, "Replicated").get_instance(rt, mol, mg, basis)
            n = lb.ntasks()
            out[ex] = time.time() - t0
            info = lb.task_info()
            print(f"{w}: LoadBalancer[{ex}] {out[ex]:.2f} s, {n} tasks, {int(info['npts'].sum())} points "
                  f"(molgrid {t_grid:.2f} s, {os.cpu_count()} host cores)", flush=True)
            del lb


if __name__ == "__main__":
    main()
