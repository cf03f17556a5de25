#!/usr/bin/env python
"""Measured parity of the Device path against the CPU oracle on the five BASELINE.json configurations
(run on a B200; writes one line per configuration).  Small systems are compared on their full grids,
the large ones on every `stride`-th task of their real task lists (what tests/test_gpu_parity.py
asserts with thresholds; here the actual deviations are recorded).
usage (under gpurun): python tools/parity_report.py > gpurun_out/parity.txt"""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import gauxc_b200 as gx  # noqa: E402
import pyoracle as orc  # noqa: E402
from gauxc_b200.driver import System  # noqa: E402

orc.init_blas()
CASES = [("water", 1), ("benzene", 1), ("taxol", 300), ("ubiquitin", 2500), ("water833", 6000)]
if len(sys.argv) > 1:
    CASES = [c for c in CASES if c[0] in sys.argv[1:]]

for workload, stride in CASES:
    t0 = time.time()
    try:
        s = System(workload, device=True)
        if stride > 1:
            full = s.lb.export_tasks()
            nt = len(full["npts"])
            pick = np.unique(np.r_[np.arange(0, nt, stride), full["nbe"].argmax(), full["npts"].argmax()])
            poff = np.r_[0, np.cumsum(full["npts"])]
            soff = np.r_[0, np.cumsum(full["nshells"])]
            pts = np.concatenate([full["points"][poff[t]:poff[t + 1]] for t in pick])
            w = np.concatenate([full["weights"][poff[t]:poff[t + 1]] for t in pick])
            sl = np.concatenate([full["shell_lists"][soff[t]:soff[t + 1]] for t in pick])
            s.lb.set_tasks(full["npts"][pick], full["iParent"][pick], full["dist_nearest"][pick], pts, w,
                           full["nshells"][pick], sl, False)
        gx.MolecularWeightsFactory("Device", "Default", "SSF").get_instance().modify_weights(s.lb)
        tasks = s.lb.export_tasks()
        integ = gx.XCIntegratorFactory("Device").get_instance(gx.Functional(s.func_name), s.lb)
        exc, vxc = integ.eval_exc_vxc(s.P)
        nel = integ.stats()["n_el"]
        ref = orc.exc_vxc(s.basis.flat(), s.nbf, s.P, tasks, s.func_name)
        print(f"{workload:10s} {s.func_name:6s} {s.grid:14s} nbf {s.nbf:6d} tasks {len(tasks['npts']):7d} "
              f"points {int(tasks['npts'].sum()):9d} max nbe {int(tasks['nbe'].max()):5d}  "
              f"EXC {exc:.12f}  |dEXC| {abs(exc - ref['exc']):.2e}  max|dVXC| {np.abs(vxc - ref['vxc']).max():.2e}  "
              f"|dN_el| {abs(nel - ref['nel']):.2e}  VXC symmetric {bool(np.array_equal(vxc, vxc.T))}  "
              f"({'full grid' if stride == 1 else 'every %dth task' % stride}, {time.time() - t0:.0f} s)", flush=True)
        del integ, s
    except Exception as e:  # keep going: one line per configuration
        print(f"{workload:10s} FAILED: {e}", flush=True)
