#!/usr/bin/env python
"""EXC gradient on the Device: parity against the reference's fixtures / the oracle and device times.
usage (GPU box): python tools/grad_report.py > profiles/rNN_exc_grad.txt"""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))
sys.path.insert(0, os.path.join(ROOT, "tests"))
import gauxc_b200 as gx  # noqa: E402
from gauxc_b200 import systems  # noqa: E402
from gauxc_b200.driver import System  # noqa: E402
import pyoracle as orc  # noqa: E402  (checker only)
from conftest import make_lb  # noqa: E402


def centers(atoms, basis):
    xyz = np.array([a[1:] for a in atoms])
    return np.linalg.norm(basis.flat()[5][:, None, :] - xyz[None, :, :], axis=2).argmin(1).astype(np.int32)


def main():
    orc.init_blas()
    gold = np.load(os.path.join(ROOT, "tests", "golden", "benzene_exc_grad.npz"))
    print("# fixture parity (reference tests/xc_integrator.cxx:276-297 accepts rms < 1e-8)")
    for name, func in (("benzene_svwn5_cc-pvdz_ufg_ssf", "SVWN5"), ("benzene_pbe0_cc-pvdz_ufg_ssf", "PBE0")):
        atoms, shells, P, _, _ = systems.golden_system(name)
        for s in shells:
            s["tol"] = np.finfo(float).eps
        _, basis, lb = make_lb(atoms, shells, "UltraFineGrid", "Unpruned", normalize=False, device=True)
        gx.MolecularWeightsFactory("Device", "Default", "SSF").get_instance().modify_weights(lb)
        tasks = lb.export_tasks()
        integ = gx.XCIntegratorFactory("Device").get_instance(gx.Functional(func), lb)
        coords = np.array([a[1:] for a in atoms])
        for key, wd in (("EXC_GRAD_HELLFEY", False), ("EXC_GRAD_FULL", True)):
            g = integ.eval_exc_grad(P, len(atoms), include_weight_derivatives=wd)
            t0 = time.time()
            g = integ.eval_exc_grad(P, len(atoms), include_weight_derivatives=wd)
            dt = time.time() - t0
            o = orc.exc_grad(basis.flat(), centers(atoms, basis), coords, basis.nbf(), P, tasks, func, wd)
            ref = gold[f"{name}:{key}"]
            print(f"{name} {func} {key}: rms vs fixture {np.linalg.norm(g - ref) / np.sqrt(3 * len(atoms)):.2e} "
                  f"max vs fixture {np.abs(g - ref).max():.2e} max vs oracle {np.abs(g - o).max():.2e} "
                  f"|sum over atoms| {np.abs(g.sum(0)).max():.2e} device call {dt * 1e3:.1f} ms "
                  f"(local work {integ.stats()['local_work_ms']:.1f} ms)")
    print("# taxol def2-SVP PBE SuperFine, full grid, one GPU: device time per eval_exc_grad call")
    s = System("taxol", device=True)
    gx.MolecularWeightsFactory("Device", "Default", "SSF").get_instance().modify_weights(s.lb)
    integ = gx.XCIntegratorFactory("Device").get_instance(gx.Functional(s.func_name), s.lb)
    for wd in (False, True):
        g = integ.eval_exc_grad(s.P, len(s.atoms), include_weight_derivatives=wd)
        t0 = time.time()
        g = integ.eval_exc_grad(s.P, len(s.atoms), include_weight_derivatives=wd)
        dt = time.time() - t0
        print(f"taxol include_weight_derivatives={wd}: {dt * 1e3:.1f} ms per call (local work "
              f"{integ.stats()['local_work_ms']:.1f} ms), max|g| {np.abs(g).max():.6f} "
              f"|sum over atoms| {np.abs(g.sum(0)).max():.2e}")
    t0 = time.time()
    exc, _ = integ.eval_exc_vxc(s.P)
    exc, _ = integ.eval_exc_vxc(s.P)
    print(f"taxol eval_exc_vxc on the same integrator afterwards: EXC {exc:.10f}")


if __name__ == "__main__":
    main()
