#!/usr/bin/env python
"""One EXC gradient call (with weight derivatives) on a workload, for ncu launch lists.
usage: python tools/grad_once.py taxol"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402
import gauxc_b200 as gx  # noqa: E402
from gauxc_b200.driver import System  # noqa: E402

w = sys.argv[1] if len(sys.argv) > 1 else "taxol"
s = System(w, device=True)
gx.MolecularWeightsFactory("Device", "Default", "SSF").get_instance().modify_weights(s.lb)
integ = gx.XCIntegratorFactory("Device").get_instance(gx.Functional(s.func_name), s.lb)
g = integ.eval_exc_grad(s.P, len(s.atoms), include_weight_derivatives=True)
print(w, "max|g|", float(np.abs(g).max()), "local work ms", integ.stats()["local_work_ms"])
