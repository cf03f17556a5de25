#!/usr/bin/env python
"""SSF weights on the device for the given workloads: kernel ms (CUDA events) and, with --check, the error against
the oracle's host algorithm on a random sample of tasks.  usage: python tools/ssf_time.py [--check] ubiquitin water833"""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))
sys.path.insert(0, os.path.join(ROOT, "tests"))


def main():
    from gauxc_b200.driver import System
    check = "--check" in sys.argv
    for w in [a for a in sys.argv[1:] if not a.startswith("--")]:
        t0 = time.time()
        s = System(w, device=True)
        raw = s.lb.export_tasks() if check else None
        t1 = time.time()
        ms = s.modify_weights()
        t2 = time.time()
        line = f"{w}: setup {t1 - t0:.1f} s, ssf kernel {ms:.1f} ms, modify_weights wall {t2 - t1:.2f} s, npts {s.npts_local}"
        if check:
            import pyoracle as orc
            from test_gpu_parity import ssf_sample_error
            err, n = ssf_sample_error(orc, s.atoms, raw, s.lb.export_tasks(), 60)
            line += f", max |w - w_oracle| {err:.2e} over {n} points"
        print(line, flush=True)


if __name__ == "__main__":
    main()
