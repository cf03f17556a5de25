#!/usr/bin/env python
"""DMMA slot model of the fused kernel (CPU only): executed m8n8k4 slots per SM sub-partition against the
algorithmic 2 npts nbe^2 flop of X = P_sub B, from the real task list of a workload.  Quantifies what the
tile / chunk / stage granularity costs:
  rows    : 8-row blocks, rounded to 16 rows per warp except in the split-K mode of tiles <= 64 points
            (LDA kernel: exact)
  columns : 64-column chunks, a chunk costs a full chunk whatever its fill
  K       : 16-row stages; LDA stops at the diagonal chunk
usage: python tools/dmma_slot_model.py taxol ubiquitin > profiles/r01_dmma_slot_model.txt"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from gauxc_b200.driver import System  # noqa: E402

for wl in sys.argv[1:]:
    s = System(wl, device=False)
    t = s.lb.export_tasks()
    npts_t = t["npts"].astype(np.int64)
    nbe_t = t["nbe"].astype(np.int64)
    lda = s.func_name.upper() in ("SVWN5", "LDA", "SPW92", "VWN5")
    res = {}
    for exact_rows in (0, 1):
        for narrow in (0, 1):
            ideal = model = 0.0
            for npts, nbe in zip(npts_t, nbe_t):
                nk = (nbe + 15) // 16
                nn = (nbe + 63) // 64
                lastw = 0.5 if (narrow and nbe - (nn - 1) * 64 <= 32) else 1.0
                if lda:
                    stages = sum(min(nk, 4 * (c + 1)) for c in range(nn - 1)) + lastw * min(nk, 4 * nn)
                else:
                    stages = (nn - 1) * nk + lastw * nk
                full, rem = divmod(npts, 128)

                def blocks(n):
                    even = lambda mi: 2 * ((mi + 1) // 2)  # noqa: E731
                    if n <= 64:
                        mi = (n + 7) // 8
                        return mi if exact_rows else even(mi)
                    return 8 + even((n - 64 + 7) // 8)

                mb = full * 16 + (blocks(rem) if rem else 0)
                model += 16 * 8 * mb * stages             # clocks of one sub-partition's DMMA pipe
                ideal += (2.0 * npts * nbe * nbe / 128.0) * (0.5 if lda else 1.0)
            res[(exact_rows, narrow)] = model / ideal
    f = 148 * 1.965e9
    print(f"{wl}: tasks {len(npts_t)}, points {npts_t.sum()}, algorithmic X time at the DMMA peak "
          f"{ideal / f * 1e3:.1f} ms{' (triangular)' if lda else ''}")
    print(f"  executed / algorithmic slots: as built (GGA kernel) {res[(0, 0)]:.3f}; exact row blocks in short "
          f"tiles (LDA kernel) {res[(1, 0)]:.3f}; + half-width last chunk {res[(1, 1)]:.3f} "
          f"(even rows + half-width last chunk {res[(0, 1)]:.3f})")
