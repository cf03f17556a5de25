#!/usr/bin/env python
"""One-screen summary of a bench.py JSON line (GPU session logs)."""
import json
import sys

for f in sys.argv[1:]:
    try:
        d = json.loads([l for l in open(f) if l.startswith("{")][-1])
    except Exception as e:
        print(f, "unreadable:", e)
        continue
    r = d["roofline"]
    print(f"{d['config']['workload']}: {d['ms_per_step']:.1f} ms/step ({d['value'] / 1e6:.1f} Mpts/s), "
          f"e2e {d['e2e']['ms_per_step'] if d.get('e2e') else None}, ssf {d['ssf_weights_ms']:.0f} ms, n_gpus {d['n_gpus']}")
    print("  kernels ms:", {k: round(v, 1) for k, v in r["kernel_ms_per_step"].items()},
          "frac:", {k["kernel"].split(" ")[0]: round(k["frac"], 3) for k in r["per_kernel"]})
    print("  parity:", d.get("parity"))
    if d.get("cpu_baseline"):
        print("  cpu:", d["cpu_baseline"]["value"] / 1e6, "Mpts/s", d["cpu_baseline"]["cores"], "cores")
    for w, o in (d.get("others") or {}).items():
        print("  other", w, {k: o.get(k) for k in ("ms_per_step", "e2e_ms_per_step", "frac", "parity")})
    print("  clocks:", d.get("clocks"))
