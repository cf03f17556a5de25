#!/usr/bin/env python
"""Stall samples of an .ncu-rep grouped by SASS index ranges (warp roles of a warp-specialised kernel).
usage: python tools/ncu_role_samples.py <rep> [name:lo-hi ...]   (no ranges: print landmark instructions)"""
import csv, subprocess, sys, collections
rep = sys.argv[1]
roles = []
for a in sys.argv[2:]:
    n, r = a.split(":"); lo, hi = r.split("-"); roles.append((n, int(lo), int(hi)))
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
h0 = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
hdr = rows[h0]; body = [r for r in rows[h0 + 1:] if len(r) >= len(hdr)]
ix = {h: i for i, h in enumerate(hdr)}
stall_cols = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
if not roles:
    marks = ["DMMA", "LDGSTS", "UBLKCP", "STG.E", "LDG.E", "SYNCS.PHASECHK", "BAR.SYNC", "ATOMG", "RED", "SETMAXREG", "USETMAXREG", "EXIT", "MUFU"]
    last = None
    for n, r in enumerate(body):
        src = r[ix["Source"]]
        m = next((k for k in marks if k in src), None)
        if m and (m != last or m in ("SYNCS.PHASECHK", "BAR.SYNC", "SETMAXREG", "USETMAXREG", "EXIT")):
            print(n, src.strip()[:100], "samples", r[ix["# Samples"]])
        if m: last = m
    sys.exit(0)
tot = sum(int(r[ix["# Samples"]] or 0) for r in body)
for name, lo, hi in roles:
    sub = body[lo:hi + 1]
    s = sum(int(r[ix["# Samples"]] or 0) for r in sub)
    by = collections.Counter()
    for r in sub:
        for c in stall_cols:
            by[c] += int(r[ix[c]] or 0)
    spin = sum(int(r[ix["# Samples"]] or 0) for i, r in enumerate(sub)
               if ("BRA" in r[ix["Source"]] and i > 0 and "TRYWAIT" in sub[i - 1][ix["Source"]]) or "TRYWAIT" in r[ix["Source"]])
    print(f"{name:10s} #{lo}-{hi}: {100*s/tot:5.1f}% of samples; mbarrier spin {100*spin/max(s,1):5.1f}% of role; "
          + ", ".join(f"{c[6:]}={100*v/max(s,1):.0f}%" for c, v in by.most_common(5)))
