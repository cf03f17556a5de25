#!/usr/bin/env python
"""Top stall-sample SASS instructions of an .ncu-rep (source page), with main stall reason.
usage: python tools/ncu_sass_hot.py <rep> [topN]"""
import csv, subprocess, sys
rep = sys.argv[1]; top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hi = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
hdr = rows[hi]; body = rows[hi + 1:]
ix = {h: i for i, h in enumerate(hdr)}
stall_cols = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
tot = 0; recs = []
for n, r in enumerate(body):
    if len(r) < len(hdr): continue
    try: s = int(r[ix["# Samples"]])
    except ValueError: continue
    tot += s
    st = sorted(((int(r[ix[c]] or 0), c) for c in stall_cols), reverse=True)[:2]
    recs.append((s, n, r[ix["Source"]].strip()[:90], int(r[ix["Instructions Executed"]] or 0), st))
print("total samples", tot)
for s, n, src, ie, st in sorted(recs, reverse=True)[:top]:
    print(f"{100*s/tot:5.1f}% #{n:5d} exec={ie:9d} {src:90s} {st[0][1]}={st[0][0]} {st[1][1]}={st[1][0]}")
