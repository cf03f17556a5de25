#!/usr/bin/env python
"""Per-kernel totals of an `ncu --metrics gpu__time_duration.sum --csv` launch list.
usage: python tools/ncu_launch_summary.py <launches.csv>
The number of eval_exc_vxc calls in the capture is COUNTED (one reduce_partials_kernel launch per call), not assumed."""
import collections, csv, sys
rows = [r for r in csv.reader(open(sys.argv[1])) if len(r) > 10]
hdr = rows[0]; ix = {h: i for i, h in enumerate(hdr)}
scale = {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3}
agg = collections.defaultdict(lambda: [0, 0.0])
for r in rows[1:]:
    if r[ix["Metric Name"]] != "gpu__time_duration.sum":
        continue
    k = r[ix["Kernel Name"]].replace("gxb::<unnamed>::", "").split("(")[0]
    agg[k][0] += 1
    agg[k][1] += float(r[ix["Metric Value"]].replace(",", "")) * scale[r[ix["Metric Unit"]]]
tot = sum(v[1] for v in agg.values())
calls = max(1, sum(v[0] for k, v in agg.items() if "reduce_partials" in k))
print(f"# {sys.argv[1]}: {sum(v[0] for v in agg.values())} launches, {tot:.2f} ms under ncu (serialised, cold cache), "
      f"{calls} eval_exc_vxc call(s) in the capture (= reduce_partials_kernel launches)")
print(f"{'kernel':50s} {'launches':>8s} {'total ms':>10s} {'ms/call':>9s} {'share':>7s}")
for k, v in sorted(agg.items(), key=lambda x: -x[1][1]):
    per = v[1] / calls if ("ssf" not in k and "probe" not in k) else v[1]
    print(f"{k[:50]:50s} {v[0]:8d} {v[1]:10.2f} {per:9.2f} {100*v[1]/tot:6.1f}%")
