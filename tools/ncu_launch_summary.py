#!/usr/bin/env python3
"""Condenses an `ncu --metrics gpu__time_duration.sum --csv` launch list into per-kernel totals and shares.
Usage: tools/ncu_launch_summary.py launches.csv "<header comment>" > profiles/rNN_launches_summary.csv"""
import csv, sys, collections
rows = list(csv.DictReader(l for l in open(sys.argv[1]) if not l.startswith("==")))
agg = collections.OrderedDict()
for r in rows:
    if r["Metric Name"] != "gpu__time_duration.sum":
        continue
    name = r["Kernel Name"].split("(")[0].replace("void ", "").split("<")[0]
    v = float(r["Metric Value"].replace(",", ""))
    v *= {"ns": 1e-3, "us": 1.0, "ms": 1e3}.get(r["Metric Unit"], 1e-3)
    a = agg.setdefault(name, [0, 0.0])
    a[0] += 1
    a[1] += v
total = sum(a[1] for a in agg.values())
print("# " + sys.argv[2])
print("# per-launch times are cold-cache and serialised: compare SHARES.")
print("kernel,launches,mean_us_per_launch,total_us,share")
for n, (c, t) in agg.items():
    print("%s,%d,%.1f,%.1f,%.3f" % (n, c, t / c, t, t / total))
