#!/usr/bin/env python
"""Summarises an `ncu --metrics gpu__time_duration.sum --csv` launch list: per-kernel launch count,
total / mean duration and share.  usage: launch_summary.py launches.csv [first_id last_id]"""
import csv, sys, collections
rows = []
with open(sys.argv[1]) as f:
    lines = [l for l in f if not l.startswith("==")]
for r in csv.DictReader(lines):
    if r.get("Metric Name") != "gpu__time_duration.sum":
        continue
    v = float(r["Metric Value"].replace(",", ""))
    unit = r["Metric Unit"]
    us = v / 1e3 if unit in ("ns", "nsecond") else v * 1e3 if unit in ("ms", "msecond") else v
    rows.append((int(r["ID"]), r["Kernel Name"].split("(")[0], us))
if len(sys.argv) > 3:
    rows = [r for r in rows if int(sys.argv[2]) <= r[0] <= int(sys.argv[3])]
agg = collections.OrderedDict()
for _, k, us in rows:
    a = agg.setdefault(k, [0, 0.0]); a[0] += 1; a[1] += us
tot = sum(a[1] for a in agg.values())
print("%-60s %7s %12s %10s %6s" % ("kernel", "n", "total_us", "mean_us", "share"))
for k, (n, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print("%-60s %7d %12.1f %10.2f %5.1f%%" % (k[:60], n, t, t / n, 100 * t / tot))
print("%-60s %7d %12.1f" % ("TOTAL", len(rows), tot))
