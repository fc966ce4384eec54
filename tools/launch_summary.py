#!/usr/bin/env python3
"""Sum an `ncu --metrics gpu__time_duration.sum --csv` launch list per kernel.  Usage: launch_summary.py file.csv"""
import csv, sys
rows = [r for r in csv.reader(open(sys.argv[1])) if len(r) > 5]
hdr = [i for i, r in enumerate(rows) if "Kernel Name" in r][0]
h = rows[hdr]; ki = h.index("Kernel Name"); vi = h.index("Metric Value"); ui = h.index("Metric Unit")
agg = {}
for r in rows[hdr + 1:]:
    v = float(r[vi].replace(",", "")); u = r[ui]
    v = v / 1e6 if u == "ns" else (v / 1e3 if u == "us" else v)
    agg.setdefault(r[ki][:90], []).append(v)
tot = sum(sum(v) for v in agg.values())
for k, v in sorted(agg.items(), key=lambda kv: -sum(kv[1])):
    print(f"  {sum(v):9.3f} ms  {100 * sum(v) / tot:5.1f} %  x{len(v):3d}  {k}")
