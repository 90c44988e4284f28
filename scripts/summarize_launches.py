"""Summarises an `ncu --metrics gpu__time_duration.sum --csv` launch list: per-launch table of the LAST step
and per-kernel totals.  usage: summarize_launches.py launches.csv launches_per_step [--all]"""
import csv
import sys
from collections import OrderedDict

path, per_step = sys.argv[1], int(sys.argv[2])
rows = []
with open(path) as f:
    lines = [l for l in f if not l.startswith("==")]
for r in csv.DictReader(lines):
    if r.get("Metric Name") != "gpu__time_duration.sum":
        continue
    v = float(r["Metric Value"].replace(",", ""))
    unit = r["Metric Unit"]
    us = v / 1000.0 if unit in ("ns", "nsecond") else v if unit in ("us", "usecond") else v * 1000.0 if unit in ("ms", "msecond") else v
    rows.append((r["Kernel Name"].split("(")[0], r.get("Grid Size", ""), r.get("Block Size", ""), us))
step = rows[-per_step:]
total = sum(r[3] for r in step)
print(f"launches in file: {len(rows)}; last step: {len(step)} launches, {total:.1f} us (serialised, cold-cache)")
if "--all" in sys.argv:
    for i, (k, g, b, us) in enumerate(step):
        print(f"{i:4d} {us:9.1f} us  {100*us/total:5.1f}%  {k[:60]:60s} grid {g} block {b}")
agg = OrderedDict()
for k, g, b, us in step:
    a = agg.setdefault(k, [0, 0.0])
    a[0] += 1
    a[1] += us
print("\nper kernel (last step):")
for k, (n, us) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f"{us:10.1f} us {100*us/total:5.1f}%  x{n:3d}  {k[:80]}")
