#!/usr/bin/env python
"""DRAM traffic per kernel per step from an `ncu --set full` raw export (dram__bytes_read.sum + dram__bytes_write.sum summed over the
kernel's launches of ONE eager step) -> profiles/ncu_traffic.json, which bench.py reports as roofline.traffic.

    python scripts/make_ncu_traffic.py gpurun_out/r02/prof_full_raw.csv c2 [--merge profiles/ncu_traffic.json]"""
import csv
import json
import sys

path, wl = sys.argv[1], sys.argv[2]
rows = list(csv.reader(open(path)))
hdr, units, data = rows[0], rows[1], rows[2:]
col = {h: i for i, h in enumerate(hdr)}
scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}


def val(r, name):
    try:
        return float(r[col[name]].replace(",", "")) * scale.get(units[col[name]], 1.0)
    except Exception:
        return 0.0


out = {}
if data and "pack_kernel" in data[0][col["Kernel Name"]]:
    data = data[1:]          # the engine's initial full re-pack precedes the step
for r in data:
    name = r[col["Kernel Name"]].split("(")[0].split("::")[-1].split("<")[0].strip().replace("void ", "")
    d = out.setdefault(name, {"launches": 0, "dram_bytes": 0.0, "us": 0.0})
    d["launches"] += 1
    d["dram_bytes"] += val(r, "dram__bytes_read.sum") + val(r, "dram__bytes_write.sum")
res = {k: int(v["dram_bytes"]) for k, v in out.items()}
res["_launches"] = {k: v["launches"] for k, v in out.items()}
res["_step_total_bytes"] = int(sum(v["dram_bytes"] for v in out.values()))
merged = {}
if "--merge" in sys.argv:
    try:
        merged = json.load(open(sys.argv[sys.argv.index("--merge") + 1]))
    except Exception:
        merged = {}
merged[wl] = res
json.dump(merged, sys.stdout, indent=1)
print()
