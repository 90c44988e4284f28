"""Shrinks an `ncu --page source --csv` export: keeps only executed instructions and a few columns.
usage: ncu_source_compact.py in.csv out.csv"""
import csv
import sys

csv.field_size_limit(10 ** 9)
keep = ["Source", "# Samples", "Instructions Executed"]
with open(sys.argv[1]) as f, open(sys.argv[2], "w", newline="") as g:
    w = csv.writer(g)
    hdr = None
    idx = 0
    for r in csv.reader(f):
        if not r:
            continue
        if r[0] == "Kernel Name":
            w.writerow(r[:2]); hdr = None; idx = 0
            continue
        if hdr is None:
            hdr = r
            col = {h: i for i, h in enumerate(hdr)}
            stall = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
            w.writerow(["#"] + keep + ["top stalls"])
            continue
        ex = int(r[col["Instructions Executed"]] or 0)
        if ex > 0:
            st = sorted(((int(r[col[c]] or 0), c[6:]) for c in stall), reverse=True)[:3]
            w.writerow([idx] + [r[col[k]].strip() for k in keep] + [" ".join(f"{c}:{v}" for v, c in st if v)])
        idx += 1
