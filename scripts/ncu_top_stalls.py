"""Top SASS instructions by warp-stall samples from `ncu -i X.ncu-rep --page source --csv`.
usage: ncu_top_stalls.py source.csv [kernel_index] [top_n]"""
import csv
import sys

csv.field_size_limit(10 ** 9)
rows = list(csv.reader(open(sys.argv[1])))
kidx = int(sys.argv[2]) if len(sys.argv) > 2 else 0
topn = int(sys.argv[3]) if len(sys.argv) > 3 else 40
starts = [i for i, r in enumerate(rows) if r and r[0] == "Kernel Name"]
starts.append(len(rows))
blk = rows[starts[kidx]:starts[kidx + 1]]
print(blk[0][1])
hdr = blk[1]
col = {h: i for i, h in enumerate(hdr)}
ins = blk[2:]
stall_cols = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
tot = sum(int(r[col["# Samples"]] or 0) for r in ins)
print("total samples", tot, "instructions", len(ins))
order = sorted(range(len(ins)), key=lambda i: -int(ins[i][col["# Samples"]] or 0))[:topn]
for i in sorted(order):
    r = ins[i]
    n = int(r[col["# Samples"]] or 0)
    st = sorted(((int(r[col[c]] or 0), c[6:]) for c in stall_cols), reverse=True)[:3]
    st = " ".join(f"{c}:{v}" for v, c in st if v)
    print(f"{i:5d} {100.0 * n / max(tot, 1):5.1f}% exec {r[col['Instructions Executed']]:>8s}  {r[col['Source']].strip()[:70]:70s} {st}")
