"""Per-launch digest of an `ncu --set full ... ; ncu -i X.ncu-rep --page raw --csv` export:
duration, tensor-pipe utilisation, DRAM bytes, L2 throughput.  usage: summarize_ncu_raw.py raw.csv"""
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
hdr, units, data = rows[0], rows[1], rows[2:]
col = {h: i for i, h in enumerate(hdr)}


def f(r, name, default=0.0):
    try:
        return float(r[col[name]].replace(",", ""))
    except Exception:
        return default


def scale(name, to):
    u = units[col[name]]
    k = {"ns": 1e-3, "nsecond": 1e-3, "us": 1.0, "usecond": 1.0, "ms": 1e3, "msecond": 1e3, "s": 1e6, "second": 1e6}
    b = {"byte": 1e-6, "Kbyte": 1e-3, "Mbyte": 1.0, "Gbyte": 1e3}
    return (k if to == "us" else b).get(u, 1.0)


print(f"{'#':>3} {'kernel':34s} {'grid':>14s} {'us':>8s} {'tensor%':>8s} {'dramR MB':>9s} {'dramW MB':>9s} {'dram%':>6s} {'L2%':>6s} {'SM%':>6s} {'regs':>5s}")
for i, r in enumerate(data):
    name = r[col["Kernel Name"]].split("(")[0].replace("sv::<unnamed>::", "").replace("sv::", "")[:34]
    grid = r[col["Grid Size"]] if "Grid Size" in col else ""
    us = f(r, "gpu__time_duration.sum") * scale("gpu__time_duration.sum", "us")
    tp = f(r, "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed")
    dr = f(r, "dram__bytes_read.sum") * scale("dram__bytes_read.sum", "MB")
    dw = f(r, "dram__bytes_write.sum") * scale("dram__bytes_write.sum", "MB")
    dp = f(r, "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed")
    l2 = f(r, "lts__throughput.avg.pct_of_peak_sustained_elapsed")
    sm = f(r, "sm__throughput.avg.pct_of_peak_sustained_elapsed")
    regs = f(r, "launch__registers_per_thread")
    print(f"{i:3d} {name:34s} {grid:>14s} {us:8.1f} {tp:8.1f} {dr:9.1f} {dw:9.1f} {dp:6.1f} {l2:6.1f} {sm:6.1f} {regs:5.0f}")
