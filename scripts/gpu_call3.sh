#!/bin/bash
# GPU visit 3: ncu --set full of the halo conv kernels, digest only (the raw source page is ~80 MB: kept in /tmp on the box)
TAG=${1:-call3}
OUT=gpurun_out/$TAG
mkdir -p $OUT
for cfg in "SV_HWG_SPLITS=18" "SV_HWG_SPLITS=26" "SV_HWG_SPLITS=37 SV_WG_CTAS=148" "SV_HWG_SPLITS=37 SV_WG_CTAS=74"; do
  echo "== $cfg"
  env $cfg timeout 300 python bench.py --steps 30 --warmup 5 --no-cpu-baseline 2>&1 | cut -c1-140
done
timeout 900 ncu --set full --clock-control none --import-source on -k 'regex:halo_conv_kernel|halo4_kernel' -c 8 -f -o /tmp/prof_halo \
    python scripts/profile_step.py --workload c2 --steps 1 > $OUT/ncu_halo.log 2>&1
ncu -i /tmp/prof_halo.ncu-rep --page raw --csv > $OUT/prof_halo_raw.csv 2>/dev/null
ncu -i /tmp/prof_halo.ncu-rep --page source --csv > /tmp/prof_halo_source.csv 2>/dev/null
ncu -i /tmp/prof_halo.ncu-rep --page details --csv > $OUT/prof_halo_details.csv 2>/dev/null
for k in 0 1 4 6; do python scripts/ncu_top_stalls.py /tmp/prof_halo_source.csv $k 45 > $OUT/stalls_$k.txt 2>&1; done
python - <<'PY' > $OUT/details_digest.txt 2>&1
import csv
rows=list(csv.reader(open("gpurun_out/call3/prof_halo_details.csv")))
hdr=rows[0]; col={h:i for i,h in enumerate(hdr)}
want=("Duration","Elapsed Cycles","SM Active Cycles","Achieved Occupancy","Theoretical Occupancy","Block Limit","Waves Per SM","Registers Per","Shared Memory Config","Dynamic Shared","Achieved Active Warps","Issued Warp","No Eligible","Eligible Warps","Active Warps Per","One or More Eligible","Executed Ipc","Issue Slots Busy","L2 Cache Throughput","DRAM Throughput","Memory Throughput","Compute (SM)","L1/TEX Hit","L2 Hit","Mem Busy","Max Bandwidth","Mem Pipes","Avg. Active Threads","Stall")
for r in rows[1:]:
    if r[col["ID"]] in ("0","1","4","6") and any(w in r[col["Metric Name"]] for w in want):
        print(r[col["ID"]], r[col["Section Name"]][:28].ljust(28), r[col["Metric Name"]][:44].ljust(44), r[col["Metric Value"]], r[col["Metric Unit"]])
PY
ls -la $OUT
