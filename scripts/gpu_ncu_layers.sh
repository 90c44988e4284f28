#!/bin/bash
# ncu --set full (+ source-level stall samples) of single layer passes timed by scripts/bench_layers.py.
#   usage: bash scripts/gpu_ncu_layers.sh <tag> <layer filter> <kernel regex> <launches of that regex in one train step> <count>
TAG=$1; FILTER=$2; KREGEX=$3; SKIP=$4; COUNT=$5
OUT=gpurun_out/$TAG; mkdir -p $OUT
timeout 600 ncu --set full --import-source on --clock-control none -k "regex:$KREGEX" --launch-skip $SKIP -c $COUNT -f -o /tmp/layers \
    python scripts/bench_layers.py --workload c2 --filter "$FILTER" --reps 1 > $OUT/ncu.log 2>&1
tail -3 $OUT/ncu.log
ncu -i /tmp/layers.ncu-rep --page raw --csv > $OUT/raw.csv 2>/dev/null
ncu -i /tmp/layers.ncu-rep --page source --csv > /tmp/src.csv 2>/dev/null
for i in $(seq 0 $((COUNT-1))); do python scripts/ncu_top_stalls.py /tmp/src.csv $i 45 > $OUT/stalls_$i.txt 2>&1; done
python scripts/summarize_ncu_raw.py $OUT/raw.csv > $OUT/summary.txt 2>&1; cat $OUT/summary.txt
ls -la $OUT
