#!/bin/bash
OUT=gpurun_out/call8; mkdir -p $OUT
timeout 300 python scripts/timeline.py --workload c2 > $OUT/timeline.txt 2> $OUT/timeline.err; head -12 $OUT/timeline.txt; tail -3 $OUT/timeline.err
for cfg in SV_NS_SPLIT_WIDE=1; do
  echo "== $cfg"
  env $cfg timeout 300 python bench.py --steps 30 --warmup 5 --no-cpu-baseline 2>&1 | cut -c1-140
  env $cfg timeout 300 python scripts/bench_layers.py --workload c2 --filter decoder_x.d4 2>&1 | tail -4
done
