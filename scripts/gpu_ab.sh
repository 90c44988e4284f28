#!/bin/bash
TAG=${1:-ab}   # usage: bash scripts/gpu_ab.sh <tag> [VAR=VALUE ...]: tests, bench, bench under each knob, layer timings, ncu launch list
OUT=gpurun_out/$TAG
mkdir -p $OUT
timeout 900 python -m pytest tests -m gpu -x -q > $OUT/pytest_gpu.log 2>&1; echo "pytest exit $?" >> $OUT/pytest_gpu.log
tail -8 $OUT/pytest_gpu.log
timeout 600 python bench.py --steps 30 --warmup 5 > $OUT/bench_c2.json 2> $OUT/bench_c2.err; cut -c1-330 $OUT/bench_c2.json; tail -3 $OUT/bench_c2.err
for cfg in "$@"; do
  [ "$cfg" = "$TAG" ] && continue
  echo "== $cfg"
  env $cfg timeout 300 python bench.py --steps 30 --warmup 5 --no-cpu-baseline 2>&1 | cut -c1-140
done
SV_TC_VERBOSE=1 timeout 300 python scripts/bench_layers.py --workload c2 > $OUT/layers_c2.txt 2> $OUT/layers_c2.err
cat $OUT/layers_c2.txt
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file $OUT/launches_c2.csv \
    python scripts/profile_step.py --workload c2 --steps 2 > $OUT/profile_step.log 2>&1
python scripts/summarize_launches.py $OUT/launches_c2.csv $(grep -o 'step 1: [0-9]*' $OUT/profile_step.log | grep -o '[0-9]*$') --all > $OUT/launches_c2_summary.txt 2>&1
tail -26 $OUT/launches_c2_summary.txt
