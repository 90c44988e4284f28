#!/bin/bash
# GPU visit: parity tests with the new planner paths (falling back one feature at a time to localise a failure),
# bench A/B per feature, per-layer timings, ncu launch list.
TAG=${1:-call1}
OUT=gpurun_out/$TAG
mkdir -p $OUT
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $OUT/smi.txt 2>&1
timeout 900 python -m pytest tests -m gpu -x -q > $OUT/pytest_gpu.log 2>&1; RC=$?; echo "pytest exit $RC" >> $OUT/pytest_gpu.log
tail -15 $OUT/pytest_gpu.log
if [ $RC -ne 0 ]; then
  for cfg in SV_S2_FWD_HALO=0 SV_FIRST_PAIR=0 SV_S2_DGRAD_HALO=0 SV_OLD_REDUCE=1 "SV_S2_FWD_HALO=0 SV_FIRST_PAIR=0 SV_S2_DGRAD_HALO=0 SV_OLD_REDUCE=1"; do
    env $cfg timeout 900 python -m pytest tests -m gpu -q > $OUT/pytest_fallback.log 2>&1; echo "== $cfg: pytest exit $? $(tail -1 $OUT/pytest_fallback.log)"
    grep -E "^FAILED|^ERROR" $OUT/pytest_fallback.log | head -8
  done
fi
timeout 600 python bench.py --steps 30 --warmup 5 > $OUT/bench_c2.json 2> $OUT/bench_c2.err; cut -c1-330 $OUT/bench_c2.json; tail -3 $OUT/bench_c2.err
for cfg in SV_S2_FWD_HALO=0 SV_FIRST_PAIR=0 SV_S2_DGRAD_HALO=0 SV_OLD_REDUCE=1 "SV_HALO_TH=32 SV_HALO_TW=32" SV_WG_CTAS=148 SV_HWG_SPLITS=74 SV_HWG_SPLITS=296; do
  echo "== $cfg"
  env $cfg timeout 300 python bench.py --steps 30 --warmup 5 --no-cpu-baseline 2>&1 | cut -c1-140
done
SV_TC_VERBOSE=1 timeout 300 python scripts/bench_layers.py --workload c2 > $OUT/layers_c2.txt 2> $OUT/layers_c2.err
cat $OUT/layers_c2.txt | grep -E "e1|e2|e3|sum"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file $OUT/launches_c2.csv \
    python scripts/profile_step.py --workload c2 --steps 2 > $OUT/profile_step.log 2>&1
python scripts/summarize_launches.py $OUT/launches_c2.csv $(grep -o 'step 1: [0-9]*' $OUT/profile_step.log | grep -o '[0-9]*$') --all > $OUT/launches_c2_summary.txt 2>&1
tail -26 $OUT/launches_c2_summary.txt
