#!/bin/bash
# Round-end evidence run on ONE B200 (gpurun): parity tests, bench lines (default line carries all workloads), per-layer timings,
# ncu launch list, ncu --set full digest + DRAM traffic per kernel.   usage: bash scripts/gpu_profile.sh [tag]
TAG=${1:-r02}
OUT=gpurun_out/$TAG
mkdir -p $OUT
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active --format=csv > $OUT/smi.txt 2>&1
timeout 1200 python -m pytest tests -m gpu -q > $OUT/pytest_gpu.log 2>&1; echo "pytest exit $?" >> $OUT/pytest_gpu.log
tail -4 $OUT/pytest_gpu.log
timeout 900 python bench.py --steps 50 --warmup 10 > $OUT/bench_c2.json 2> $OUT/bench_c2.err; cut -c1-300 $OUT/bench_c2.json; tail -3 $OUT/bench_c2.err
timeout 600 python bench.py --steps 50 --warmup 10 --precision bf16 --no-cpu-baseline --no-other-workloads > $OUT/bench_c2_bf16.json 2> $OUT/bench_c2_bf16.err; cut -c1-200 $OUT/bench_c2_bf16.json
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > $OUT/bench_reference.json 2>&1; cut -c1-300 $OUT/bench_reference.json
timeout 300 python scripts/bench_layers.py --workload c2 > $OUT/layers_c2.txt 2> $OUT/layers_c2.err
tail -1 $OUT/layers_c2.txt
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file $OUT/launches_c2.csv \
    python scripts/profile_step.py --workload c2 --steps 2 > $OUT/profile_step.log 2>&1
python scripts/summarize_launches.py $OUT/launches_c2.csv $(grep -o 'step 1: [0-9]*' $OUT/profile_step.log | grep -o '[0-9]*$') --all > $OUT/launches_c2_summary.txt 2>&1
tail -28 $OUT/launches_c2_summary.txt
if [ -z "$SKIP_FULL" ]; then
  timeout 1500 ncu --set full --clock-control none -k 'regex:igemm|halo_conv_kernel|halo_wgrad_kernel|wgrad_kernel|nsconv_kernel|pconv_kernel|pixel_loss_kernel|adam_kernel|pack_kernel|upsample2x|wgrad_reduce_vec|colsum_multi_partial' \
      -c 125 -f -o /tmp/prof_full python scripts/profile_step.py --workload c2 --steps 1 > $OUT/ncu_full.log 2>&1
  ncu -i /tmp/prof_full.ncu-rep --page raw --csv > $OUT/prof_full_raw.csv 2>/dev/null
  python scripts/summarize_ncu_raw.py $OUT/prof_full_raw.csv > $OUT/ncu_full_summary.txt 2>&1
  head -5 $OUT/ncu_full_summary.txt; wc -l $OUT/ncu_full_summary.txt
fi
ls -la $OUT | head -30
