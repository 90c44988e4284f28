#!/bin/bash
# One GPU-box visit: parity tests, bench line, per-layer timings, ncu launch list, ncu --set full of the GEMM kernels.
# usage (from the repo root, under gpurun): bash scripts/gpu_check.sh [tag]
TAG=${1:-run}
OUT=gpurun_out/$TAG
mkdir -p $OUT
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active --format=csv > $OUT/smi.txt 2>&1
if [ -z "$SKIP_TESTS" ]; then
  timeout 900 python -m pytest tests -m gpu -x -q > $OUT/pytest_gpu.log 2>&1; echo "pytest exit $?" >> $OUT/pytest_gpu.log
  tail -5 $OUT/pytest_gpu.log
fi
timeout 600 python bench.py --steps 30 --warmup 5 > $OUT/bench_c2.json 2> $OUT/bench_c2.err; tail -c 2500 $OUT/bench_c2.json
timeout 300 python scripts/bench_layers.py --workload c2 > $OUT/layers_c2.txt 2>&1
if [ -z "$SKIP_NCU" ]; then
  timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file $OUT/launches_c2.csv \
      python scripts/profile_step.py --workload c2 --steps 2 > $OUT/profile_step.log 2>&1
  python scripts/summarize_launches.py $OUT/launches_c2.csv $(grep -o 'step 1: [0-9]*' $OUT/profile_step.log | grep -o '[0-9]*$') --all > $OUT/launches_c2_summary.txt 2>&1
  NG=$(grep -c 'igemm_kernel\|halo_conv_kernel\|wgrad_kernel' $OUT/launches_c2_summary.txt)
  timeout 900 ncu --set full --clock-control none -k 'regex:igemm_kernel|halo_conv_kernel|halo_wgrad_kernel|wgrad_kernel$' \
      -s ${NCU_SKIP:-64} -c ${NCU_COUNT:-64} -f -o $OUT/prof_gemm python scripts/profile_step.py --workload c2 --steps 2 > $OUT/ncu_full.log 2>&1
  timeout 300 ncu --set full --clock-control none --import-source on -k 'regex:pixel_loss_kernel|adam_kernel|pack_kernel' \
      -s 3 -c 3 -f -o $OUT/prof_hbm python scripts/profile_step.py --workload c2 --steps 2 > $OUT/ncu_full_hbm.log 2>&1
  ncu -i $OUT/prof_gemm.ncu-rep --page raw --csv > $OUT/prof_gemm_raw.csv 2>/dev/null
  ncu -i $OUT/prof_hbm.ncu-rep --page raw --csv > $OUT/prof_hbm_raw.csv 2>/dev/null
  [ $(stat -c %s $OUT/prof_gemm.ncu-rep 2>/dev/null || echo 0) -gt 40000000 ] && rm -f $OUT/prof_gemm.ncu-rep
fi
ls -la $OUT
