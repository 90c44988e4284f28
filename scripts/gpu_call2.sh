#!/bin/bash
# GPU visit 2: new eval/checkpoint tests, knob sweeps, ncu --set full of the halo conv kernels (with source-level stalls)
TAG=${1:-call2}
OUT=gpurun_out/$TAG
mkdir -p $OUT
timeout 600 python -m pytest tests/test_gpu_eval.py -m gpu -x -q > $OUT/pytest_eval.log 2>&1; echo "pytest exit $?" >> $OUT/pytest_eval.log
tail -15 $OUT/pytest_eval.log
for cfg in SV_HWG_SPLITS=37 SV_HWG_SPLITS=56 SV_HWG_SPLITS=74 SV_HWG_SPLITS=111 "SV_HWG_SPLITS=74 SV_WG_CTAS=148" "SV_HWG_SPLITS=74 SV_WG_CTAS=222" \
           "SV_HWG_SPLITS=74 SV_S2_FWD_HALO=0" "SV_HWG_SPLITS=74 SV_HALO_TH=32 SV_HALO_TW=32"; do
  echo "== $cfg"
  env $cfg timeout 300 python bench.py --steps 30 --warmup 5 --no-cpu-baseline 2>&1 | cut -c1-140
done
timeout 900 ncu --set full --clock-control none --import-source on -k 'regex:halo_conv_kernel|halo4_kernel' -c 8 -f -o $OUT/prof_halo \
    python scripts/profile_step.py --workload c2 --steps 1 > $OUT/ncu_halo.log 2>&1
ncu -i $OUT/prof_halo.ncu-rep --page raw --csv > $OUT/prof_halo_raw.csv 2>/dev/null
ncu -i $OUT/prof_halo.ncu-rep --page source --csv > $OUT/prof_halo_source.csv 2>/dev/null
ncu -i $OUT/prof_halo.ncu-rep --page details --csv > $OUT/prof_halo_details.csv 2>/dev/null
python scripts/summarize_ncu_raw.py $OUT/prof_halo_raw.csv
for k in 0 1 6; do python scripts/ncu_top_stalls.py $OUT/prof_halo_source.csv $k 25 > $OUT/stalls_$k.txt 2>&1; done
[ $(stat -c %s $OUT/prof_halo.ncu-rep 2>/dev/null || echo 0) -gt 30000000 ] && rm -f $OUT/prof_halo.ncu-rep
ls -la $OUT
