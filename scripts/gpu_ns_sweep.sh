#!/bin/bash
TAG=${1:-sweep}
OUT=gpurun_out/$TAG
mkdir -p $OUT
timeout 900 python -m pytest tests/test_gpu_tc_layers.py tests/test_gpu_parity.py -x -q > $OUT/pytest_gpu.log 2>&1; echo "pytest exit $?" >> $OUT/pytest_gpu.log
tail -4 $OUT/pytest_gpu.log
for cfg in "default" "SV_NS_NOSPLIT=1" "SV_NS_STAGES=2" "SV_NS_STAGES=3"; do
  echo "== $cfg"
  if [ "$cfg" = "default" ]; then E=""; else E="$cfg"; fi
  env $E SV_TC_VERBOSE=1 timeout 300 python scripts/bench_layers.py --workload c2 --filter decoder_x. > $OUT/layers_$cfg.txt 2> $OUT/layers_$cfg.err
  grep -E "d3|d4|d5" $OUT/layers_$cfg.txt | grep -v wgrad
done
timeout 600 python bench.py --steps 30 --warmup 5 --no-cpu-baseline > $OUT/bench_c2.json 2> $OUT/bench_c2.err; cut -c1-200 $OUT/bench_c2.json; tail -3 $OUT/bench_c2.err
