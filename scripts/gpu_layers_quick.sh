#!/bin/bash
# Layer-level check of one planner change on one B200: the per-layer tensor-core tests, the isolated layer timings, then short bench lines.
#   usage: bash scripts/gpu_layers_quick.sh <tag> "ENV=a" "ENV=b" ...   ("-" = no overrides)
TAG=$1; shift
OUT=gpurun_out/$TAG; mkdir -p $OUT
timeout 600 python -m pytest tests/test_gpu_tc_layers.py tests/test_gpu_split_layers.py -x -q > $OUT/pytest_layers.log 2>&1; echo "pytest exit $?" >> $OUT/pytest_layers.log; tail -5 $OUT/pytest_layers.log
timeout 300 python scripts/bench_layers.py --workload c2 > $OUT/layers_c2.txt 2> $OUT/layers_c2.err; grep -E "e1|e2|sum" $OUT/layers_c2.txt | head -12
i=0
for SETTING in "$@"; do
  [ "$SETTING" = "-" ] && SETTING=""
  echo "== [$SETTING]"
  env $SETTING timeout 300 python bench.py --steps 30 --warmup 5 --no-cpu-baseline --no-other-workloads --no-fast-mode > $OUT/bench_$i.json 2> $OUT/bench_$i.err
  python -c "import sys,json; d=json.load(open('$OUT/bench_$i.json')); print('ms/step', round(d['ms_per_step'],4), 'e2e', round(d['e2e']['ms_per_step'],4))" || tail -5 $OUT/bench_$i.err
  i=$((i+1))
done
