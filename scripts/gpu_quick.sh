#!/bin/bash
# Quick A/B on one B200 (gpurun): optional GPU tests, then one short device-resident bench line per environment setting.
#   usage: bash scripts/gpu_quick.sh <tag> <run tests: 0|1> "ENV1=a ENV2=b" "ENV3=c" ...     ("-" = no overrides)
TAG=$1; TESTS=$2; shift 2
OUT=gpurun_out/$TAG; mkdir -p $OUT
if [ "$TESTS" = "1" ]; then
  timeout 900 python -m pytest tests -m gpu -x -q > $OUT/pytest_gpu.log 2>&1; echo "pytest exit $?" >> $OUT/pytest_gpu.log; tail -5 $OUT/pytest_gpu.log
fi
i=0
for SETTING in "$@"; do
  [ "$SETTING" = "-" ] && SETTING=""
  echo "== [$SETTING]"
  env $SETTING timeout 300 python bench.py --steps 30 --warmup 5 --no-cpu-baseline --no-other-workloads --no-fast-mode $BENCH_ARGS > $OUT/bench_$i.json 2> $OUT/bench_$i.err
  python -c "import sys,json; d=json.load(open('$OUT/bench_$i.json')); print('ms/step', round(d['ms_per_step'],4), 'e2e', round(d['e2e']['ms_per_step'],4))" || tail -5 $OUT/bench_$i.err
  i=$((i+1))
done
