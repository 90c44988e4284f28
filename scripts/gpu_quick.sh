#!/bin/bash
# Quick GPU visit: parity tests (with fallbacks to localise a failure), bench line, per-layer timings.
TAG=${1:-quick}
OUT=gpurun_out/$TAG
mkdir -p $OUT
timeout 900 python -m pytest tests -m gpu -x -q > $OUT/pytest_gpu.log 2>&1; RC=$?; echo "pytest exit $RC" >> $OUT/pytest_gpu.log
tail -30 $OUT/pytest_gpu.log
if [ $RC -ne 0 ]; then
  SV_NO_NSCONV=1 timeout 900 python -m pytest tests -m gpu -x -q > $OUT/pytest_gpu_nons.log 2>&1; echo "pytest(no nsconv) exit $?" >> $OUT/pytest_gpu_nons.log
  tail -5 $OUT/pytest_gpu_nons.log
  SV_NO_NSCONV=1 SV_NO_SPLITK=1 timeout 900 python -m pytest tests -m gpu -x -q > $OUT/pytest_gpu_nons_nosk.log 2>&1; echo "pytest(no nsconv, no splitk) exit $?" >> $OUT/pytest_gpu_nons_nosk.log
  tail -5 $OUT/pytest_gpu_nons_nosk.log
fi
timeout 600 python bench.py --steps 30 --warmup 5 ${BENCH_ARGS} > $OUT/bench_c2.json 2> $OUT/bench_c2.err; tail -c 1500 $OUT/bench_c2.json; tail -5 $OUT/bench_c2.err
SV_TC_VERBOSE=1 timeout 300 python scripts/bench_layers.py --workload c2 > $OUT/layers_c2.txt 2> $OUT/layers_c2.err
grep -E "d4|d5|e4_mean|d1 " $OUT/layers_c2.txt; tail -1 $OUT/layers_c2.txt
