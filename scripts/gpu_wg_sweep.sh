#!/bin/bash
TAG=${1:-wg}
OUT=gpurun_out/$TAG
mkdir -p $OUT
timeout 900 python -m pytest tests -m gpu -x -q > $OUT/pytest_gpu.log 2>&1; echo "pytest exit $?" >> $OUT/pytest_gpu.log
tail -4 $OUT/pytest_gpu.log
for cfg in "SV_WG_GPC=1 SV_WG_CTAS=148" "SV_WG_GPC=1 SV_WG_CTAS=296" "SV_WG_GPC=2 SV_WG_CTAS=148" "SV_WG_GPC=8 SV_WG_CTAS=296"; do
  echo "== $cfg"
  env $cfg timeout 300 python scripts/bench_layers.py --workload c2 --filter _x. > "$OUT/layers_$cfg.txt" 2>&1
  grep -E "wgrad" "$OUT/layers_$cfg.txt" | grep -v "d4\|d5\|d3"
  env $cfg timeout 300 python bench.py --steps 30 --warmup 5 --no-cpu-baseline 2>/dev/null | cut -c1-160
done
