#!/bin/bash
# per config: per-layer parity (LGVae 64x64), bench, wgrad layer timings
for cfg in "$@"; do
  echo "== $cfg"
  env $cfg timeout 300 python -m pytest tests/test_gpu_tc_layers.py -m gpu -q -k "test_tc_layers_match_reference and lgvae-64" 2>&1 | tail -1
  env $cfg timeout 300 python bench.py --steps 40 --warmup 5 --no-cpu-baseline 2>/dev/null | cut -c1-140
  env $cfg timeout 300 python scripts/bench_layers.py --workload c2 --filter decoder_x.d 2>/dev/null | grep wgrad | grep "d3\|d4\|d5"
done
