#!/bin/bash
# bench under each "VAR=VALUE[ VAR=VALUE]" argument (no tests): planner knob sweeps
for cfg in "$@"; do
  echo "== $cfg"
  env $cfg timeout 300 python bench.py --steps 40 --warmup 5 --no-cpu-baseline 2>/dev/null | cut -c1-140
done
