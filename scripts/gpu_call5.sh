#!/bin/bash
OUT=gpurun_out/call5; mkdir -p $OUT
SV_HALO_TRACE=1 timeout 300 python scripts/halo_trace.py --workload c2 > $OUT/halo_trace.txt 2>&1
cat $OUT/halo_trace.txt
