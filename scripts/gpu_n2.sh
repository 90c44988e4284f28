#!/bin/bash
# 2-GPU data-parallel bench (NCCL all-reduce of the gradient buckets captured in the step graph) + the reference arm under torchrun
OUT=gpurun_out/n2; mkdir -p $OUT
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 30 --warmup 5 --no-cpu-baseline > $OUT/bench_n2.json 2> $OUT/bench_n2.err
echo "exit $?"; cut -c1-400 $OUT/bench_n2.json; tail -5 $OUT/bench_n2.err
