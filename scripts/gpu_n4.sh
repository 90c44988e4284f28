#!/bin/bash
OUT=gpurun_out/n4; mkdir -p $OUT
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus 4 --steps 30 --warmup 5 > $OUT/bench_n4.json 2> $OUT/bench_n4.err
echo "exit $?"; cut -c1-300 $OUT/bench_n4.json; tail -3 $OUT/bench_n4.err
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29514 bench.py --impl reference --gpus 4 --steps 2 --warmup 1 2>/dev/null | cut -c1-200
