#!/bin/bash
# Weak-scaling sweep on one box (gpurun --gpus 8): bench.py at N = 1, 2, 4, 8 (256 images per GPU), NCCL init / algorithm lines kept.
TAG=${1:-r02_scale}
OUT=gpurun_out/$TAG
mkdir -p $OUT
python bench.py --gpus 1 --steps 40 --warmup 5 --no-cpu-baseline --no-fast-mode --no-other-workloads > $OUT/bench_n1.json 2> $OUT/bench_n1.err
for N in 2 4 8; do
  NCCL_DEBUG=INFO NCCL_DEBUG_SUBSYS=INIT,TUNING python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $((29600 + N)) \
      bench.py --gpus $N --steps 40 --warmup 5 --no-cpu-baseline --no-fast-mode $([ $N -lt 8 ] && echo --no-other-workloads) > $OUT/bench_n$N.json 2> $OUT/bench_n$N.err
  python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $((29700 + N)) bench.py --gpus $N --steps 40 --warmup 5 --no-cpu-baseline --no-fast-mode --no-other-workloads --nccl > $OUT/bench_n${N}_nccl.json 2> /dev/null
  grep -E "NCCL INFO (AllReduce|Connected|comm|Channel|NVLS|Using)" $OUT/bench_n$N.err | grep -v "Channel [0-9]*/[0-9]* :" | sort | uniq -c | sort -rn | head -12 > $OUT/nccl_n$N.txt
done
python - <<PY
import json
base = None
for n in (1, 2, 4, 8):
    try:
        d = json.load(open("$OUT/bench_n%d.json" % n))
    except Exception as ex:
        print(n, "failed", ex); continue
    base = base or d["value"]
    print(f"N={n}: {d['value']:.0f} img/s  {d['ms_per_step']:.3f} ms/step  efficiency {d['value'] / (n * base):.3f}  e2e {d['e2e']['value']:.0f}", d.get("other_workloads", {}).get("c4", {}).get("value"))
PY
python - <<PY
import json
for n in (2, 4, 8):
    try:
        d = json.load(open("$OUT/bench_n%d_nccl.json" % n)); print(f"NCCL path N={n}: {d['value']:.0f} img/s  {d['ms_per_step']:.3f} ms/step")
    except Exception as ex:
        print(n, "nccl line missing", ex)
PY
