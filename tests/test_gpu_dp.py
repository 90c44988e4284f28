"""GPU, >= 2 devices: the data-parallel train step ON HARDWARE equals the oracle at the global batch: N = 2 ranks, each
one graph-captured step with the bucketed NCCL all-reduce and the per-segment optimizer overlap inside the graph, against the CPU
oracle run on the whole batch of 2b images (tests/dp_worker.py).  Skipped on a one-GPU box; `gpurun --gpus 2` runs it
(log kept under profiles/)."""
import json
import os
import socket
import subprocess
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


@pytest.mark.skipif(not torch.cuda.is_available() or torch.cuda.device_count() < 2, reason="needs two GPUs (gpurun --gpus 2)")
# (bf16x3 cases use >= 32 images of 32x32 / 8 of 64x64 in total: with fewer, ONE flipped ReLU unit of a small layer - probability
#  ~1e-6 per unit at the pair precision - moves that layer's gradient by 1/sqrt(batch * units) > 1 %)
@pytest.mark.parametrize("path", ["nccl", "nvls"])
@pytest.mark.parametrize("model,H,b,p,beta,prec", [("lgvae", 32, 4, 4, 40.0, "fp32"), ("lgvae", 32, 16, 4, 40.0, "bf16x3"),
                                                   ("lggmvae", 32, 16, 4, 40.0, "bf16x3"), ("lgvae", 64, 4, 8, 120.0, "bf16x3")])
def test_two_gpu_graph_step_equals_oracle_at_global_batch(model, H, b, p, beta, prec, path):
    """path nccl: bucketed NCCL all-reduce + per-segment Adam; path nvls: the fused multimem reduce-scatter + Adam(shard) + all-gather."""
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
           "--master-port", str(_free_port()), os.path.join(ROOT, "tests", "dp_worker.py"), model, str(H), str(b), str(p), str(beta), prec, path]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600, cwd=ROOT)
    lines = [l for l in r.stdout.split("\n") if l.startswith("{")]
    assert lines, (r.stdout[-2000:], r.stderr[-3000:])
    d = json.loads(lines[-1])
    print(d)
    assert r.returncode == 0 and d["ok"], (d["messages"], r.stderr[-2000:])
    assert d["world"] == 2 and d["iterations"] == 1
