"""Runs N eager (un-captured) train steps of one workload so that `ncu` sees every kernel of the step.

    ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches.csv \
        python scripts/profile_step.py --workload c2 --steps 2
"""
import argparse
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from bench import WORKLOADS
from splitvae_b200.engine import Engine

ap = argparse.ArgumentParser()
ap.add_argument("--workload", default="c2")
ap.add_argument("--steps", type=int, default=2)
ap.add_argument("--batch", type=int, default=0)
args = ap.parse_args()
model, H, B, patch, beta, alpha, desc = WORKLOADS[args.workload]
B = args.batch or B
e = Engine(model=model, height=H, width=H, batch=B, beta=beta, alpha=alpha)
e.init_params(seed=5)
x = torch.rand(B, H, H, 6, device="cuda") * 2 - 1
torch.cuda.synchronize()
for i in range(args.steps):
    l0 = e.launch_count
    e.train_step(x)
    torch.cuda.synchronize()
    print(f"step {i}: {e.launch_count - l0} launches, total={e.scalars()['total']:.3f}", flush=True)
