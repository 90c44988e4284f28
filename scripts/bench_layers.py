"""Times every tensor-core layer launch (fwd / dgrad / wgrad) of one workload in isolation with CUDA events.
Env knobs read by the planner (SV_HALO_TW, SV_HALO_TH, SV_NO_HALO, ...) can be swept from the shell."""
import argparse
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from bench import WORKLOADS
from splitvae_b200.engine import Engine

ap = argparse.ArgumentParser()
ap.add_argument("--workload", default="c2")
ap.add_argument("--reps", type=int, default=10)
ap.add_argument("--filter", default="")
ap.add_argument("--precision", default="bf16x3")
args = ap.parse_args()
model, H, B, patch, beta, alpha, desc = WORKLOADS[args.workload]
e = Engine(model=model, height=H, width=H, batch=B, beta=beta, alpha=alpha, precision=args.precision)
e.init_params(seed=5)
x = torch.rand(B, H, H, 6, device="cuda") * 2 - 1
e.train_step(x)
torch.cuda.synchronize()
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
tot = 0.0
for i, L in enumerate(e.debug_layers()):
    name = L.name.decode()
    if args.filter and args.filter not in name:
        continue
    macs = B * L.Ho * L.Wo * L.Co * L.kh * L.kw * L.Ci
    for p, pname, ok in ((0, "fwd", L.tc_fwd), (1, "dgrad", L.tc_dgrad), (2, "wgrad", L.tc_wgrad)):
        if not ok:
            continue
        ts = []
        for r in range(args.reps):
            flush.zero_()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            e.debug_run_layer(i, p, 1, x)
            b.record()
            torch.cuda.synchronize()
            ts.append(a.elapsed_time(b) * 1e3)
        ts.sort()
        us = ts[len(ts) // 2]
        tot += us
        print(f"{name:28s} {pname:5s} {us:8.1f} us  {2 * macs / us / 1e6:8.1f} TFLOP/s  ({L.kh}x{L.kw} s{L.stride} {L.Ci}->{L.Co} @{L.Ho}x{L.Wo})", flush=True)
print(f"sum {tot:.1f} us")
