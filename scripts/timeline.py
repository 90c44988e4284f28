"""Kernel timeline of one eager (un-captured) train step through torch.profiler (CUPTI): per-stream busy time, overlap and a
text Gantt.  usage: python scripts/timeline.py --workload c2 > gpurun_out/timeline.txt"""
import argparse
import json
import os
import sys
import tempfile

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from torch.profiler import ProfilerActivity, profile

from bench import WORKLOADS
from splitvae_b200.engine import Engine

ap = argparse.ArgumentParser()
ap.add_argument("--workload", default="c2")
ap.add_argument("--graph", action="store_true", help="profile CUDA-graph replays of the captured step (trainer.StepRunner) instead of eager steps")
args = ap.parse_args()
model, H, B, patch, beta, alpha, desc = WORKLOADS[args.workload]
e = Engine(model=model, height=H, width=H, batch=B, beta=beta, alpha=alpha)
e.init_params(seed=5)
x = torch.rand(B, H, H, 6, device="cuda") * 2 - 1
if args.graph:
    from splitvae_b200.trainer import StepRunner
    runner = StepRunner(e, use_graph=True)
    runner.inputs.copy_(x)
    step = runner.step
else:
    step = lambda: e.train_step(x)
for _ in range(3):
    step()
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
    for _ in range(3):
        step()
    torch.cuda.synchronize()
path = os.path.join(tempfile.gettempdir(), "trace.json")
prof.export_chrome_trace(path)
ev = [v for v in json.load(open(path))["traceEvents"] if v.get("cat") == "kernel"]
ev.sort(key=lambda v: v["ts"])
# split into steps at the first-layer staging kernel
starts = [i for i, v in enumerate(ev) if "stage_first" in v["name"]]
steps = starts[::2]
lo, hi = steps[-1], len(ev)
k = ev[lo:hi]
t0 = k[0]["ts"]
end = max(v["ts"] + v["dur"] for v in k)
print(f"step span {end - t0:.1f} us, {len(k)} kernels, sum of durations {sum(v['dur'] for v in k):.1f} us")
streams = sorted({v["args"]["stream"] for v in k})
for s in streams:
    ks = [v for v in k if v["args"]["stream"] == s]
    print(f"stream {s}: {len(ks)} kernels, busy {sum(v['dur'] for v in ks):.1f} us, first {ks[0]['ts'] - t0:.1f}, last end {ks[-1]['ts'] + ks[-1]['dur'] - t0:.1f}")
# concurrency histogram (time with n kernels in flight)
pts = sorted([(v["ts"], 1) for v in k] + [(v["ts"] + v["dur"], -1) for v in k])
n, last, hist = 0, t0, {}
for t, d in pts:
    hist[n] = hist.get(n, 0.0) + (t - last)
    n += d
    last = t
print("time with n kernels in flight:", {a: round(b, 1) for a, b in sorted(hist.items())})
print("\nstart_us  dur_us  stream  grid  kernel")
for v in k:
    name = v["name"].replace("(anonymous namespace)::", "").split("(")[0].replace("sv::<unnamed>::", "").replace("sv::", "").replace("void ", "")[:38]
    g = v["args"].get("grid", "")
    print(f"{v['ts'] - t0:8.1f} {v['dur']:7.1f}  {streams.index(v['args']['stream'])}  {str(g):16s} {name}")
