"""Where the roles of nsconv_kernel wait (SV_NS_DEBUG bit 2): per-CTA cycle sums of the TMA producer, the MMA-issuing thread and one
epilogue warp, medians over the CTAs of one launch.

    SV_BUILD_DEFINES=-DSV_NS_TRACE python splitvae_b200/build.py --force        # the counters are compiled out of the default build
    python scripts/ns_trace.py --workload c2 --layers decoder_x.d5:0 decoder_x.d4:0 decoder_x.d4:1 [--debug 0|1|2|3]
(layer:pass with pass 0 = forward, 1 = dgrad; --debug adds the SV_NS_DEBUG bits 0 / 1: epilogue drains only / no MMAs)"""
import argparse
import ctypes as C
import os
import sys

ap = argparse.ArgumentParser()
ap.add_argument("--workload", default="c2")
ap.add_argument("--layers", nargs="+", default=["decoder_x.d5:0", "decoder_x.d4:0", "decoder_x.d3:0", "decoder_x.d4:1", "decoder_x.d3:1"])
ap.add_argument("--debug", type=int, default=0)
ap.add_argument("--timing", action="store_true", help="bit 3: CTA start / end wall clocks (globaltimer) instead of the role sums")
ap.add_argument("--reps", type=int, default=10)
args = ap.parse_args()
os.environ["SV_NS_DEBUG"] = str((8 if args.timing else 4) | args.debug)
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

from bench import WORKLOADS
from splitvae_b200.engine import Engine

model, H, B, patch, beta, alpha, desc = WORKLOADS[args.workload]
e = Engine(model=model, height=H, width=H, batch=B, beta=beta, alpha=alpha)
e.init_params(seed=5)
x = torch.rand(B, H, H, 6, device="cuda") * 2 - 1
e.train_step(x)
torch.cuda.synchronize()
layers = {L.name.decode(): i for i, L in enumerate(e.debug_layers())}
buf = np.zeros(8192 * 8, dtype=np.uint64)
names = ["CTA lifetime", "weights landed (from CTA start)", "MMA thread: halo_full waits", "MMA thread: acc_empty waits", "MMA thread: issue + commit",
         "producer: halo_empty waits", "epilogue warp: acc_full waits", "epilogue warp: work"]
for spec in args.layers:
    name, p = spec.split(":")
    i = layers[name]
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e.debug_run_layer(i, int(p), 1, x)
    torch.cuda.synchronize()
    ev0.record()
    for _ in range(args.reps):
        e.debug_run_layer(i, int(p), 1, x)
    ev1.record()
    torch.cuda.synchronize()
    e.lib.sv_debug_halo_trace(buf.ctypes.data_as(C.c_void_p), 8192)
    t = buf.reshape(-1, 8)[:148].astype(np.int64)
    print(f"== {spec}  (SV_NS_DEBUG={os.environ['SV_NS_DEBUG']})  {ev0.elapsed_time(ev1) * 1e3 / args.reps:.1f} us per launch ({args.reps} back to back)")
    if args.timing:
        st, en, life = t[:, 1], t[:, 2], t[:, 0]
        print(f"   first CTA start -> last CTA end  {(en.max() - st.min()) / 1e3:8.1f} us   (last launch)")
        print(f"   CTA start spread                 {(st.max() - st.min()) / 1e3:8.1f} us   end spread {(en.max() - en.min()) / 1e3:.1f} us")
        print(f"   CTA lifetime                     median {np.median(en - st) / 1e3:.1f} us = {np.median(life):.0f} cycles -> {np.median(life / np.maximum(en - st, 1)):.3f} cycles/ns")
        print(f"   SMs used {len(set(t[:, 3].tolist()))}")
        continue
    for k, nm in enumerate(names):
        print(f"   {nm:34s} median {np.median(t[:, k]):9.0f}  p10 {np.percentile(t[:, k], 10):9.0f}  p90 {np.percentile(t[:, k], 90):9.0f} cycles")
