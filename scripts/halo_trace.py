"""Per-CTA phase timeline of the halo convolution kernel (SV_HALO_TRACE=1): which phase of a CTA's life takes the time.

    SV_HALO_TRACE=1 python scripts/halo_trace.py --workload c2 --layers encoder_x.e1:0 encoder_x.e2:0 encoder_x.e2:1 decoder_x.d5:1
"""
import argparse
import ctypes as C
import os
import sys

os.environ.setdefault("SV_HALO_TRACE", "1")
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

from bench import WORKLOADS
from splitvae_b200.engine import Engine

ap = argparse.ArgumentParser()
ap.add_argument("--workload", default="c2")
ap.add_argument("--layers", nargs="+", default=["encoder_x.e1:0", "encoder_x.e2:0", "encoder_x.e2:1", "decoder_x.d5:1"])
args = ap.parse_args()
model, H, B, patch, beta, alpha, desc = WORKLOADS[args.workload]
e = Engine(model=model, height=H, width=H, batch=B, beta=beta, alpha=alpha)
e.init_params(seed=5)
x = torch.rand(B, H, H, 6, device="cuda") * 2 - 1
e.train_step(x)
torch.cuda.synchronize()
layers = {L.name.decode(): i for i, L in enumerate(e.debug_layers())}
buf = np.zeros(8192 * 8, dtype=np.uint64)
for spec in args.layers:
    name, p = spec.split(":")
    i = layers[name]
    for _ in range(2):
        e.debug_run_layer(i, int(p), 1, x)
    torch.cuda.synchronize()
    n = e.lib.sv_debug_halo_trace(buf.ctypes.data_as(C.c_void_p), 8192)
    t = buf.reshape(-1, 8)[:n].astype(np.int64)
    t = t[t[:, 1] > 0]
    start, setup, halo, mma, accdone, epi = (t[:, k] for k in (1, 2, 3, 4, 5, 6))
    gt = t[:, 7]
    ok = (epi > start) & (gt >= gt.max() - 300000)      # the buffer is never cleared: keep the last launch only (300 us window)
    t, start, setup, halo, mma, accdone, epi, gt = (a[ok] for a in (t, start, setup, halo, mma, accdone, epi, gt))
    med = lambda a: float(np.median(a))
    print(f"== {spec}: {len(t)} CTAs traced; kernel span {(gt.max() - gt.min()) / 1e3:.1f} us (first to last CTA start)")
    print(f"   setup            {med(setup - start):9.0f} cycles   (barrier init, TMEM alloc, first sync)")
    print(f"   halo wait        {med(halo - setup):9.0f} cycles   (TMA halo load until landed)")
    print(f"   MMA issue        {med(mma - halo):9.0f} cycles   (all taps issued, incl. weight-ring waits)")
    print(f"   MMA drain        {med(accdone - mma):9.0f} cycles   (commit -> epilogue sees the accumulators)")
    print(f"   epilogue         {med(epi - accdone):9.0f} cycles")
    print(f"   CTA lifetime     {med(epi - start):9.0f} cycles   p10 {np.percentile(epi - start, 10):.0f}  p90 {np.percentile(epi - start, 90):.0f}")
    sm = t[:, 0]
    per_sm = np.bincount(sm.astype(np.int64))
    print(f"   CTAs per SM      max {per_sm.max()}  min {per_sm[per_sm > 0].min()}  SMs used {(per_sm > 0).sum()}")
    # concurrency on one SM: how many CTAs overlap in time on SM of the first CTA
    s0 = sm[0]
    sel = np.argsort(start[sm == s0])
    st, en = start[sm == s0][sel], epi[sm == s0][sel]
    print("   SM %d timeline (start, end) rel cycles: %s" % (s0, ", ".join(f"({a - st[0]},{b - st[0]})" for a, b in list(zip(st, en))[:10])))
