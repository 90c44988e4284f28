#!/usr/bin/env python
"""Per-tensor gradient error of a tensor-core precision mode against the repo's own fp32 mode (which is oracle-checked to
2e-4) at BASELINE.json's FULL sizes, where the fp64 CPU oracle would take minutes.

    python scripts/grad_error_fullsize.py [--workload c2] [--precision bf16x3 bf16] [--batch 256]

Prints scalars, then rel-L2 and cosine per gradient tensor (worst first)."""
import argparse
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from splitvae_b200.engine import Engine  # noqa: E402

WL = {"c1": ("lgvae", 32, 64, 1.0), "c2": ("lgvae", 64, 256, 120.0), "c3": ("lggmvae", 32, 256, 40.0), "c4": ("lggmvae", 64, 256, 120.0)}


def inputs_for(model, H, B, seed=0):
    g = torch.Generator(device="cuda").manual_seed(seed)
    k = torch.randint(0, 256, (B, H, H, 6), generator=g, device="cuda")
    x = (k.float() / 255.0 * 2 - 1).contiguous()
    eg = torch.randn(B, 128, generator=g, device="cuda")
    el = torch.randn(B, 128, generator=g, device="cuda")
    u = torch.rand(B, 30, generator=g, device="cuda").clamp_(1e-6, 1 - 1e-6) if model != "lgvae" else None
    return x, eg, el, u


def step(e, x, eg, el, u):
    e.forward(x, eg, el, u)
    e.loss_fwd_bwd(x)
    for s in range(len(e.segments)):
        e.backward_segment(s)
    torch.cuda.synchronize()
    return e.scalars(), e.grads.clone()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--workload", default="c2")
    ap.add_argument("--precision", nargs="+", default=["bf16x3", "bf16"])
    ap.add_argument("--batch", type=int, default=0)
    ap.add_argument("--top", type=int, default=8)
    a = ap.parse_args()
    model, H, B, beta = WL[a.workload]
    B = a.batch or B
    x, eg, el, u = inputs_for(model, H, B)
    ref = Engine(model=model, height=H, width=H, batch=B, beta=beta, alpha=40.0, precision="fp32")
    ref.init_params(seed=7)
    rsc, rg = step(ref, x, eg, el, u)
    print(f"{a.workload} {model} {H}x{H} B={B}: fp32-mode scalars {rsc}")
    for prec in a.precision:
        try:
            e = Engine(model=model, height=H, width=H, batch=B, beta=beta, alpha=40.0, precision=prec)
        except Exception as ex:   # an older library without this mode
            print(f"precision {prec}: unavailable ({ex})")
            continue
        e.params.copy_(ref.params)
        e.params_updated()
        sc, g = step(e, x, eg, el, u)
        srel = {k: abs(sc[k] - v) / max(abs(v), 1e-30) for k, v in rsc.items()}
        rows = []
        for name, shape, off, cnt in e.table:
            p, q = g[off:off + cnt].double(), rg[off:off + cnt].double()
            den = q.norm().item()
            if den < 1e-12:
                continue
            rows.append(((p - q).norm().item() / den, float(torch.dot(p, q) / (p.norm() * q.norm() + 1e-300)), name))
        rows.sort(reverse=True)
        print(f"precision {prec}: scalar rel errors " + ", ".join(f"{k} {v:.1e}" for k, v in srel.items()))
        print(f"  worst gradient rel-L2 {rows[0][0]:.3e} ({rows[0][2]}), median {rows[len(rows) // 2][0]:.3e}")
        for r, c, n in rows[:a.top]:
            print(f"    {r:.3e}  cos {c:.6f}  {n}")
        del e


if __name__ == "__main__":
    main()
