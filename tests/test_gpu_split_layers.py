"""bf16x3 forward, layer by layer: after ONE sv_forward every stored activation is a bf16 pair (hi + lo); each layer's device
output must equal the CPU oracle's fp64 layer applied to the device's own input pair, to ~2^-16 (the pair's precision plus the dropped lo*lo term) - a
layer whose lo operands, split weight sections or lo stores were wrong would sit at the single-bf16 level (2^-9 = 2e-3).

Checked per layer: the three-MMA product (hi*hi + lo*hi + hi*lo), the [W_hi | W_lo] weight packs of every view (plain, pixel
pairs, first-layer window / in-pixel pairs), the epilogues' lo stores, split-K finishes, and the paired bilinear resize, the
reparameterisation / gumbel / residual-add glue that writes pairs."""
import numpy as np
import pytest
import torch

from oracle import splitvae_oracle as O
from helpers import make_case, make_engine, to_dev

pytestmark = pytest.mark.gpu

FUSED = {"e4_mean": [("e4_mean", None), ("e4_sd", "softplus")],
         "y_block.0": [("y_block.0", "elu"), ("e1", "elu")],
         "h_top_dense": [("h_top_dense", "elu"), ("z_prior_mean", None), ("z_prior_sig", "softplus")],
         "z_mean": [("z_mean", None), ("z_sig", "softplus")]}
ACT = {"e1": "relu", "e2": "relu", "e3": "relu", "d1": "relu", "d2": "relu", "d3": "relu", "d4": "relu", "d5": None,
       "h_block.0": "elu", "h_block.1": "elu", "h_block.2": "elu", "y_block.2": "elu", "y_dense": None}


def _pair(e, ptr, ptr_lo, elems, dt):
    v = e.debug_view(ptr, elems, dt).double()
    if ptr_lo:
        v = v + e.debug_view(ptr_lo, elems, 1).double()
    return v.cpu()


def _rel(a, b):
    return float((a - b).norm() / max(float(b.norm()), 1e-30))


def check_layers(model, H, B, p=4, verbose=False):
    params, batch = make_case(model, H, B, p)
    e = make_engine(model, H, B, "bf16x3", 40.0)
    e.load_params(params)
    x = to_dev(batch["inputs"])
    e.forward(x, to_dev(batch["eps_g"]), to_dev(batch["eps_l"]), to_dev(batch["u"]) if model != "lgvae" else None)
    torch.cuda.synchronize()
    P = {k: torch.tensor(v, dtype=torch.float64) for k, v in params.items()}
    worst, rows, outs = 0.0, [], {}
    for L in e.debug_layers():
        name = L.name.decode()
        prefix, first = name.split(".", 1)
        parts = FUSED[first] if first in FUSED else [(first, ACT[first])]
        dense = L.kh == 1 and L.kw == 1 and L.Hi == 1
        if not L.in_:                                  # first conv: the image batch, as the bf16 pair the staging kernel writes
            coff = 3 if prefix == "encoder_x_hat" else 0
            xin = torch.tensor(batch["inputs"][..., coff:coff + 3], dtype=torch.float32)
            hi = xin.bfloat16().float()
            vin = (hi + (xin - hi).bfloat16().float()).double() if L.split_fwd else hi.double()
        else:
            vin = _pair(e, L.in_, L.in_lo, L.in_elems, L.in_dt).view(-1, L.in_ld)[:, L.in_coff:L.in_coff + L.Ci]
            vin = vin.reshape(B, L.Hi, L.Wi, L.Ci) if not dense else vin.reshape(B, L.Ci)
        ref = []
        for pn, act in parts:
            w, b = P[f"{prefix}.{pn}.kernel"], P[f"{prefix}.{pn}.bias"]
            w32 = w.float()
            whi = w32.bfloat16().float()
            wq = (whi + (w32 - whi).bfloat16().float()).double() if L.split_fwd else whi.double()   # the device's weight pair
            pre = (vin @ wq + b) if dense else O.conv2d_same(vin, wq, b, L.stride)
            ref.append(O._act(pre, act))
        ref = torch.cat([r.reshape(-1, r.shape[-1]) for r in ref], dim=1)
        out = _pair(e, L.out, L.out_lo, L.out_elems, L.out_dt).view(-1, L.out_ld)[:, :L.Co]
        outs[name] = out
        r = _rel(out, ref)
        # stored pair: 2^-17; single bf16 (d5 input path) / fp32 outputs of single-bf16 products are exact up to accumulation order
        rows.append((name, r, "pair" if L.out_lo else ("f32" if L.out_dt == 0 else "bf16")))
        worst = max(worst, r)
        tol = 5e-5 if (L.out_lo or L.out_dt == 0) else 4e-3
        if verbose:
            print(f"   {name:30s} rel-L2 {r:.2e}  ({rows[-1][2]})")
        assert r < tol, (name, r, rows)
    return e, batch, rows, worst


@pytest.mark.parametrize("model,H,B", [("lgvae", 32, 4), ("lgvae", 64, 3), ("lgvae", 64, 4), ("lggmvae", 32, 5), ("lggmvae", 64, 2), ("lgvae", 32, 130)])
def test_split_forward_layers(model, H, B):
    e, batch, rows, worst = check_layers(model, H, B)
    print(f"{model} H={H} B={B}: {len(rows)} layers, worst rel-L2 {worst:.2e}")


@pytest.mark.parametrize("model,H,B", [("lgvae", 32, 4), ("lggmvae", 32, 4)])
def test_split_forward_glue(model, H, B):
    """resize pairs, z pair, y pair, h = e1 + h_top pair: consumer input == fp64 function of the producer output."""
    e, batch, rows, _ = check_layers(model, H, B)
    info = {L.name.decode(): L for L in e.debug_layers()}
    for dec in ("decoder_x", "decoder_x_hat"):
        for src, dst in (("d2", "d3"), ("d3", "d4"), ("d4", "d5")):
            S, D = info[f"{dec}.{src}"], info[f"{dec}.{dst}"]
            lo_res = _pair(e, S.out, S.out_lo, S.out_elems, S.out_dt).view(B, S.Ho, S.Wo, S.Co)
            up = O.resize2x(lo_res)
            got = _pair(e, D.in_, D.in_lo, D.in_elems, D.in_dt).view(B, D.Hi, D.Wi, D.Ci)
            r = _rel(got, up)
            assert r < (2e-5 if D.in_lo else 4e-3), (dec, dst, r)
    z = torch.cat([e.output("z_x"), e.output("z_x_hat")], dim=1).double().cpu()
    D1 = info["decoder_x.d1"]
    assert _rel(_pair(e, D1.in_, D1.in_lo, D1.in_elems, D1.in_dt).view(B, 256), z) < 2e-5
    if model == "lggmvae":
        YH, ZH, YB = info["encoder_x.h_top_dense"], info["encoder_x.z_mean"], info["encoder_x.y_block.0"]
        y = e.output("y").double().cpu()
        assert _rel(_pair(e, YH.in_, YH.in_lo, YH.in_elems, YH.in_dt).view(B, 32)[:, :30], y) < 2e-5
        e1out = _pair(e, YB.out, YB.out_lo, YB.out_elems, YB.out_dt).view(B, 1536)[:, 1024:]
        h_top = _pair(e, YH.out, None, YH.out_elems, YH.out_dt).view(B, 768)[:, :512]
        assert _rel(_pair(e, ZH.in_, ZH.in_lo, ZH.in_elems, ZH.in_dt).view(B, 512), e1out + h_top) < 2e-5


if __name__ == "__main__":      # debugging aid: per-layer table
    import sys
    model, H, B = (sys.argv[1], int(sys.argv[2]), int(sys.argv[3])) if len(sys.argv) > 3 else ("lgvae", 32, 4)
    try:
        check_layers(model, H, B, verbose=True)
    except AssertionError as ex:
        print("FAILED:", str(ex)[:300])
