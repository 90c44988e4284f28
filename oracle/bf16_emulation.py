"""Storage-rounding models of libsplitvae's two tensor-core precision modes, on the CPU oracle.

TEST INFRASTRUCTURE ONLY (same rules as splitvae_oracle.py).

Both modes keep fp32 accumulation, fp32 master weights, fp32 loss arithmetic and fp32 "head" outputs; they differ
in the operands of the FORWARD tensor-core products:

  mode "bf16"    every tensor-core operand (activations, activation gradients, weight copies) is ONE bfloat16.
                 Forward roundings of 2^-9 flip the sign of ~0.5 % of the near-zero ReLU pre-activations per layer;
                 a flipped unit changes its whole gradient contribution, so gradients of the early layers sit 6-11 %
                 (rel-L2) from an fp32 run although every kernel is right.  Fast mode, not the parity mode.
  mode "bf16x3"  (default on the device) forward operands are bfloat16 PAIRS hi + lo (hi = bf16(v), lo = bf16(v - hi):
                 16-17 significant bits) and every forward product is three MMAs hi*hi + lo*hi + hi*lo with fp32
                 accumulation; activations that feed another forward layer are stored as such pairs.  The LAST
                 decoder layer d5 (no ReLU after it) and the whole BACKWARD pass (dgrad, wgrad, stored gradients)
                 use single-bf16 operands: their error is smooth (~0.5 %), not mask flips.  Gradients then sit
                 < 1e-2 from fp64 (tests/test_gpu_parity.py holds exactly that).

This module restates the reference model (splitvae_oracle.py) with the SAME rounding points as the device, so the
kernels can be checked tightly (differences left: summation order and 1-ulp ties).

Rounding points:
  forward : conv/dense weights; the first-layer image; every stored activation (after its activation function);
            upsampled tensors; z (decoder input); y; h = e1 + h_top.   "bf16": one bf16.  "bf16x3": a bf16 pair,
            except d5's input / weights (one bf16).
            NOT rounded: biases, encoder heads (z_mean, z_sig), y_logits, prior heads, h_top, decoder outputs.
  backward: every stored activation gradient (one bf16) = gradient w.r.t. a layer's pre-activation, w.r.t. an
            upsampled tensor, w.r.t. z / y / h; dgrad multiplies it with bf16(W), wgrad with bf16(X) - in both
            modes.  Activation derivatives are taken from the stored OUTPUT of the activation, as on the device.
"""
from __future__ import annotations

import numpy as np
import torch
import torch.nn.functional as F

from . import splitvae_oracle as O


def _r(t):
    return t.to(torch.bfloat16).to(t.dtype)


def _r2(t):
    """bf16 pair hi + lo (what the bf16x3 mode stores / multiplies)."""
    hi = _r(t)
    return hi + _r(t - hi)


_FWD = {"bf16": _r, "bf16x3": _r2}


class _Lin(torch.autograd.Function):
    """pre = op(fr(x), fr(w)) + b in the forward; the backward multiplies single-bf16 operands: dX = r(g) * r(W),
    dW = r(X) * r(g), db = sum r(g) - on the device g is STORED in bf16 (the caller rounds it, see _G)."""

    @staticmethod
    def forward(ctx, x, w, b, stride, conv, fr):
        ctx.save_for_backward(x, w)
        ctx.stride, ctx.conv = stride, conv
        xx, ww = fr(x), fr(w)
        return O.conv2d_same(xx, ww, b, stride) if conv else O.dense(xx, ww, b)

    @staticmethod
    def backward(ctx, g):
        x, w = ctx.saved_tensors
        with torch.enable_grad():
            xr = _r(x).detach().requires_grad_()
            wr = _r(w).detach().requires_grad_()
            y = O.conv2d_same(xr, wr, None, ctx.stride) if ctx.conv else xr @ wr
            gx, gw = torch.autograd.grad(y, (xr, wr), g)
        gb = g.reshape(-1, g.shape[-1]).sum(0)
        return gx, gw, gb, None, None, None


class _G(torch.autograd.Function):
    """pass the value, round the gradient to one bf16 (a stored activation gradient)"""
    @staticmethod
    def forward(ctx, x):
        return x.view_as(x)

    @staticmethod
    def backward(ctx, g):
        return _r(g)


class _Q(torch.autograd.Function):
    """round the value with `fr`, pass the gradient"""
    @staticmethod
    def forward(ctx, x, fr):
        return fr(x)

    @staticmethod
    def backward(ctx, g):
        return g, None


class _ActQ(torch.autograd.Function):
    """y = fr(act(pre)) if stored in bf16 (pairs) else act(pre); the backward uses act'(hi(y)) from the stored output."""
    @staticmethod
    def forward(ctx, pre, kind, fr):
        if kind == "relu":
            y = torch.relu(pre)
        elif kind == "elu":
            y = F.elu(pre)
        elif kind == "softplus":
            y = F.softplus(pre)
        else:
            y = pre.clone()
        if fr is not None:
            y = fr(y)
        ctx.kind = kind
        ctx.stored = fr is not None
        ctx.save_for_backward(y)
        return y

    @staticmethod
    def backward(ctx, g):
        (y,) = ctx.saved_tensors
        if ctx.stored:
            y = _r(y)            # the backward kernels read the hi plane
        if ctx.kind == "relu":
            d = (y > 0).to(g.dtype)
        elif ctx.kind == "elu":
            d = torch.where(y > 0, torch.ones_like(y), y + 1)
        elif ctx.kind == "softplus":
            d = 1 - torch.exp(-y)
        else:
            d = torch.ones_like(y)
        return g * d, None, None


def _layer(x, P, names, stride, acts, fr, store, conv=True):
    """One (possibly fused) device layer: pre = G(op(fr(x), fr(W)) + b); y = ActQ(pre)."""
    outs = []
    for name, act in zip(names, acts):
        pre = _Lin.apply(x, P[name + ".kernel"], P[name + ".bias"], stride, conv, fr)
        outs.append(_ActQ.apply(_G.apply(pre), act, fr if store else None))
    return outs[0] if len(outs) == 1 else outs


def encoder_conv(P, pre, x, eps, fr):
    h = _layer(x, P, [pre + ".e1"], 2, ["relu"], fr, True)
    h = _layer(h, P, [pre + ".e2"], 2, ["relu"], fr, True)
    h = _layer(h, P, [pre + ".e3"], 2, ["relu"], fr, True)
    h = h.reshape(h.shape[0], -1)
    z_mean, z_sig = _layer(h, P, [pre + ".e4_mean", pre + ".e4_sd"], 1, [None, "softplus"], fr, False, conv=False)
    return z_mean + z_sig * eps, z_mean, z_sig


def encoder_gmvae(P, pre, x, eps, u, tau, fr):
    h = x
    for i in range(3):
        h = _layer(h, P, [f"{pre}.h_block.{i}"], 2, ["elu"], fr, True)
    h = h.reshape(h.shape[0], -1)
    yh1, e1out = _layer(h, P, [pre + ".y_block.0", pre + ".e1"], 1, ["elu", "elu"], fr, True, conv=False)
    yh2 = _layer(yh1, P, [pre + ".y_block.2"], 1, ["elu"], fr, True, conv=False)
    y_logits = _layer(yh2, P, [pre + ".y_dense"], 1, [None], fr, False, conv=False)
    y = torch.softmax((y_logits - torch.log(-torch.log(u))) / tau, dim=1)
    yt = _G.apply(_Q.apply(y, fr))
    h_top, zpm, zps = _layer(yt, P, [pre + ".h_top_dense", pre + ".z_prior_mean", pre + ".z_prior_sig"], 1,
                             ["elu", None, "softplus"], fr, False, conv=False)
    hsum = _G.apply(_Q.apply(e1out + h_top, fr))
    z_mean, z_sig = _layer(hsum, P, [pre + ".z_mean", pre + ".z_sig"], 1, [None, "softplus"], fr, False, conv=False)
    return z_mean + z_sig * eps, z_mean, z_sig, y, y_logits, zpm, zps


def decoder(P, pre, zq, H, W, fr):
    h = _layer(_G.apply(zq), P, [pre + ".d1"], 1, ["relu"], fr, True, conv=False)
    h = h.reshape(-1, H // 8, W // 8, 128)
    h = _layer(h, P, [pre + ".d2"], 1, ["relu"], fr, True)
    h = _layer(_G.apply(_Q.apply(O.resize2x(h), fr)), P, [pre + ".d3"], 1, ["relu"], fr, True)
    h = _layer(_G.apply(_Q.apply(O.resize2x(h), fr)), P, [pre + ".d4"], 1, ["relu"], fr, True)
    # d5: linear outputs, nothing discontinuous downstream -> single-bf16 operands in both modes
    h = _layer(_G.apply(_Q.apply(O.resize2x(h), _r)), P, [pre + ".d5"], 1, [None], _r, False)
    return h[..., :3], h[..., 3:]


def model_forward(P, model, inputs, eps_g, eps_l, u=None, tau=0.4, mode="bf16"):
    fr = _FWD[mode]
    H, W = inputs.shape[1], inputs.shape[2]
    x, x_hat = inputs[..., :3], inputs[..., 3:]
    out = {}
    if model == "gmvae":
        z_x, zm_x, zs_x, y, y_logits, zpm, zps = encoder_gmvae(P, "encoder_x", x, eps_g, u, tau, fr)
        x_mean, x_ls = decoder(P, "decoder_x", _Q.apply(z_x, fr), H, W, fr)
        return dict(x_mean=x_mean, x_log_scale=x_ls, z_x=z_x, z_mean_x=zm_x, z_sig_x=zs_x, y=y, y_logits=y_logits,
                    z_prior_mean=zpm, z_prior_sig=zps)
    if model == "lgvae":
        z_x, zm_x, zs_x = encoder_conv(P, "encoder_x", x, eps_g, fr)
    else:
        z_x, zm_x, zs_x, y, y_logits, zpm, zps = encoder_gmvae(P, "encoder_x", x, eps_g, u, tau, fr)
        out.update(y=y, y_logits=y_logits, z_prior_mean=zpm, z_prior_sig=zps)
    z_xh, zm_xh, zs_xh = encoder_conv(P, "encoder_x_hat", x_hat, eps_l, fr)
    zcat = _Q.apply(torch.cat([z_x, z_xh], dim=1), fr)
    x_mean, x_ls = decoder(P, "decoder_x", zcat, H, W, fr)
    xh_mean, xh_ls = decoder(P, "decoder_x_hat", zcat[:, 128:], H, W, fr)
    out.update(x_mean=x_mean, x_log_scale=x_ls, z_x=z_x, z_mean_x=zm_x, z_sig_x=zs_x, z_x_hat=z_xh,
               x_hat_mean=xh_mean, x_hat_log_scale=xh_ls, z_mean_x_hat=zm_xh, z_sig_x_hat=zs_xh)
    return out


def forward_backward(params, model, inputs, eps_g, eps_l, u=None, *, beta, alpha=40.0, tau=0.4, y_size=30, mode="bf16"):
    """Same contract as splitvae_oracle.forward_backward, fp32 arithmetic with the device's storage roundings."""
    dtype = torch.float32
    P = O.to_torch(params, dtype)
    tin = torch.tensor(np.asarray(inputs), dtype=dtype)
    te_g = torch.tensor(np.asarray(eps_g), dtype=dtype)
    te_l = torch.tensor(np.asarray(eps_l), dtype=dtype)
    tu = None if u is None else torch.tensor(np.asarray(u), dtype=dtype)
    out = model_forward(P, model, tin, te_g, te_l, tu, tau, mode)
    L = O.step_losses(out, tin, model, beta, alpha, y_size)
    L["total"].backward()
    scalars = {k: float(v.detach()) for k, v in L.items()}
    grads = {k: t.grad.detach().numpy().copy() for k, t in P.items()}
    return scalars, grads
