"""Storage-rounding model of libsplitvae's bf16 tensor-core path, on the CPU oracle.

TEST INFRASTRUCTURE ONLY (same rules as splitvae_oracle.py).

The bf16 path keeps fp32 accumulation, fp32 master weights, fp32 loss arithmetic and fp32
"head" outputs, but STORES activations, activation gradients and the tensor-core copies of the
weights in bfloat16.  Those roundings perturb gradients of the early layers by several percent
relative to an all-fp32 run (8 mantissa bits, compounding through ~10 stored tensors), which
says nothing about whether the kernels are right.  This module restates the reference model
(splitvae_oracle.py) with the SAME rounding points as the device, so the bf16 kernels can be
checked tightly (differences left: summation order and 1-ulp ties).

Rounding points (see DESIGN.md "bf16 data path"):
  forward : conv/dense weights; the first-layer image; every stored activation (after its
            activation function); upsampled tensors; z (decoder input); y; h = e1 + h_top.
            NOT rounded: biases, encoder heads (z_mean, z_sig), y_logits, prior heads, h_top,
            decoder outputs (mean, log_scale).
  backward: every stored activation gradient = gradient w.r.t. a layer's pre-activation, w.r.t.
            an upsampled tensor, w.r.t. z / y / h.  Activation derivatives are taken from the
            (rounded) stored OUTPUT of the activation, as on the device.
"""
from __future__ import annotations

import torch
import torch.nn.functional as F

from . import splitvae_oracle as O


def _r(t):
    return t.to(torch.bfloat16).to(t.dtype)


class _Q(torch.autograd.Function):
    """round the value, pass the gradient"""
    @staticmethod
    def forward(ctx, x):
        return _r(x)

    @staticmethod
    def backward(ctx, g):
        return g


class _G(torch.autograd.Function):
    """pass the value, round the gradient"""
    @staticmethod
    def forward(ctx, x):
        return x.view_as(x)

    @staticmethod
    def backward(ctx, g):
        return _r(g)


class _ActQ(torch.autograd.Function):
    """y = round(act(pre)) if stored in bf16 else act(pre); backward uses act'(y) from the stored output."""
    @staticmethod
    def forward(ctx, pre, kind, store_bf16):
        if kind == "relu":
            y = torch.relu(pre)
        elif kind == "elu":
            y = F.elu(pre)
        elif kind == "softplus":
            y = F.softplus(pre)
        else:
            y = pre.clone()
        if store_bf16:
            y = _r(y)
        ctx.kind = kind
        ctx.save_for_backward(y)
        return y

    @staticmethod
    def backward(ctx, g):
        (y,) = ctx.saved_tensors
        if ctx.kind == "relu":
            d = (y > 0).to(g.dtype)
        elif ctx.kind == "elu":
            d = torch.where(y > 0, torch.ones_like(y), y + 1)
        elif ctx.kind == "softplus":
            d = 1 - torch.exp(-y)
        else:
            d = torch.ones_like(y)
        return g * d, None, None


Q, G = _Q.apply, _G.apply


def _layer(x, P, names, stride, acts, store_bf16, conv=True):
    """One (possibly fused) device layer: pre = G(conv(x, Q(W)) + b); y = ActQ(pre)."""
    outs = []
    for name, act in zip(names, acts):
        w, b = Q(P[name + ".kernel"]), P[name + ".bias"]
        pre = O.conv2d_same(x, w, b, stride) if conv else O.dense(x, w, b)
        outs.append(_ActQ.apply(G(pre), act, store_bf16))
    return outs[0] if len(outs) == 1 else outs


def encoder_conv(P, pre, x, eps):
    h = _layer(Q(x), P, [pre + ".e1"], 2, ["relu"], True)
    h = _layer(h, P, [pre + ".e2"], 2, ["relu"], True)
    h = _layer(h, P, [pre + ".e3"], 2, ["relu"], True)
    h = h.reshape(h.shape[0], -1)
    z_mean, z_sig = _layer(h, P, [pre + ".e4_mean", pre + ".e4_sd"], 1, [None, "softplus"], False, conv=False)
    return z_mean + z_sig * eps, z_mean, z_sig


def encoder_gmvae(P, pre, x, eps, u, tau):
    h = Q(x)
    for i in range(3):
        h = _layer(h, P, [f"{pre}.h_block.{i}"], 2, ["elu"], True)
    h = h.reshape(h.shape[0], -1)
    yh1, e1out = _layer(h, P, [pre + ".y_block.0", pre + ".e1"], 1, ["elu", "elu"], True, conv=False)
    yh2 = _layer(yh1, P, [pre + ".y_block.2"], 1, ["elu"], True, conv=False)
    y_logits = _layer(yh2, P, [pre + ".y_dense"], 1, [None], False, conv=False)
    y = torch.softmax((y_logits - torch.log(-torch.log(u))) / tau, dim=1)
    yt = G(Q(y))
    h_top, zpm, zps = _layer(yt, P, [pre + ".h_top_dense", pre + ".z_prior_mean", pre + ".z_prior_sig"], 1,
                             ["elu", None, "softplus"], False, conv=False)
    hsum = G(Q(e1out + h_top))
    z_mean, z_sig = _layer(hsum, P, [pre + ".z_mean", pre + ".z_sig"], 1, [None, "softplus"], False, conv=False)
    return z_mean + z_sig * eps, z_mean, z_sig, y, y_logits, zpm, zps


def decoder(P, pre, zq, H, W):
    h = _layer(G(zq), P, [pre + ".d1"], 1, ["relu"], True, conv=False)
    h = h.reshape(-1, H // 8, W // 8, 128)
    h = _layer(h, P, [pre + ".d2"], 1, ["relu"], True)
    h = _layer(G(Q(O.resize2x(h))), P, [pre + ".d3"], 1, ["relu"], True)
    h = _layer(G(Q(O.resize2x(h))), P, [pre + ".d4"], 1, ["relu"], True)
    h = _layer(G(Q(O.resize2x(h))), P, [pre + ".d5"], 1, [None], False)
    return h[..., :3], h[..., 3:]


def model_forward(P, model, inputs, eps_g, eps_l, u=None, tau=0.4):
    H, W = inputs.shape[1], inputs.shape[2]
    x, x_hat = inputs[..., :3], inputs[..., 3:]
    out = {}
    if model == "lgvae":
        z_x, zm_x, zs_x = encoder_conv(P, "encoder_x", x, eps_g)
    else:
        z_x, zm_x, zs_x, y, y_logits, zpm, zps = encoder_gmvae(P, "encoder_x", x, eps_g, u, tau)
        out.update(y=y, y_logits=y_logits, z_prior_mean=zpm, z_prior_sig=zps)
    z_xh, zm_xh, zs_xh = encoder_conv(P, "encoder_x_hat", x_hat, eps_l)
    zcat = Q(torch.cat([z_x, z_xh], dim=1))
    x_mean, x_ls = decoder(P, "decoder_x", zcat, H, W)
    xh_mean, xh_ls = decoder(P, "decoder_x_hat", zcat[:, 128:], H, W)
    out.update(x_mean=x_mean, x_log_scale=x_ls, z_x=z_x, z_mean_x=zm_x, z_sig_x=zs_x, z_x_hat=z_xh,
               x_hat_mean=xh_mean, x_hat_log_scale=xh_ls, z_mean_x_hat=zm_xh, z_sig_x_hat=zs_xh)
    return out


def forward_backward(params, model, inputs, eps_g, eps_l, u=None, *, beta, alpha=40.0, tau=0.4, y_size=30):
    """Same contract as splitvae_oracle.forward_backward, fp32 arithmetic with the device's bf16 storage."""
    import numpy as np
    dtype = torch.float32
    P = O.to_torch(params, dtype)
    tin = torch.tensor(np.asarray(inputs), dtype=dtype)
    te_g = torch.tensor(np.asarray(eps_g), dtype=dtype)
    te_l = torch.tensor(np.asarray(eps_l), dtype=dtype)
    tu = None if u is None else torch.tensor(np.asarray(u), dtype=dtype)
    out = model_forward(P, model, tin, te_g, te_l, tu, tau)
    L = O.step_losses(out, tin, model, beta, alpha, y_size)
    L["total"].backward()
    scalars = {k: float(v.detach()) for k, v in L.items()}
    grads = {k: t.grad.detach().numpy().copy() for k, t in P.items()}
    return scalars, grads
