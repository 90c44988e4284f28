"""Shared helpers of the GPU parity tests: run the oracle and the CUDA engine on identical
weights, inputs and noise and compare scalars / gradients / updated parameters."""
import numpy as np
import torch

from oracle import splitvae_oracle as O


def rel_l2(a, b):
    a = np.asarray(a, np.float64).ravel()
    b = np.asarray(b, np.float64).ravel()
    d = np.linalg.norm(b)
    return float(np.linalg.norm(a - b) / d) if d > 0 else float(np.linalg.norm(a))


def make_case(model, H, B, p, seed_base=0, y_size=30, ls_bias=None):
    params = O.init_params(model, H, H, seed=5 + seed_base, y_size=y_size, decoder_ls_bias=ls_bias)
    batch = O.synthetic_batch(B, H, p, y_size=y_size, seed_base=seed_base)
    return params, batch


def make_engine(model, H, B, precision, beta, alpha=40.0, tau=0.4, lr=1e-4, y_size=30, no_tc=False, world_size=1):
    from splitvae_b200.engine import Engine
    return Engine(model=model, height=H, width=H, batch=B, y_size=y_size, tau=tau, beta=beta, alpha=alpha,
                  learning_rate=lr, precision=precision, no_tc=no_tc, world_size=world_size)


def to_dev(a):
    return torch.from_numpy(np.ascontiguousarray(a)).cuda()


def run_engine_step(e, batch, model, adam=True):
    x = to_dev(batch["inputs"])
    eg, el = to_dev(batch["eps_g"]), to_dev(batch["eps_l"])
    u = to_dev(batch["u"]) if model != "lgvae" else None
    e.forward(x, eg, el, u)
    e.loss_fwd_bwd(x)
    for s in range(len(e.segments)):
        e.backward_segment(s)
    torch.cuda.synchronize()
    sc = e.scalars()
    grads = e.get_grads()
    if adam:
        e.adam_step()
        torch.cuda.synchronize()
    return sc, grads


def compare_grads(grads, ref, tol, floor=1e-7):
    bad = []
    worst = 0.0
    for k in ref:
        r = rel_l2(grads[k], ref[k])
        if np.linalg.norm(ref[k]) < floor:
            continue
        worst = max(worst, r)
        if not (r <= tol):
            bad.append((k, r))
    return worst, bad
