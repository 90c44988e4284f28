"""Generates tests/golden/*.json from the CPU oracle (fp64 for scalars/gradients, fp32 numpy Adam).

The reference (TF 2.0) cannot run in this image and ships no fixtures, so these vectors
pin the ORACLE against regressions and give the GPU tests a committed target; they are not outputs of the
TF reference itself ("parity unpinned", see oracle/splitvae_oracle.py).  Inputs, noise and weights are
regenerated from seeds by oracle.synthetic_batch / oracle.init_params, so only results are stored.

    python scripts/make_golden.py
"""
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import splitvae_oracle as O

CASES = [  # name, model, H, B, patch, beta, alpha, seed_base, steps
    ("c1_lgvae_h32", "lgvae", 32, 4, 1, 1.0, 40.0, 0, 3),
    ("c2_lgvae_h64", "lgvae", 64, 2, 8, 120.0, 40.0, 0, 2),
    ("c3_lggmvae_h32", "lggmvae", 32, 4, 4, 40.0, 40.0, 0, 3),
    ("c4_lggmvae_h64", "lggmvae", 64, 2, 8, 120.0, 40.0, 0, 2),
    ("tiny_lgvae_h16", "lgvae", 16, 3, 2, 7.0, 40.0, 40, 3),
    ("tiny_lggmvae_h16", "lggmvae", 16, 3, 4, 7.0, 3.0, 40, 3),
]


def summarise(a):
    a = np.asarray(a, np.float64).ravel()
    return {"l2": float(np.linalg.norm(a)), "sum": float(a.sum()), "head": [float(v) for v in a[:4]]}


def make(name, model, H, B, p, beta, alpha, seed_base, steps):
    params = O.init_params(model, H, H, seed=5 + seed_base)
    batch = O.synthetic_batch(B, H, p, seed_base=seed_base)
    u = batch["u"] if model == "lggmvae" else None
    sc64, g64 = O.forward_backward(params, model, batch["inputs"], batch["eps_g"], batch["eps_l"], u, beta=beta, alpha=alpha,
                                   dtype=torch.float64)
    out = {"case": dict(model=model, H=H, B=B, patch=p, beta=beta, alpha=alpha, seed_base=seed_base, steps=steps, lr=1e-4),
           "scalars_fp64": sc64, "grads_fp64": {k: summarise(v) for k, v in g64.items()},
           "inputs_sum": float(np.asarray(batch["inputs"], np.float64).sum()),
           "params_sum": float(sum(np.asarray(v, np.float64).sum() for v in params.values()))}
    st = O.TrainState(params)
    traj = []
    for _ in range(steps):
        sc, _ = O.train_step(st, model, batch["inputs"], batch["eps_g"], batch["eps_l"], u, beta=beta, alpha=alpha,
                             lr=float(np.float32(1e-4)))
        traj.append(sc)
    out["train_scalars_fp32"] = traj
    out["params_after"] = {k: summarise(v - params[k]) for k, v in st.params.items()}
    return out


if __name__ == "__main__":
    torch.set_num_threads(os.cpu_count() or 1)
    os.makedirs(os.path.join(ROOT, "tests", "golden"), exist_ok=True)
    for c in CASES:
        path = os.path.join(ROOT, "tests", "golden", c[0] + ".json")
        with open(path, "w") as f:
            json.dump(make(*c), f, indent=1)
        print("wrote", path)
