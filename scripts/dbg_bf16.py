import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__)))); sys.path.insert(0, 'tests')
import numpy as np, torch
from oracle import splitvae_oracle as O
from helpers import *
model, H, B, p, beta = sys.argv[1], int(sys.argv[2]), int(sys.argv[3]), 4, float(sys.argv[4])
params, batch = make_case(model, H, B, p)
ref_sc, ref_g = O.forward_backward(params, model, batch["inputs"], batch["eps_g"], batch["eps_l"], batch["u"] if model=="lggmvae" else None, beta=beta, alpha=40.0, dtype=torch.float64)
res = {}
for prec in ("fp32", "bf16"):
    e = make_engine(model, H, B, prec, beta)
    e.load_params(params)
    sc, g = run_engine_step(e, batch, model, adam=False)
    res[prec] = (sc, g)
    print(prec, sc)
print("ref", ref_sc)
for k in ref_g:
    print(f"{k:40s} |g|={np.linalg.norm(ref_g[k]):.3e} fp32:{rel_l2(res['fp32'][1][k], ref_g[k]):.2e} bf16:{rel_l2(res['bf16'][1][k], ref_g[k]):.2e}")
from oracle import bf16_emulation as E
emu_sc, emu_g = E.forward_backward(params, model, batch["inputs"], batch["eps_g"], batch["eps_l"], batch["u"] if model=="lggmvae" else None, beta=beta, alpha=40.0)
print("emu", emu_sc)
for k in ref_g:
    print(f"{k:40s} bf16-vs-emu:{rel_l2(res['bf16'][1][k], emu_g[k]):.2e}  emu-vs-fp64:{rel_l2(emu_g[k], ref_g[k]):.2e}")
