import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__)))); sys.path.insert(0, 'tests')
import numpy as np, torch
from oracle import splitvae_oracle as O, bf16_emulation as E
from helpers import *
model, H, B, p, beta = sys.argv[1], int(sys.argv[2]), int(sys.argv[3]), int(sys.argv[4]), float(sys.argv[5])
params, batch = make_case(model, H, B, p)
u = batch["u"] if model=="lggmvae" else None
emu_sc, emu_g = E.forward_backward(params, model, batch["inputs"], batch["eps_g"], batch["eps_l"], u, beta=beta, alpha=40.0)
ref_sc, ref_g = O.forward_backward(params, model, batch["inputs"], batch["eps_g"], batch["eps_l"], u, beta=beta, alpha=40.0, dtype=torch.float64)
res = {}
for name, no_tc in (("ref", True), ("tc", False)):
    e = make_engine(model, H, B, "bf16", beta, no_tc=no_tc)
    e.load_params(params)
    sc, g = run_engine_step(e, batch, model, adam=False)
    res[name] = (sc, g)
    print(name, sc)
print("emu", emu_sc)
for k in emu_g:
    print(f"{k:36s} |g|={np.linalg.norm(emu_g[k]):.2e} ref-emu:{rel_l2(res['ref'][1][k], emu_g[k]):.1e} tc-emu:{rel_l2(res['tc'][1][k], emu_g[k]):.1e} tc-ref:{rel_l2(res['tc'][1][k], res['ref'][1][k]):.1e} emu-f64:{rel_l2(emu_g[k], ref_g[k]):.1e} tc-f64:{rel_l2(res['tc'][1][k], ref_g[k]):.1e}")
