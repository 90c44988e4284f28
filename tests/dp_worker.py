"""Worker of tests/test_gpu_dp.py, one process per GPU (launched by torch.distributed.run):

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port P tests/dp_worker.py MODEL H b PATCH BETA PRECISION

Every rank runs ONE graph-captured data-parallel train step (trainer.StepRunner: bucketed NCCL all-reduce inside the captured graph,
per-segment Adam on the optimizer stream) on its contiguous shard of a global batch of world*b images; rank 0 then checks the step
against the CPU oracle run on the WHOLE batch: loss scalars, every (all-reduced) gradient tensor, and the Adam-updated weights.
Prints one JSON line; exits non-zero on a mismatch."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import numpy as np
import torch
import torch.distributed as dist


def main():
    model, H, b, p, beta, prec = sys.argv[1], int(sys.argv[2]), int(sys.argv[3]), int(sys.argv[4]), float(sys.argv[5]), sys.argv[6]
    path = sys.argv[7] if len(sys.argv) > 7 else "nccl"            # "nvls": the fused multimem reduce + Adam + broadcast kernel
    from oracle import splitvae_oracle as O
    from splitvae_b200.engine import Engine
    from splitvae_b200.parallel import init_from_env
    from splitvae_b200.trainer import StepRunner
    rank, world, local = init_from_env()
    torch.cuda.set_device(local)
    lr = float(np.float32(1e-3))
    params = O.init_params(model, H, H)
    batch = O.synthetic_batch(world * b, H, p)
    sl = slice(rank * b, (rank + 1) * b)
    dev = lambda a: torch.from_numpy(np.ascontiguousarray(a)).cuda()
    nvls = None
    if path == "nvls":
        from splitvae_b200.parallel import NvlsArenas
        if not NvlsArenas.available():
            if rank == 0:
                print(json.dumps({"ok": True, "skipped": "no NVLS multicast support on this box", "world": world, "iterations": 1, "messages": []}), flush=True)
            dist.barrier()
            os._exit(0)
        nvls = NvlsArenas()
    e = Engine(model=model, height=H, width=H, batch=b, beta=beta, alpha=40.0, learning_rate=lr, world_size=world, precision=prec,
               arena_alloc=nvls.alloc if nvls else None)
    e.load_params(params)
    runner = StepRunner(e, use_graph=True, explicit_noise=True, nvls=nvls, write_reduced_grads=True)
    u = dev(batch["u"][sl]) if model != "lgvae" else None
    runner.step(dev(batch["inputs"][sl]), dev(batch["eps_g"][sl]), dev(batch["eps_l"][sl]), u)
    torch.cuda.synchronize()
    sc = runner.scalars()
    grads, new_params = e.get_grads(), e.get_params()
    # every rank must hold the same reduced gradients and the same updated weights (replicas stay in lock-step)
    gsum = torch.tensor([float(sum(np.abs(g).sum() for g in grads.values())), float(sum(np.abs(w).sum() for w in new_params.values()))],
                        dtype=torch.float64, device="cuda")
    lo, hi = gsum.clone(), gsum.clone()
    dist.all_reduce(lo, op=dist.ReduceOp.MIN)
    dist.all_reduce(hi, op=dist.ReduceOp.MAX)
    ok, msgs = True, []
    if not torch.equal(lo, hi):
        ok = False
        msgs.append(f"replicas diverged: {lo.tolist()} vs {hi.tolist()}")
    if rank == 0:
        un = batch["u"] if model != "lgvae" else None
        st = O.TrainState(params)
        ref_sc, ref_g = O.train_step(st, model, batch["inputs"], batch["eps_g"], batch["eps_l"], un, beta=beta, alpha=40.0, lr=lr, dtype=torch.float64)
        stol, gtol = (1e-5, 3e-4) if prec == "fp32" else (1e-3, 1.25e-2)
        for k, v in ref_sc.items():
            if abs(sc[k] - v) > stol * max(abs(v), 1.0 if prec == "fp32" else 1e-3):
                ok = False
                msgs.append(f"scalar {k}: {sc[k]} vs {v}")
        worst = ("", 0.0)
        for k, g in ref_g.items():
            n = float(np.linalg.norm(g))
            if n < 1e-7:
                continue
            r = float(np.linalg.norm(grads[k] - g)) / n
            if r > worst[1]:
                worst = (k, r)
        if worst[1] > gtol:
            ok = False
            msgs.append(f"gradient {worst[0]}: rel-L2 {worst[1]:.3e} > {gtol}")
        # Adam normalises the step: compare the parameter displacement
        wd = 0.0
        for k in params:
            dr = st.params[k] - params[k]
            if np.linalg.norm(dr) < 1e-9:
                continue
            wd = max(wd, float(np.linalg.norm((new_params[k] - params[k]) - dr) / np.linalg.norm(dr)))
        # (bf16x3: at step 1 Adam's update is -lr * sign(g) per element, so the ~0.5 % gradient error flips the sign of the elements
        #  with |g| below it: a few percent of them, each contributing 2 * lr)
        if wd > (0.02 if prec == "fp32" else 0.5):
            ok = False
            msgs.append(f"Adam displacement rel-L2 {wd:.3e}")
        print(json.dumps({"ok": ok, "path": path, "world": world, "model": model, "H": H, "per_gpu_batch": b, "precision": prec, "iterations": e.iterations,
                          "worst_gradient": worst, "adam_displacement_rel": wd, "scalars": sc, "messages": msgs}), flush=True)
    flag = torch.tensor([0 if ok else 1], device="cuda")
    dist.all_reduce(flag, op=dist.ReduceOp.MAX)
    runner.graph = None
    torch.cuda.synchronize()
    dist.barrier()
    sys.stdout.flush()
    os._exit(int(flag.item()))


if __name__ == "__main__":
    main()
