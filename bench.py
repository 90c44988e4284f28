#!/usr/bin/env python
"""Benchmark of the SPLIT-VAE train step (BASELINE.json metric: train images/sec, CelebA64 shape).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference] [--workload c2|c1|c3|c4]

One "step" = one train_step_lg_vae (vae/trainer.py:120-144): forward, fused loss fwd+bwd, backward,
Keras Adam, captured in a CUDA graph.  Prints ONE JSON line (rank 0).  See DESIGN.md "Measurement".
"""
from __future__ import annotations

import argparse
import json
import os
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

WORKLOADS = {
    # name: (model, H, per-GPU batch, patch, beta, alpha, description)
    "c1": ("lgvae", 32, 64, 1, 1.0, 40.0, "SPLIT-VAE SVHN-shape 32x32x3 --beta 1 --patch_size 1 batch 64"),
    "c2": ("lgvae", 64, 256, 8, 120.0, 40.0, "SPLIT-VAE CelebA64-shape 64x64x3 --beta 120 --patch_size 8 -no_label batch 256/GPU"),
    "c3": ("lggmvae", 32, 256, 4, 40.0, 40.0, "SPLIT-GMVAE SVHN-shape --beta 40 --alpha 40 --y_size 30 --patch_size 4 batch 256/GPU"),
    "c4": ("lggmvae", 64, 256, 8, 120.0, 40.0, "SPLIT-GMVAE CelebA64-shape --beta 120 --alpha 40 --y_size 30 --patch_size 8 batch 256/GPU"),
}
TRAIN_GFLOP_PER_IMAGE = {"c1": 0.5623, "c2": 2.2492, "c3": 0.8011, "c4": 3.1994}
# operand type per layer class (BASELINE.md section 3 asks for it); accumulation is fp32 (TMEM) everywhere
OPERANDS = {
    "bf16x3": {"forward conv/dense except d5": "bf16 pairs hi+lo, 3 tcgen05 MMAs (hi*hi + lo*hi + hi*lo)", "forward d5": "bf16",
               "dgrad": "bf16", "wgrad": "bf16", "stored activations": "bf16 pairs (d5 input: bf16)", "stored activation gradients": "bf16",
               "accumulate": "fp32", "loss / KL / reparameterisation": "fp32", "master weights / Adam": "fp32",
               "parity": "every scalar rel 1e-3, every gradient tensor rel-L2 1e-2 vs the fp64 oracle (tests/test_gpu_parity.py)"},
    "bf16": {"forward": "bf16", "dgrad": "bf16", "wgrad": "bf16", "stored activations / gradients": "bf16", "accumulate": "fp32",
             "loss / KL / reparameterisation": "fp32", "master weights / Adam": "fp32",
             "parity": "ELBO rel 1e-3; gradients only 5-10 % (rel-L2) from fp64: 2^-9 forward roundings flip ReLU masks of near-zero units"},
    "fp32": {"all": "fp32 SIMT reference kernels"},
}
LOSS_BYTES_PER_IMAGE = {32: 122880, 64: 491520}  # fused loss fwd+bwd, fp32 in/out


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as f:
            d = json.load(f)
        return dict(hbm=d["hbm_gbs"], tensor_burst=d["bf16_tflops"], tensor_sustained=d["bf16_tflops_sustained"], source="measured")
    return dict(hbm=6650.0, tensor_burst=1590.0, tensor_sustained=1400.0, source="fallback")


class ClockSampler(threading.Thread):
    """Samples SM clock and throttle reasons through NVML during the timed region."""

    def __init__(self, index=0, period=0.1):
        super().__init__(daemon=True)
        self.period, self.index = period, index
        self.samples, self.reasons, self.max_mhz = [], set(), None
        self._halt = threading.Event()
        self.ok = False
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.dev = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.dev, pynvml.NVML_CLOCK_SM)
            self.ok = True
        except Exception:
            self.ok = False

    def run(self):
        if not self.ok:
            return
        nv = self.nv
        names = {nv.nvmlClocksThrottleReasonHwSlowdown: "hw_slowdown",
                 nv.nvmlClocksThrottleReasonHwThermalSlowdown: "hw_thermal_slowdown",
                 nv.nvmlClocksThrottleReasonSwThermalSlowdown: "sw_thermal_slowdown",
                 nv.nvmlClocksThrottleReasonSwPowerCap: "sw_power_cap"}
        while not self._halt.is_set():
            try:
                self.samples.append(nv.nvmlDeviceGetClockInfo(self.dev, nv.NVML_CLOCK_SM))
                r = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.dev)
                for bit, name in names.items():
                    if r & bit:
                        self.reasons.add(name)
            except Exception:
                pass
            time.sleep(self.period)

    def stop(self):
        self._halt.set()
        if self.ok:
            self.join(timeout=2)
        s = sorted(self.samples)
        return {"sm_mhz": s[len(s) // 2] if s else None, "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons)}


def cpu_baseline(model, H, batch, patch, beta, alpha, steps, threads=None, min_seconds=0.0):
    """The oracle (CPU restatement of the TF reference; TF itself is not installable) timed on the host
    cores over a bounded sample: `steps` train steps of `batch` images of this workload's shape (more steps
    until `min_seconds` of CPU work have been timed)."""
    import torch
    from oracle import splitvae_oracle as O
    threads = threads or os.cpu_count() or 1
    torch.set_num_threads(threads)
    params = O.init_params(model, H, H)
    st = O.TrainState(params)
    b = O.synthetic_batch(batch, H, patch)
    u = b["u"] if model == "lggmvae" else None
    O.train_step(st, model, b["inputs"], b["eps_g"], b["eps_l"], u, beta=beta, alpha=alpha)  # warm-up
    times = []
    while len(times) < steps or sum(times) < min_seconds:
        t0 = time.perf_counter()
        O.train_step(st, model, b["inputs"], b["eps_g"], b["eps_l"], u, beta=beta, alpha=alpha)
        times.append(time.perf_counter() - t0)
    return batch * len(times) / sum(times), threads, times


def run_reference(args, wl):
    """--impl reference: the reference's CPU implementation of the path (oracle port; TF 2.0 cannot be installed)."""
    model, H, B, patch, beta, alpha, desc = WORKLOADS[wl]
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    sample_batch = B          # the workload's own per-GPU batch: same config as the GPU arm
    ips, threads, times = cpu_baseline(model, H, sample_batch, patch, beta, alpha, max(1, args.steps))
    ms = 1000.0 * sum(times) / len(times)
    line = {"impl": "reference", "metric": "train images/sec", "value": ips, "unit": "images/s", "n_gpus": args.gpus,
            "steps": args.steps, "warmup": 1, "ms_per_step": ms, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": desc, "model": model, "global_batch": sample_batch, "parallelism": "cpu",
                       "per_step_sample": f"one CPU step = the workload's batch of {sample_batch} images"},
            "cpu_baseline": {"value": ips, "unit": "images/s", "cores": threads, "kind": "port",
                             "sample": f"{args.steps} oracle train steps of {sample_batch} images ({H}x{H}x3), torch CPU fp32"},
            "e2e": {"value": ips, "unit": "images/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    emit(line)


_REAL_STDOUT = None


def _stdout_json_only():
    """The driver reads ONE JSON line from stdout: send everything else that lands on fd 1 (NCCL's 'NCCL version ...' banner, library
    chatter) to stderr and keep a private duplicate of the real stdout for the result line."""
    global _REAL_STDOUT
    if _REAL_STDOUT is None:
        sys.stdout.flush()
        _REAL_STDOUT = os.dup(1)
        os.dup2(2, 1)


def emit(line):
    text = json.dumps(line) + "\n"
    if _REAL_STDOUT is None:
        sys.stdout.write(text)
        sys.stdout.flush()
    else:
        sys.stdout.flush()
        os.write(_REAL_STDOUT, text.encode())


def main():
    _stdout_json_only()
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=30)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", type=str, default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", type=str, default="c2", choices=sorted(WORKLOADS))
    ap.add_argument("--precision", type=str, default="bf16x3", choices=["bf16x3", "bf16", "fp32"],
                    help="bf16x3 (default, the parity mode): forward on bf16 pairs, backward single bf16; bf16: single-bf16 operands everywhere")
    ap.add_argument("--no-fast-mode", action="store_true", help="skip the secondary single-bf16 throughput measurement")
    ap.add_argument("--no-other-workloads", action="store_true", help="skip the short device-resident lines of the other BASELINE.json configs")
    ap.add_argument("--no-graph", action="store_true")
    ap.add_argument("--nccl", action="store_true", help="data parallel through the bucketed NCCL all-reduce even when NVLS multicast is available")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--hang-dump", type=int, default=0, help="dump all Python stacks after this many seconds (debugging)")
    args = ap.parse_args()
    if args.hang_dump:
        import faulthandler
        faulthandler.dump_traceback_later(args.hang_dump, exit=True)
    wl = args.workload
    if args.impl == "reference":
        return run_reference(args, wl)

    import torch
    import torch.distributed as dist
    from splitvae_b200.augmentation import Augmentator
    from splitvae_b200.engine import Engine
    from splitvae_b200.parallel import init_from_env
    from splitvae_b200.trainer import StepRunner

    model, H, B, patch, beta, alpha, desc = WORKLOADS[wl]
    rank, world, local_rank = init_from_env()
    if world != args.gpus and world > 1:
        raise SystemExit(f"--gpus {args.gpus} but WORLD_SIZE={world}")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    K, Wm = args.steps, max(3, args.warmup)

    nvls = None
    if world > 1 and not args.nccl:
        from splitvae_b200.parallel import NvlsArenas
        if NvlsArenas.available():
            nvls = NvlsArenas()
    e = Engine(model=model, height=H, width=H, batch=B, beta=beta, alpha=alpha, learning_rate=1e-4, world_size=world,
               precision=args.precision, rng_stream=rank, arena_alloc=nvls.alloc if nvls else None)
    e.init_params(seed=5)  # same seed on every rank: replicated weights
    runner = StepRunner(e, use_graph=not args.no_graph, nvls=nvls)
    aug = Augmentator("scramble", patch)

    # synthetic data: a pool of pinned uint8 batches (distinct per rank) + per-image patch permutations
    g = torch.Generator().manual_seed(1000 + rank)
    pool = 4
    host_u8 = [torch.randint(0, 256, (B, H, H, 3), dtype=torch.uint8, generator=g).pin_memory() for _ in range(pool)]
    n_patch = (H // patch) ** 2
    host_perm = [torch.stack([torch.randperm(n_patch, generator=g) for _ in range(B)]).to(torch.int32).pin_memory() for _ in range(pool)]
    dev_u8 = torch.empty(B, H, H, 3, dtype=torch.uint8, device=dev)
    dev_perm = torch.empty(B, n_patch, dtype=torch.int32, device=dev)
    host_scalars = torch.empty(8, dtype=torch.float32).pin_memory()
    l2_flush = torch.empty(192 * 1024 * 1024, dtype=torch.uint8, device=dev)  # > 126 MB L2

    def stage(i):
        dev_u8.copy_(host_u8[i % pool], non_blocking=True)
        dev_perm.copy_(host_perm[i % pool], non_blocking=True)
        aug.scramble(dev_u8, dev_perm, out=runner.inputs)

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- launches per step (eager, un-captured) + warm-up ------------------------------------
    stage(0)
    l0 = e.launch_count
    runner._issue()
    torch.cuda.synchronize()
    launches_per_step = e.launch_count - l0
    if not args.no_graph:
        runner.capture(warmup=1)
    for i in range(Wm):
        runner.step()
    barrier()

    # ---- (A) device-resident throughput: K graph replays, inputs already in HBM -----------------
    sampler = ClockSampler(local_rank)
    sampler.start()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    ev0.record()
    for i in range(K):
        runner.step()
    ev1.record()
    barrier()
    t_dev = ev0.elapsed_time(ev1) / 1000.0

    # ---- (B) end to end: pinned host uint8 -> H2D -> scramble -> step -> D2H scalars, every step.  The input pipeline is
    # double-buffered: a copy stream uploads and scrambles batch i+1 into a staging tensor while the graph of step i runs;
    # the main stream then copies the staged batch into the graph's input buffer (device to device) and replays.
    main = torch.cuda.current_stream()
    copy_stream = torch.cuda.Stream()
    staging = [torch.empty_like(runner.inputs) for _ in range(2)]
    dev_u8s = [torch.empty_like(dev_u8) for _ in range(2)]
    dev_perms = [torch.empty_like(dev_perm) for _ in range(2)]
    ev_ready = [torch.cuda.Event() for _ in range(2)]
    ev_consumed = [torch.cuda.Event() for _ in range(2)]

    def stage_async(i):
        b = i % 2
        with torch.cuda.stream(copy_stream):
            copy_stream.wait_event(ev_consumed[b])
            dev_u8s[b].copy_(host_u8[i % pool], non_blocking=True)
            dev_perms[b].copy_(host_perm[i % pool], non_blocking=True)
            aug.scramble(dev_u8s[b], dev_perms[b], out=staging[b])
            ev_ready[b].record(copy_stream)

    def e2e_steps(n):
        stage_async(0)
        for i in range(n):
            stage_async(i + 1)
            main.wait_event(ev_ready[i % 2])
            runner.inputs.copy_(staging[i % 2], non_blocking=True)
            ev_consumed[i % 2].record(main)
            runner.step()
            host_scalars.copy_(e.output("scalars"), non_blocking=True)
        main.wait_event(ev_ready[n % 2])    # the one batch staged ahead of the loop's end

    for b in range(2):
        ev_consumed[b].record(main)
    e2e_steps(2)
    barrier()
    ev2, ev3 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev2.record()
    e2e_steps(K)
    ev3.record()
    barrier()
    t_e2e = ev2.elapsed_time(ev3) / 1000.0
    clocks = sampler.stop()
    final_total = float(host_scalars[5])

    if world > 1:
        t = torch.tensor([t_dev, t_e2e], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        t_dev, t_e2e = t.tolist()

    peaks_for_others = measured_peaks()
    # ---- secondary number: the single-bf16 fast mode (same workload, device-resident graph replays); NOT the headline: its gradients
    # are 5-10 % from fp32
    fast = None
    if args.precision == "bf16x3" and not args.no_fast_mode and not args.no_graph:
        ef = Engine(model=model, height=H, width=H, batch=B, beta=beta, alpha=alpha, learning_rate=1e-4, world_size=world,
                    precision="bf16", rng_stream=rank)
        ef.init_params(seed=5)
        rf = StepRunner(ef, use_graph=True)
        rf.inputs.copy_(runner.inputs)
        rf.capture(warmup=1)
        for i in range(Wm):
            rf.step()
        barrier()
        f0, f1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        f0.record()
        for i in range(K):
            rf.step()
        f1.record()
        barrier()
        t_fast = f0.elapsed_time(f1) / 1000.0
        if world > 1:
            tf_ = torch.tensor([t_fast], device=dev, dtype=torch.float64)
            dist.all_reduce(tf_, op=dist.ReduceOp.MAX)
            t_fast = float(tf_[0])
        fast = {"precision": "bf16", "value": world * B * K / t_fast, "unit": "images/s", "ms_per_step": 1000.0 * t_fast / K,
                "note": "single-bf16 operands everywhere (round-1 default): faster, but gradients 5-10 % rel-L2 from fp32; not the headline"}
        rf.graph = None
        del rf, ef

    # ---- the other BASELINE.json configurations, short device-resident runs (same precision, same data-parallel world): C1 is the
    # reference's own CPU-runnable case, C3 / C4 the SPLIT-GMVAE shapes (C4 is quoted on 8 GPUs: it rides along in every --gpus N run)
    others = None
    if not args.no_other_workloads and not args.no_graph and wl == "c2":
        others = {}
        for ow in ("c1", "c3", "c4"):
            om, oH, oB, op_, ob, oa, odesc = WORKLOADS[ow]
            nvo = None
            if nvls is not None:
                from splitvae_b200.parallel import NvlsArenas
                nvo = NvlsArenas()
            eo = Engine(model=om, height=oH, width=oH, batch=oB, beta=ob, alpha=oa, learning_rate=1e-4, world_size=world,
                        precision=args.precision, rng_stream=rank, arena_alloc=nvo.alloc if nvo else None)
            eo.init_params(seed=5)
            ro = StepRunner(eo, use_graph=True, nvls=nvo)
            go = torch.Generator().manual_seed(2000 + rank)
            u8o = torch.randint(0, 256, (oB, oH, oH, 3), dtype=torch.uint8, generator=go).to(dev)
            Augmentator("scramble", op_, seed=rank).scramble(u8o, out=ro.inputs)
            ro.capture(warmup=1)
            for i in range(3):
                ro.step()
            barrier()
            o0, o1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            Ko = max(10, K // 2)
            o0.record()
            for i in range(Ko):
                ro.step()
            o1.record()
            barrier()
            to = o0.elapsed_time(o1) / 1000.0
            if world > 1:
                tt = torch.tensor([to], device=dev, dtype=torch.float64)
                dist.all_reduce(tt, op=dist.ReduceOp.MAX)
                to = float(tt[0])
            others[ow] = {"workload": odesc, "value": world * oB * Ko / to, "unit": "images/s", "ms_per_step": 1000.0 * to / Ko,
                          "global_batch": world * oB, "steps": Ko,
                          "step_tensor_frac": oB * TRAIN_GFLOP_PER_IMAGE[ow] * 1e9 / (to / Ko) / 1e12 / peaks_for_others["tensor_sustained"]}
            ro.graph = None
            del ro, eo

    # ---- roofline: every tensor-core launch of the step timed alone (CUDA events on the launching stream, L2 flushed
    # before each launch), grouped by kernel; the kernel with the largest share of the step is the one reported ------
    peaks = measured_peaks()
    roof = None
    if rank == 0:
        def timed(fn, reps=5):
            ts = []
            for _ in range(reps):
                l2_flush.zero_()
                a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                a.record()
                fn()
                b.record()
                torch.cuda.synchronize()
                ts.append(a.elapsed_time(b) / 1000.0)
            return sorted(ts)[len(ts) // 2]

        from splitvae_b200._lib import KERNEL_NAMES
        by_kernel = {}
        for i, L in enumerate(e.debug_layers()):
            flops = 2.0 * B * L.Ho * L.Wo * L.Co * L.kh * L.kw * L.Ci        # algorithmic (unpadded) FLOPs of one pass
            for p, kern in ((0, L.kern_fwd), (1, L.kern_dgrad), (2, L.kern_wgrad)):
                if not kern:
                    continue
                t = timed(lambda: e.debug_run_layer(i, p, 1, runner.inputs))
                k = by_kernel.setdefault(KERNEL_NAMES[kern], {"launch_groups": 0, "us": 0.0, "gflop": 0.0, "sm_us": 0.0})
                k["launch_groups"] += 1
                k["us"] += t * 1e6
                k["gflop"] += flops / 1e9
                # the halo weight-gradient launches occupy a fixed number of CTAs, one per SM (planner defaults, tc_kernels.cu:plan_halo_wgrad)
                ctas = 148
                if KERNEL_NAMES[kern] == "halo_wgrad_kernel":
                    ctas = int(os.environ.get("SV_HWG_SPLITS_S2", 37)) if L.stride == 2 else int(os.environ.get("SV_HWG_SPLITS", 28))
                k["sm_us"] += t * 1e6 * min(ctas, 148) / 148.0
        for k in by_kernel.values():
            k["tflops"] = k["gflop"] / k["us"] * 1e3 if k["us"] else 0.0   # GFLOP/us = PFLOP/s
        # Dominant kernel = largest share of the step's SM-TIME (isolated time x the fraction of the 148 SMs its launches occupy).  The
        # halo weight-gradient kernel is launched on 28-37 CTAs (one per SM) on purpose - it runs on auxiliary streams beside the dgrad
        # chain, SV_HWG_SPLITS sweep in DESIGN.md - so its isolated time is 4-5x its share of the step; every other tensor-core kernel
        # fills the GPU.  Both orderings are reported (by_kernel[*].us and .sm_us).
        for n, k in by_kernel.items():
            k.setdefault("sm_us", k["us"])
            k["sm_frac"] = k["sm_us"] / k["us"] if k["us"] else 1.0
        dom = max(by_kernel, key=lambda n: by_kernel[n]["sm_us"])
        d = by_kernel[dom]
        traffic = None
        tp = os.path.join(ROOT, "profiles", "ncu_traffic.json")   # dram bytes per launch from the committed ncu --set full capture
        if os.path.exists(tp):
            with open(tp) as f:
                traffic = json.load(f).get(wl, {}).get(dom)
        roof = {"kernel": dom, "bound": "tensor", "achieved": d["tflops"], "peak": peaks["tensor_burst"], "unit": "TFLOP/s",
                "frac": d["tflops"] / peaks["tensor_burst"], "traffic": traffic, "peak_source": peaks["source"],
                "algorithmic_gflop": d["gflop"], "us": d["us"], "launch_groups": d["launch_groups"],
                "note": "bf16x3: the forward launches of every layer but d5 issue THREE tcgen05 MMAs per algorithmic product (two where the weight pair is "
                        "N-stacked), so a forward kernel at frac f keeps the tensor pipe 2-3 f busy; achieved counts algorithmic FLOPs only. "
                        "sum over this kernel's launches in one step of (2*MAC, unpadded) / sum of their CUDA-event times, each launch "
                        "timed alone after an L2 flush; peak = bf16 burst (kernel timed in isolation); a 'launch group' is one layer pass "
                        "(wgrad groups include their split-K reduce launch); traffic = dram read+write bytes of this kernel's launches "
                        "in one step from the committed ncu --set full capture (profiles/ncu_traffic.json)",
                "dominant_by": "SM-time (isolated time x fraction of SMs occupied)",
                "longest_isolated": max(by_kernel, key=lambda n: by_kernel[n]["us"]),
                "by_kernel": by_kernel}
        if dom == "halo_wgrad_kernel":
            # the halo wgrads are launched on ~37 of the 148 SMs on purpose: they run on auxiliary streams beside the dgrad chain
            # (148-CTA launches starved the chain: 1.78 vs 1.68 ms/step), so their isolated time overstates their share of the step
            roof["sms_used"] = round(148.0 * d["sm_frac"], 1)
            roof["frac_of_sms_used"] = roof["frac"] / d["sm_frac"]
            roof["note"] += "; this kernel is launched on 28-37 of 148 SMs by design (it overlaps the dgrad chain), frac_of_sms_used = frac / (share of the SMs)"
        t_loss = timed(lambda: e.debug_pixel_loss(runner.inputs), reps=10)      # the likelihood kernel alone (no scalar reduction behind it)
        loss_bytes = B * LOSS_BYTES_PER_IMAGE[H]
        ach = loss_bytes / t_loss / 1e9
        roof["hbm_kernels"] = {"pixel_loss_kernel": {"achieved": ach, "peak": peaks["hbm"], "unit": "GB/s", "frac": ach / peaks["hbm"],
                                                      "algorithmic_bytes_per_launch": loss_bytes, "us_per_launch": t_loss * 1e6}}
        try:   # multi-tensor Keras Adam over an arena-sized scratch copy: 28 B / parameter; never allowed to break the line
            import ctypes as C
            from splitvae_b200 import _lib
            n = int(e.arena_floats)
            bufs = [torch.zeros(n, device=dev) for _ in range(4)]
            bufs[1].normal_()
            lib = _lib.load()
            vp = lambda t: C.c_void_p(t.data_ptr())
            run = lambda: lib.sv_adam_flat(vp(bufs[0]), vp(bufs[1]), vp(bufs[2]), vp(bufs[3]), C.c_int64(n), C.c_float(1e-4),
                                           C.c_void_p(torch.cuda.current_stream().cuda_stream))
            t_adam = timed(run, reps=10)
            adam_bytes = 28 * n
            roof["hbm_kernels"]["adam_kernel"] = {"achieved": adam_bytes / t_adam / 1e9, "peak": peaks["hbm"], "unit": "GB/s",
                                                  "frac": adam_bytes / t_adam / 1e9 / peaks["hbm"],
                                                  "algorithmic_bytes_per_launch": adam_bytes, "us_per_launch": t_adam * 1e6}
            del bufs
        except Exception as ex:   # pragma: no cover
            roof["hbm_kernels"]["adam_kernel"] = {"error": str(ex)[:200]}
        step_flops = B * TRAIN_GFLOP_PER_IMAGE[wl] * 1e9
        roof["step_tensor"] = {"achieved_tflops": step_flops / (t_dev / K) / 1e12, "peak_tflops": peaks["tensor_sustained"],
                               "frac": step_flops / (t_dev / K) / 1e12 / peaks["tensor_sustained"],
                               "note": "whole train step, algorithmic FLOPs (BASELINE.md section 3) / step time, vs sustained bf16 peak"}

    if rank == 0:
        cpu = None
        if not args.no_cpu_baseline and world == 1:     # (the CPU baseline is reported by the 1-GPU run only)
            sb = B                                      # the workload's own batch
            ips, threads, times = cpu_baseline(model, H, sb, patch, beta, alpha, 3, min_seconds=10.0)
            cpu = {"value": ips, "unit": "images/s", "cores": threads, "kind": "port",
                   "sample": f"{len(times)} oracle train steps of {sb} images ({H}x{H}x3) = {sum(times):.1f} s of CPU work after 1 warm-up, "
                             f"torch CPU fp32 on {threads} threads (TF 2.0 not installable)"}
        total_images = world * B * K
        line = {
            "metric": "train images/sec", "value": total_images / t_dev, "unit": "images/s", "n_gpus": world, "steps": K,
            "warmup": Wm, "ms_per_step": 1000.0 * t_dev / K, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": {"bf16x3": "bf16x3", "bf16": "bf16", "fp32": "f32"}[args.precision], "data": "synthetic",
            "config": {"workload": desc, "model": model, "global_batch": world * B, "parallelism": f"dp{world}",
                       "gradient_exchange": ("none" if world == 1 else "fused NVLS multimem reduce-scatter + Adam(shard) + all-gather kernel" if nvls
                                             else "bucketed NCCL all-reduce (3 buckets) overlapped with backward"),
                       "cuda_graph": not args.no_graph, "l2": "per-step working set (activations + weights + Adam state) exceeds the 126 MB L2; no flush between steps",
                       "noise": "in-kernel Philox", "precision": args.precision, "operands": OPERANDS[args.precision]},
            "e2e": {"value": total_images / t_e2e, "unit": "images/s", "h2d_bytes_per_step": int(dev_u8.numel() + dev_perm.numel() * 4),
                    "d2h_bytes_per_step": 32, "ms_per_step": 1000.0 * t_e2e / K},
            "gpu_launches": int(launches_per_step * K),
            "launches_per_step": int(launches_per_step),
            "clocks": clocks, "roofline": roof, "cpu_baseline": cpu, "final_total_loss": final_total,
        }
        if fast is not None:
            line["fast_mode"] = fast
        if others is not None:
            line["other_workloads"] = others
        emit(line)
    if world > 1:
        # release the captured graph (it holds NCCL kernels) before tearing the communicator down; destroy_process_group()
        # was observed to block forever with a live captured collective, so leave the teardown to process exit
        runner.graph = None
        torch.cuda.synchronize()
        dist.barrier()
        sys.stdout.flush()
        sys.stderr.flush()
        os._exit(0)


if __name__ == "__main__":
    main()
