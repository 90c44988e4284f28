"""GPU property tests at BASELINE.json's FULL sizes (LGVae 64x64 B=256, LGGMVae 64x64 B=256), where the fp64 oracle would take
minutes: size-independent properties of the train step, through the C-ABI.

  * the tensor-core path and the fp32-accumulate SIMT reference kernels (same bf16 storage) agree on the step's scalars and on
    every gradient tensor at full size;
  * the step is deterministic: split-K and bias-gradient reductions run in a fixed order, so two runs from the same state are
    bit-identical (this is what makes data-parallel replicas stay in lock-step);
  * a CUDA-graph replay of the captured step equals the eager step bit for bit;
  * gradients are linear in the 1/(B*world) scale: world_size=2 halves every gradient exactly (power-of-two scaling).
"""
import numpy as np
import pytest
import torch

from helpers import make_engine, rel_l2

pytestmark = pytest.mark.gpu

FULL = [("lgvae", 64, 256, 120.0, 40.0), ("lggmvae", 64, 256, 120.0, 40.0)]


def _inputs(model, H, B, seed=0):
    g = torch.Generator(device="cuda").manual_seed(seed)
    k = torch.randint(0, 256, (B, H, H, 6), generator=g, device="cuda")
    x = (k.float() / 255.0 * 2 - 1).contiguous()                      # the reference's value grid (vae/data.py:52)
    eg = torch.randn(B, 128, generator=g, device="cuda")
    el = torch.randn(B, 128, generator=g, device="cuda")
    u = torch.rand(B, 30, generator=g, device="cuda").clamp_(1e-6, 1 - 1e-6) if model == "lggmvae" else None
    return x, eg, el, u


def _step(e, x, eg, el, u):
    e.forward(x, eg, el, u)
    e.loss_fwd_bwd(x)
    for s in range(len(e.segments)):
        e.backward_segment(s)
    torch.cuda.synchronize()
    return e.scalars(), e.grads.clone()


@pytest.mark.parametrize("model,H,B,beta,alpha", FULL)
def test_full_size_tensor_core_vs_simt_reference(model, H, B, beta, alpha):
    x, eg, el, u = _inputs(model, H, B)
    tc = make_engine(model, H, B, "bf16", beta, alpha)
    tc.init_params(seed=7)
    ref = make_engine(model, H, B, "bf16", beta, alpha, no_tc=True)   # same bf16 storage, fp32 FMA chains instead of tcgen05
    ref.params.copy_(tc.params)
    ref.params_updated()
    sc, g = _step(tc, x, eg, el, u)
    rsc, rg = _step(ref, x, eg, el, u)
    for k, v in rsc.items():
        tol = 1e-3 * max(1.0, abs(v)) if k in ("total", "recon_x", "recon_x_hat") else 2e-2 * max(0.05, abs(v))
        assert abs(sc[k] - v) <= tol, (k, sc[k], v)
    worst = ("", 0.0)
    for name, shape, off, cnt in tc.table:
        a, b = g[off:off + cnt].double(), rg[off:off + cnt].double()
        den = b.norm().item()
        if den < 1e-9:
            continue
        r = (a - b).norm().item() / den
        if r > worst[1]:
            worst = (name, r)
        assert r <= 8e-2, (name, r)
    print(f"{model} full size: worst gradient rel-L2 TC vs SIMT = {worst[1]:.2e} ({worst[0]})")


# The default (benchmarked) mode at the sizes the bench runs, against an INDEPENDENT checker: the repo's fp32 mode, whose kernels share
# nothing with the tensor-core path and which is itself oracle-checked to 2e-4 (tests/test_gpu_parity.py).  Same gates as against the
# fp64 oracle at small sizes: every scalar rel 1e-3, every gradient tensor rel-L2 1e-2.
FULL_X3 = FULL + [("lgvae", 32, 64, 1.0, 40.0), ("lggmvae", 32, 256, 40.0, 40.0)]


@pytest.mark.parametrize("model,H,B,beta,alpha", FULL_X3)
def test_full_size_bf16x3_vs_fp32_mode(model, H, B, beta, alpha):
    x, eg, el, u = _inputs(model, H, B)
    ref = make_engine(model, H, B, "fp32", beta, alpha)
    ref.init_params(seed=7)
    rsc, rg = _step(ref, x, eg, el, u)
    report = []
    for prec, stol, gtol in (("bf16x3", 1e-3, 1e-2), ("bf16", None, 0.2)):
        e = make_engine(model, H, B, prec, beta, alpha)
        e.params.copy_(ref.params)
        e.params_updated()
        sc, g = _step(e, x, eg, el, u)
        for k, v in rsc.items():
            tol = (stol or (1e-3 if k in ("total", "recon_x", "recon_x_hat") else 2e-2)) * max(abs(v), 1e-3)
            assert abs(sc[k] - v) <= tol, (prec, k, sc[k], v)
        rows = []
        for name, shape, off, cnt in e.table:
            a, b = g[off:off + cnt].double(), rg[off:off + cnt].double()
            den = b.norm().item()
            if den < 1e-9:
                continue
            rows.append(((a - b).norm().item() / den, float(torch.dot(a, b) / (a.norm() * b.norm())), name))
        rows.sort(reverse=True)
        report.append(f"{prec}: worst gradient rel-L2 {rows[0][0]:.2e} cos {rows[0][1]:.6f} ({rows[0][2]}), median {rows[len(rows) // 2][0]:.2e}")
        assert rows[0][0] <= gtol, (prec, rows[:4])
        del e
    print(f"{model} {H}x{H} B={B} vs fp32 mode: " + "; ".join(report))


@pytest.mark.parametrize("model,H,B,beta,alpha", FULL[:1])
def test_full_size_step_is_deterministic_and_graph_replay_is_exact(model, H, B, beta, alpha):
    from splitvae_b200.trainer import StepRunner
    x, eg, el, u = _inputs(model, H, B, seed=1)
    e = make_engine(model, H, B, "bf16x3", beta, alpha)
    e.init_params(seed=8)
    p0 = e.params.clone()
    _, g1 = _step(e, x, eg, el, u)
    _, g2 = _step(e, x, eg, el, u)
    assert torch.equal(g1, g2)                                         # fixed-order reductions: bit-identical gradients
    # eager train step vs CUDA-graph replay from the same state (explicit noise so that both see the same epsilon)
    e.train_step(x, eg, el, u)
    torch.cuda.synchronize()
    eager = e.params.clone()
    e.params.copy_(p0); e.adam_m.zero_(); e.adam_v.zero_(); e.iterations = 0
    e.params_updated()
    r = StepRunner(e, use_graph=True, explicit_noise=True)
    r.capture(warmup=1)
    r.step(x, eg, el, u)
    torch.cuda.synchronize()
    assert e.iterations == 1
    assert torch.equal(e.params, eager)


def test_gradients_scale_exactly_with_world_size():
    model, H, B = "lgvae", 64, 32
    x, eg, el, u = _inputs(model, H, B, seed=2)
    a = make_engine(model, H, B, "bf16x3", 120.0)
    a.init_params(seed=9)
    b = make_engine(model, H, B, "bf16x3", 120.0, world_size=2)
    b.params.copy_(a.params)
    b.params_updated()
    _, ga = _step(a, x, eg, el, u)
    _, gb = _step(b, x, eg, el, u)
    # the loss gradient enters the network scaled by 1/(B*world): a power of two, so every bf16 / fp32 rounding is unchanged
    assert torch.equal(ga, gb * 2)


def test_c_abi_graph_capture_and_replay_equal_eager_steps():
    """sv_capture_graph / sv_replay: a host without torch's graph API gets the captured step from the C-ABI.  Three
    replays from the same state are bit-identical to three eager sv_train_step calls (explicit noise: both see the same epsilon)."""
    model, H, B = "lgvae", 64, 32
    x, eg, el, u = _inputs(model, H, B, seed=3)
    e = make_engine(model, H, B, "bf16x3", 120.0)
    e.init_params(seed=10)
    p0 = e.params.clone()
    for _ in range(3):
        e.train_step(x, eg, el, u)
    torch.cuda.synchronize()
    eager = e.params.clone()
    e.params.copy_(p0); e.adam_m.zero_(); e.adam_v.zero_(); e.iterations = 0
    e.params_updated()
    s = torch.cuda.Stream()
    s.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(s):
        e.capture_graph(x, eg, el, u)
        assert e.iterations == 0 and torch.equal(e.params, p0)          # capture records, nothing ran
        e.replay(3)
    s.synchronize()
    assert e.iterations == 3
    assert torch.equal(e.params, eager)
