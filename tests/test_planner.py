"""CPU: the kernel planner (pure host code inside libsplitvae, reached through a plan-only handle - no CUDA call is made).

Pins which tensor-core kernel serves every layer pass of the BASELINE.json configurations, that every conv / dense pass of the
hot path HAS a tensor-core kernel (nothing silently falls back to the SIMT reference kernels in bf16 mode), and that the
workspace stays far below one B200's 180 GB."""
import pytest

from splitvae_b200._lib import KERNEL_NAMES
from splitvae_b200.engine import Engine


def _plan(model, H, B, precision="bf16"):
    e = Engine(model=model, height=H, width=H, batch=B, plan_only=True, precision=precision)
    return e, {L.name.decode(): (KERNEL_NAMES[L.kern_fwd], KERNEL_NAMES[L.kern_dgrad], KERNEL_NAMES[L.kern_wgrad]) for L in e.debug_layers()}


def test_c2_layer_to_kernel_map():
    e, plan = _plan("lgvae", 64, 256)
    for enc in ("encoder_x", "encoder_x_hat"):
        # the stride-2 weight gradients with >= 16 output columns run on the halo kernel through the pixel-pair view
        assert plan[f"{enc}.e1"] == ("pconv_kernel", "reference", "halo_wgrad_kernel")   # first conv: no input gradient
        assert plan[f"{enc}.e2"] == ("pconv_kernel", "pconv_kernel", "halo_wgrad_kernel")
        assert plan[f"{enc}.e3"] == ("igemm_kernel", "igemm_kernel", "halo_wgrad_kernel")   # 8-pixel-wide output: K step = 8 columns x 2 rows
    for dec in ("decoder_x", "decoder_x_hat"):
        assert plan[f"{dec}.d3"] == ("nsconv_kernel", "nsconv_kernel", "halo_wgrad_kernel")
        assert plan[f"{dec}.d4"] == ("nsconv_kernel", "nsconv_kernel", "halo_wgrad_kernel")
        assert plan[f"{dec}.d5"] == ("nsconv_kernel", "pconv_kernel", "halo_wgrad_kernel")
    assert e.workspace_bytes < 4 << 30
    # the halo weight-gradient launches run on a fixed few CTAs (one per SM) beside the dgrad chain: 28 for stride-1 layers, <= 37 through the pair view
    ctas = {L.name.decode(): L.wgrad_ctas for L in e.debug_layers()}
    assert ctas["decoder_x.d4"] == ctas["decoder_x.d5"] == ctas["decoder_x.d2"] == 28 and ctas["encoder_x.e1"] == 37 and 30 <= ctas["encoder_x.e2"] <= 37


@pytest.mark.parametrize("model,H,B", [("lgvae", 32, 64), ("lgvae", 64, 256), ("lggmvae", 32, 256), ("lggmvae", 64, 256), ("lgvae", 16, 3)])
@pytest.mark.parametrize("precision", ["bf16x3", "bf16"])
def test_every_pass_of_the_hot_path_has_a_tensor_core_kernel(model, H, B, precision):
    e, plan = _plan(model, H, B, precision)
    assert len(plan) == (18 if model == "lgvae" else 22)
    for name, (fwd, dgrad, wgrad) in plan.items():
        assert fwd != "reference" and wgrad != "reference", (name, fwd, wgrad)
        first_conv = name.endswith(".e1") and "encoder_x_hat" in name or name in ("encoder_x.e1", "encoder_x.h_block.0")
        if not first_conv:
            assert dgrad != "reference", name


def test_planner_knobs_are_read_at_plan_time(monkeypatch):
    monkeypatch.setenv("SV_PCONV", "0")
    monkeypatch.setenv("SV_NO_NSCONV", "1")
    monkeypatch.setenv("SV_NO_PAIR_WGRAD", "1")
    _, plan = _plan("lgvae", 64, 256)
    assert plan["decoder_x.d5"][1] == "halo_conv_kernel"
    assert plan["decoder_x.d4"][0] == "igemm_kernel"
    assert plan["encoder_x.e1"][2] == plan["encoder_x.e2"][2] == "wgrad_kernel"      # per-tap fallback of the stride-2 weight gradients
    assert plan["decoder_x.d4"][2] == "halo_wgrad_kernel"
    monkeypatch.setenv("SV_NO_NARROW_WGRAD", "1")
    _, plan = _plan("lgvae", 64, 256)
    assert plan["decoder_x.d2"][2] == plan["encoder_x.e3"][2] == "wgrad_kernel"


def test_c2_layer_to_kernel_map_bf16x3():
    """The default mode: forward products on bf16 pairs (three MMAs) need twice the operand bytes, so the planner moves the
    resident-weight forward kernels of d3 / d4 to the halo kernel with streamed weights; d5 forward (single bf16 by design)
    and the whole backward pass keep the single-bf16 plan."""
    e, plan = _plan("lgvae", 64, 256, "bf16x3")
    _, fast = _plan("lgvae", 64, 256, "bf16")
    for name in plan:
        assert plan[name][1:] == fast[name][1:], name                      # dgrad / wgrad kernels unchanged
    for dec in ("decoder_x", "decoder_x_hat"):
        assert plan[f"{dec}.d5"][0] == "nsconv_kernel"
        assert plan[f"{dec}.d4"][0] in ("halo_conv_kernel", "pconv_kernel", "nsconv_kernel")
    info = {L.name.decode(): L for L in e.debug_layers()}
    assert all(L.split_fwd == (0 if n.endswith(".d5") else 1) for n, L in info.items())
    assert e.workspace_bytes < 4 << 30


@pytest.mark.parametrize("model,H", [("lgvae", 64), ("lggmvae", 32), ("gmvae", 32)])
def test_backward_segments_partition_the_arena(model, H):
    """The gradient buckets (sv_segment_range): three segments in backward order whose ranges are disjoint and together hold every
    variable exactly once - the data-parallel all-reduce and the per-segment Adam both rely on it."""
    e = Engine(model=model, height=H, width=H, batch=8, plan_only=True)
    import ctypes as C
    lib, off, cnt = e.lib, C.c_int64(), C.c_int64()
    segs = []
    for s in range(lib.sv_num_segments(e.h)):
        r = []
        for i in range(lib.sv_segment_num_ranges(e.h, s)):
            assert lib.sv_segment_range(e.h, s, i, C.byref(off), C.byref(cnt)) == 0
            r.append((off.value, cnt.value))
        segs.append(r)
    assert len(segs) == 3 and all(1 <= len(r) <= 2 for r in segs)
    flat = sorted(x for r in segs for x in r)
    assert all(a[0] + a[1] <= b[0] for a, b in zip(flat, flat[1:]))                 # disjoint
    assert sum(c for _, c in flat) == e.arena_floats                                # nothing left out
    owner = lambda o: next(i for i, r in enumerate(segs) if any(a <= o < a + c for a, c in r))
    for name, shape, o, c in e.table:
        seg = owner(o)
        if name.startswith("decoder"):
            assert seg == 0, name
        elif any(k in name for k in (".e1.", ".e2.", "h_block.0", "h_block.1")) and not (model != "lgvae" and name.startswith("encoder_x.e1")):
            assert seg == 2, name
        else:
            assert seg == 1, name
