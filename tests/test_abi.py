"""The C-ABI library loads and exports every symbol include/splitvae.h declares; plan-only handles (no GPU
needed) expose the reference's variable inventory; the product path fails loudly without a device."""
import ctypes as C
import os
import re

import numpy as np
import pytest

from oracle import splitvae_oracle as O
from splitvae_b200 import _lib
from splitvae_b200.engine import Engine

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _header_symbols():
    src = open(os.path.join(ROOT, "include", "splitvae.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(sv_[a-z_0-9]+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    lib = _lib.load()
    syms = _header_symbols()
    assert len(syms) >= 29
    missing = [s for s in syms if not hasattr(lib, s)]
    assert not missing, missing
    assert sorted(_lib.SYMBOLS) == syms                      # the ctypes binding covers the whole header
    assert b"sm_100a" in lib.sv_version()


@pytest.mark.parametrize("model,H", [("lgvae", 32), ("lgvae", 64), ("lggmvae", 32), ("lggmvae", 64)])
def test_plan_only_inventory_matches_reference_variables(model, H):
    e = Engine(model=model, height=H, width=H, batch=8, plan_only=True)
    ref = O.init_params(model, H, H)
    assert [t[0] for t in e.table] == list(ref.keys())       # Keras creation order
    for name, shape, off, cnt in e.table:
        assert tuple(shape) == tuple(ref[name].shape), name
        assert cnt == ref[name].size and off % 64 == 0
    assert sum(t[3] for t in e.table) == sum(v.size for v in ref.values())
    assert e.arena_floats >= sum(t[3] for t in e.table) and e.workspace_bytes > 0


def test_invalid_configs_are_rejected_with_messages():
    lib = _lib.load()
    for kw, frag in [(dict(height=30), b"image size"), (dict(batch=0), b"batch"), (dict(y_size=64, model=1), b"y_size"),
                     (dict(global_latent_dims=64), b"latent")]:
        base = dict(model=0, height=32, width=32, batch=4, global_latent_dims=128, local_latent_dims=128, y_size=30, tau=0.4,
                    beta=1.0, alpha=1.0, learning_rate=1e-4, world_size=1, precision=0, flags=_lib.SV_FLAG_PLAN_ONLY)
        base.update(kw)
        if "height" in kw:
            base["width"] = kw["height"]
        cfg = _lib.SvConfig(**base)
        h = C.c_void_p()
        assert lib.sv_create(C.byref(cfg), C.byref(h)) == 1   # SV_ERR_INVALID
        assert frag in lib.sv_last_error(None)


def test_no_cpu_fallback():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(_lib.SplitVaeError):
        Engine(model="lgvae", height=32, width=32, batch=4)
    # the raw C entry point refuses too (SV_ERR_DEVICE) and a plan-only handle refuses compute calls (SV_ERR_STATE)
    lib = _lib.load()
    cfg = _lib.SvConfig(0, 32, 32, 4, 128, 128, 30, 0.4, 1.0, 1.0, 1e-4, 1, 0, 0)
    h = C.c_void_p()
    assert lib.sv_create(C.byref(cfg), C.byref(h)) == 2
    assert b"no CPU fallback" in lib.sv_last_error(None)
    e = Engine(model="lgvae", height=32, width=32, batch=4, plan_only=True)
    assert lib.sv_adam_step(e.h, None) == 3
    assert lib.sv_train_step(e.h, None, None, None, None, None) == 3
    from splitvae_b200 import trainer
    with pytest.raises(_lib.SplitVaeError):
        trainer.discretised_logistic_loss(torch.zeros(4), torch.zeros(4), torch.zeros(4))


def test_product_does_not_import_the_oracle():
    pkg = os.path.join(ROOT, "splitvae_b200")
    for dirpath, _, files in os.walk(pkg):
        for fn in files:
            if fn.endswith(".py"):
                src = open(os.path.join(dirpath, fn)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle|import_module\(.oracle", src, flags=re.M), fn


def test_header_is_plain_c_and_a_c_host_links(tmp_path):
    """The boundary is a C ABI: include/splitvae.h compiles as C99 (-pedantic, no warnings) and examples/c_host.c - a non-Python host that
    binds caller-owned buffers, captures the train step as a CUDA graph and replays it - links against the library; its plan-only mode
    runs without a GPU and prints the same variable inventory the Python host sees."""
    import shutil
    import subprocess
    gcc = shutil.which("gcc")
    if not gcc:
        pytest.skip("no gcc")
    cudart = "/usr/local/cuda/lib64"
    if not os.path.exists(os.path.join(cudart, "libcudart.so")):
        pytest.skip("no CUDA runtime to link the example's --run mode against")
    exe = str(tmp_path / "c_host")
    lib_dir = os.path.join(ROOT, "splitvae_b200")
    r = subprocess.run([gcc, "-std=c99", "-Wall", "-Wextra", "-pedantic", "-Werror", "-I" + os.path.join(ROOT, "include"),
                        os.path.join(ROOT, "examples", "c_host.c"), "-o", exe, "-L" + lib_dir, "-lsplitvae", "-Wl,-rpath," + lib_dir,
                        "-L" + cudart, "-lcudart"], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    out = subprocess.run([exe], capture_output=True, text=True, timeout=120)
    assert out.returncode == 0, out.stderr
    e = Engine(model="lgvae", height=64, width=64, batch=256, plan_only=True)
    lines = out.stdout.splitlines()
    assert "sm_100a" in lines[0]
    assert lines[1].split()[:6] == ["variables", str(len(e.table)), "arena_floats", str(e.arena_floats), "workspace_bytes", str(e.workspace_bytes)]
    for line, (name, shape, off, cnt) in zip(lines[2:], e.table):
        f = line.split()
        assert f[0] == name and f[1] == "[" + ",".join(str(d) for d in shape) + "]" and int(f[3]) == off and int(f[5]) == cnt
