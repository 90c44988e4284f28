"""Golden vectors for the patch-scramble augmentation, produced by executing the reference's own `Augmentator.scramble`
(/root/reference/augmentation.py:43-57, imported unmodified) against a numpy stand-in for the tensorflow names it uses.
The stand-in supplies the library semantics of tf.image.extract_patches (VALID, stride = size: patch q = pr * G + pc,
elements ordered row, column, channel), tf.reshape / split / unstack / concat, and tf.random.shuffle driven by an injected
permutation (shuffled[i] = patches[perm[i]]); the reassembly order is the reference's code.

    python scripts/make_reference_scramble_golden.py     # needs /root/reference (build container only)
Writes tests/golden/reference_scramble.json; tests/test_oracle.py checks oracle.scramble against it.
"""
import importlib.util
import json
import os
import sys
import types

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = "/root/reference/augmentation.py"
PERM = []


def extract_patches(x, sizes, strides, rates, padding):
    assert padding == "VALID" and sizes == strides and rates == [1, 1, 1, 1]
    p = sizes[1]
    n, H, W, C = x.shape
    out = np.zeros((n, H // p, W // p, p * p * C), x.dtype)
    for pr in range(H // p):
        for pc in range(W // p):
            out[:, pr, pc, :] = x[:, pr * p:(pr + 1) * p, pc * p:(pc + 1) * p, :].reshape(n, -1)
    return out


def install():
    tf = types.ModuleType("tensorflow")
    tf.image = types.SimpleNamespace(extract_patches=extract_patches)
    tf.expand_dims = np.expand_dims
    tf.reshape = lambda x, shape: np.reshape(x, shape)
    tf.random = types.SimpleNamespace(shuffle=lambda x: x[np.asarray(PERM.pop(0))])
    tf.split = lambda x, n, axis=0: np.split(x, n, axis=axis)
    tf.unstack = lambda x: [x[i] for i in range(x.shape[0])]
    tf.concat = lambda xs, axis: np.concatenate(list(xs), axis=axis)
    tf.convert_to_tensor = np.asarray
    sys.modules["tensorflow"] = tf
    sys.modules["tensorflow_probability"] = types.ModuleType("tensorflow_probability")


def main():
    install()
    spec = importlib.util.spec_from_file_location("reference_augmentation", REF)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    cases = []
    rng = np.random.default_rng(424242)
    for H, p in ((8, 2), (16, 4), (32, 1), (32, 4), (64, 8)):
        aug = mod.Augmentator("scramble", size=p)
        x = rng.integers(0, 256, (H, H, 3)).astype(np.float64)
        perm = rng.permutation((H // p) ** 2)
        PERM.append(perm)
        out = aug.augment(x)                                   # the reference's scramble -> [H, H, 6]
        assert out.shape == (H, H, 6) and np.array_equal(out[..., :3], x)
        cases.append({"H": H, "p": p, "x": x.astype(int).reshape(-1).tolist(), "perm": perm.tolist(),
                      "x_hat": out[..., 3:].astype(int).reshape(-1).tolist()})
    path = os.path.join(ROOT, "tests", "golden", "reference_scramble.json")
    with open(path, "w") as f:
        json.dump({"source": f"{REF} Augmentator.scramble (lines 43-57), imported unmodified", "cases": cases}, f)
    print("wrote", path, [(c["H"], c["p"]) for c in cases])


if __name__ == "__main__":
    main()
