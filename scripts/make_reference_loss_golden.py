"""Golden vectors for the loss block, produced by executing the REFERENCE'S OWN SOURCE TEXT.

TensorFlow 2.0 cannot be installed here, but `kl_divergence`, `kl_divergence_two_gauss`, `discretised_logistic_loss`
(vae/trainer.py:11-38) and the categorical-KL lines of the train step (vae/trainer.py:160-161) are pure elementwise `tf.*`
arithmetic.  This script reads those lines from /root/reference/vae/trainer.py (nothing is copied into the repo), executes
them against a small numpy stand-in for the `tf` namespace (exp / log / square / sigmoid / softplus / softmax / where /
maximum / reduce_sum / reduce_mean, evaluated in float64), and writes inputs + outputs to tests/golden/reference_losses.json.
The oracle (tests/test_oracle.py) and the CUDA kernels (tests/test_gpu_ops.py) are then checked against these vectors.

    python scripts/make_reference_loss_golden.py          # needs /root/reference (build container only)
"""
import json
import os
import re
import types

import numpy as np

REF = "/root/reference/vae/trainer.py"
OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden", "reference_losses.json")


def numpy_tf():
    tf = types.SimpleNamespace()
    tf.math = types.SimpleNamespace(log=np.log, square=np.square)
    tf.square = np.square
    tf.exp = np.exp
    tf.maximum = np.maximum
    tf.where = np.where
    tf.reduce_sum = lambda x, axis=None: np.sum(x, axis=tuple(axis) if isinstance(axis, list) else axis)
    tf.reduce_mean = lambda x, axis=None: np.mean(x, axis=tuple(axis) if isinstance(axis, list) else axis)

    def sigmoid(x):
        x = np.asarray(x, np.float64)
        e = np.exp(-np.abs(x))
        return np.where(x >= 0, 1.0 / (1.0 + e), e / (1.0 + e))

    def softmax(x, axis=-1):
        z = x - np.max(x, axis=axis, keepdims=True)
        e = np.exp(z)
        return e / np.sum(e, axis=axis, keepdims=True)

    tf.nn = types.SimpleNamespace(sigmoid=sigmoid, softplus=lambda x: np.logaddexp(0.0, x), softmax=softmax)
    return tf


def reference_functions():
    text = open(REF).read().split("\n")
    a = next(i for i, l in enumerate(text) if l.startswith("def kl_divergence("))
    b = next(i for i, l in enumerate(text) if l.startswith("def linear_assignment("))
    ns = {"tf": numpy_tf(), "np": np}
    exec("\n".join(text[a:b]), ns)                              # vae/trainer.py:11-38, verbatim
    # the categorical KL of train_step_lg_gm_vae (vae/trainer.py:160-161): two statements, de-indented, wrapped in a function
    i = next(i for i, l in enumerate(text) if re.match(r"\s+py = tf\.nn\.softmax\(y_logits, axis=1\)", l))
    body = [l.strip() for l in text[i:i + 2]]
    assert body[1].startswith("y_kl_loss = "), body
    model = types.SimpleNamespace(y_size=None)
    src = "def y_kl(y_logits, model):\n    " + "\n    ".join(body) + "\n    return y_kl_loss\n"
    exec(src, ns)
    return ns, (a + 1, b), i + 1


def main():
    ns, (la, lb), ly = reference_functions()
    rng = np.random.default_rng(20260117)
    B, D, K = 5, 128, 30
    zm, zs = rng.standard_normal((B, D)), np.exp(0.5 * rng.standard_normal((B, D)))
    pm, ps = rng.standard_normal((B, D)), np.exp(0.3 * rng.standard_normal((B, D)))
    logits = 2.0 * rng.standard_normal((B, K))
    n = 768
    k = rng.integers(0, 256, n)
    k[:64] = 0
    k[64:128] = 255                                               # both edge bins
    x = k / 255.0 * 2 - 1
    m = rng.uniform(-1.5, 1.5, n)
    ls = rng.uniform(-7.0, 2.0, n)
    m[128:256] = x[128:256] + rng.uniform(-3, 3, 128) * np.exp(ls[128:256])      # make the narrow-scale branches fire
    model = types.SimpleNamespace(y_size=K)
    G = {
        "source": f"{REF} lines {la}-{lb - 1} and {ly}-{ly + 1}, executed verbatim against a float64 numpy stand-in for tf",
        "inputs": {"z_mean": zm.tolist(), "z_sig": zs.tolist(), "prior_mean": pm.tolist(), "prior_sig": ps.tolist(),
                   "y_logits": logits.tolist(), "x": x.tolist(), "m": m.tolist(), "log_scales": ls.tolist()},
        "kl_divergence": float(ns["kl_divergence"](zm, zs)),
        "kl_divergence_two_gauss": float(ns["kl_divergence_two_gauss"](zm, zs, pm, ps)),
        "kl_divergence_two_gauss_std_normal": float(ns["kl_divergence_two_gauss"](zm, zs, 0., 1.)),
        "y_kl": float(ns["y_kl"](logits, model)),
        "discretised_logistic_loss": np.asarray(ns["discretised_logistic_loss"](x, m, ls), np.float64).tolist(),
    }
    with open(OUT, "w") as f:
        json.dump(G, f)
    print("wrote", OUT, {k: v for k, v in G.items() if isinstance(v, float)})


if __name__ == "__main__":
    main()
