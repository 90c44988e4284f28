"""Golden vectors for the model WIRING and the loss ASSEMBLY, produced by executing the reference's own source.

TensorFlow 2.0 is not installable here, so this script puts a small torch-backed (float64) stand-in for the handful of
`tensorflow` / `tf.keras` names that vae/model.py and the train steps of vae/trainer.py use into `sys.modules`, then
  * imports /root/reference/vae/model.py UNMODIFIED and builds its own `LGVae` / `LGGMVae` (layer attributes, call graphs,
    concat / slice order, activations, return-tuple order all come from the reference's code),
  * extracts the forward + loss lines of `train_step_lg_vae` / `train_step_lg_gm_vae` (vae/trainer.py, between
    `with tf.GradientTape()` and `tape.gradient`) together with the three loss functions (trainer.py:11-38) and executes them.
What the stand-in supplies is only the LIBRARY semantics: Conv2D 'same' / Dense / Flatten / Dropout(identity) / bilinear resize /
softmax etc. (the oracle's own primitives), noise drawn from a queue (eps_g, eps_l, u in the order the reference asks for them),
and the weights of oracle.init_params injected by Keras variable name.  Nothing from the reference is copied into the repo.

    python scripts/make_reference_model_golden.py        # needs /root/reference (build container only)
Writes tests/golden/reference_model_{lgvae,lggmvae}.json; tests/test_oracle.py checks oracle.model_forward / step_losses against them.
"""
import importlib.util
import json
import math
import os
import re
import sys
import types

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import splitvae_oracle as O  # noqa: E402

REF_MODEL = "/root/reference/vae/model.py"
REF_TRAINER = "/root/reference/vae/trainer.py"
NOISE = {"normal": [], "uniform": []}


# ------------------------------------------------------------------ the stand-in library
class Layer:
    def __init__(self, *a, name=None, **k):
        self.name = name

    def __call__(self, *args, **kwargs):
        return self.call(*args, **kwargs)


class Dense(Layer):
    def __init__(self, units, activation=None, name=None, bias_initializer=None, **k):
        super().__init__(name=name)
        self.units, self.activation, self.kernel, self.bias = units, activation, None, None

    def call(self, x):
        assert self.kernel.shape == (x.shape[-1], self.units), (self.kernel.shape, x.shape, self.units)
        return O._act(O.dense(x, self.kernel, self.bias), self.activation)


class Conv2D(Layer):
    def __init__(self, filters, kernel_size, strides=1, padding="valid", activation=None, name=None, **k):
        super().__init__(name=name)
        assert padding == "same"
        self.filters, self.k, self.s, self.activation, self.kernel, self.bias = filters, kernel_size, strides, activation, None, None

    def call(self, x):
        assert self.kernel.shape == (self.k, self.k, x.shape[-1], self.filters), (self.kernel.shape, x.shape)
        return O._act(O.conv2d_same(x, self.kernel, self.bias, self.s), self.activation)


class Flatten(Layer):
    def call(self, x):
        return x.reshape(x.shape[0], -1)


class Dropout(Layer):
    def __init__(self, rate=0.0, **k):
        super().__init__()

    def call(self, x, training=None):
        return x            # inactive in the reference train step (dropout note)


class BatchNormalization(Layer):
    pass                    # only used by the eval-only SVHN classifier


class Sequential(Layer):
    def __init__(self, layers=None, **k):
        super().__init__()
        self.layers = list(layers or [])

    def call(self, x, training=None):
        for l in self.layers:
            x = l(x)
        return x


def _resize(x, size):
    assert size[0] == 2 * x.shape[1] and size[1] == 2 * x.shape[2], (size, x.shape)   # the hot path only ever doubles
    return O.resize2x(x)


def _axis(a):
    return tuple(a) if isinstance(a, (list, tuple)) else a


def install():
    tf = types.ModuleType("tensorflow")
    keras = types.ModuleType("tensorflow.keras")
    layers = types.ModuleType("tensorflow.keras.layers")
    for c in (Layer, Dense, Conv2D, Flatten, Dropout, BatchNormalization):
        setattr(layers, c.__name__, c)
    keras.layers, keras.Model, keras.Sequential = layers, Layer, Sequential
    keras.initializers = types.SimpleNamespace(constant=lambda v: ("constant", v))
    tf.keras, tf.float32 = keras, "float32"
    tf.random = types.SimpleNamespace(
        normal=lambda shape, mean=0, stddev=1, dtype=None: _pop("normal", shape),
        uniform=lambda shape: _pop("uniform", shape))
    tf.reshape = lambda x, shape: x.reshape(list(shape))
    tf.concat = lambda xs, axis: torch.cat(list(xs), dim=axis)
    tf.image = types.SimpleNamespace(resize=_resize)
    tf.clip_by_value = lambda x, lo, hi: torch.clamp(x, lo, hi)
    T = lambda x: torch.as_tensor(x, dtype=torch.float64)          # TF accepts python scalars wherever it accepts tensors
    tf.exp, tf.square, tf.maximum = (lambda x: torch.exp(T(x))), (lambda x: torch.square(T(x))), lambda a, b: torch.clamp(a, min=b)
    tf.where = lambda c, a, b: torch.where(c, a, b)
    tf.reduce_sum = lambda x, axis=None: torch.sum(x) if axis is None else torch.sum(x, dim=_axis(axis))
    tf.reduce_mean = lambda x, axis=None: torch.mean(x) if axis is None else torch.mean(x, dim=_axis(axis))
    tf.math = types.SimpleNamespace(log=lambda x: torch.log(T(x)), square=lambda x: torch.square(T(x)))
    tf.nn = types.SimpleNamespace(softmax=lambda x, axis=-1: torch.softmax(x, dim=axis), sigmoid=torch.sigmoid,
                                  softplus=torch.nn.functional.softplus)
    sys.modules.update({"tensorflow": tf, "tensorflow.keras": keras, "tensorflow.keras.layers": layers})
    return tf


def _pop(kind, shape):
    t = NOISE[kind].pop(0)
    assert tuple(t.shape) == tuple(shape), (kind, t.shape, shape)
    return t


# ------------------------------------------------------------------ reference source
def load_reference_model():
    spec = importlib.util.spec_from_file_location("reference_vae_model", REF_MODEL)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)           # vae/model.py, unmodified
    return mod


def reference_step(fn_name, tf, outputs):
    """Source lines of a train step between `with tf.GradientTape() as tape:` and `gradients = tape.gradient(...)`, de-indented
    and wrapped in a function returning the named locals; plus the loss functions of trainer.py:11-38."""
    text = open(REF_TRAINER).read().split("\n")
    a = next(i for i, l in enumerate(text) if l.startswith("def kl_divergence("))
    b = next(i for i, l in enumerate(text) if l.startswith("def linear_assignment("))
    ns = {"tf": tf, "np": np}
    exec("\n".join(text[a:b]), ns)
    s = next(i for i, l in enumerate(text) if re.match(rf"\s+def {fn_name}\(model, images, optimizer\):", l))
    w = next(i for i in range(s, s + 5) if "with tf.GradientTape() as tape:" in text[i])
    e = next(i for i in range(w, w + 40) if "gradients = tape.gradient" in text[i])
    body = [l for l in text[w + 1:e] if l.strip()]
    ind = min(len(l) - len(l.lstrip()) for l in body)
    src = f"def step(model, images, config):\n" + "\n".join("    " + l[ind:] for l in body) + \
          "\n    return dict(" + ", ".join(f"{k}={v}" for k, v in outputs.items()) + ")\n"
    exec(src, ns)
    return ns["step"], (w + 2, e)


def inject(model, params):
    for name, arr in params.items():
        obj = model
        parts = name.split(".")
        for p in parts[:-1]:
            obj = obj.layers[int(p)] if p.isdigit() else getattr(obj, p)
        assert hasattr(obj, parts[-1]), name
        setattr(obj, parts[-1], torch.tensor(np.asarray(arr), dtype=torch.float64))


def digest(t):
    t = t.detach().double().reshape(-1)
    return {"n": int(t.numel()), "sum": float(t.sum()), "l2": float(t.norm()), "head": [float(v) for v in t[:4]]}


def main():
    tf = install()
    ref = load_reference_model()
    for kind, H, B, patch, beta, alpha in (("lgvae", 32, 3, 4, 7.0, 40.0), ("lggmvae", 32, 3, 4, 5.0, 3.0), ("gmvae", 32, 3, 4, 6.0, 2.0)):
        seed_base = 70
        params = O.init_params(kind, H, H, seed=5 + seed_base)
        b = O.synthetic_batch(B, H, patch, seed_base=seed_base)
        t64 = lambda a: torch.tensor(np.asarray(a), dtype=torch.float64)
        if kind == "lgvae":
            model = ref.LGVae(global_latent_dims=128, local_latent_dims=128, image_shape=[-1, H, H, 3])
            names = ["x_mean", "x_log_scale", "z_x", "z_mean_x", "z_sig_x", "z_x_hat", "x_hat_mean", "x_hat_log_scale", "z_mean_x_hat", "z_sig_x_hat"]
            outs = {"recon_x": "x_recon_loss", "recon_x_hat": "x_hat_recon_loss", "total_kl": "total_kl_loss", "kl_x": "x_kl_loss",
                    "kl_x_hat": "x_hat_kl_loss", "total": "total_loss"}
            step, lines = reference_step("train_step_lg_vae", tf, outs)
        elif kind == "gmvae":
            model = ref.GMVae(global_latent_dims=128, image_shape=[-1, H, H, 3], y_size=30, tau=0.4)
            names = ["x_mean", "x_log_scale", "z_x", "z_mean_x", "z_sig_x", "y", "y_logits", "z_prior_mean", "z_prior_sig"]
            outs = {"recon_x": "x_recon_loss", "kl_x": "x_kl_loss", "y_kl": "y_kl_loss", "total": "total_loss"}
            step, lines = reference_step("train_step_gm_vae", tf, outs)
        else:
            model = ref.LGGMVae(global_latent_dims=128, local_latent_dims=128, image_shape=[-1, H, H, 3], y_size=30, tau=0.4)
            names = ["x_mean", "x_log_scale", "z_x", "z_mean_x", "z_sig_x", "z_x_hat", "x_hat_mean", "x_hat_log_scale", "z_mean_x_hat", "z_sig_x_hat",
                     "y", "y_logits", "z_prior_mean", "z_prior_sig"]
            outs = {"recon_x": "x_recon_loss", "recon_x_hat": "x_hat_recon_loss", "kl_x": "x_kl_loss", "kl_x_hat": "x_hat_kl_loss",
                    "y_kl": "y_kl_loss", "total": "total_loss"}
            step, lines = reference_step("train_step_lg_gm_vae", tf, outs)
        inject(model, params)
        inputs = t64(b["inputs"])
        queue = lambda: NOISE.update(normal=[t64(b["eps_g"])] + ([] if kind == "gmvae" else [t64(b["eps_l"])]),
                                     uniform=[] if kind == "lgvae" else [t64(b["u"])])
        queue()
        tup = model(inputs) if kind == "lgvae" else model(inputs, training=True)
        assert len(tup) == len(names) and not NOISE["normal"] and not NOISE["uniform"]
        queue()
        sc = step(model, inputs, types.SimpleNamespace(beta=beta, alpha=alpha))
        # the rest of the model surface the visualiser / test steps call (vae/model.py:204-218, 252-275), on regenerable inputs
        api = {}
        z_a, z_b = 0.5 * t64(b["eps_g"]), 0.5 * t64(b["eps_l"])
        if kind != "gmvae":
            queue()
            api["encode"] = [digest(t) for t in model.encode(inputs)]
            api["decode_rescaled"] = [digest(t) for t in model.decode(z_a, z_b)]
            api["decode_raw"] = [digest(t) for t in model.decode(z_a, z_b, rescale=False)]
        if kind == "lggmvae":
            y_in = torch.softmax(torch.log(t64(b["u"])), dim=1)
            api["encode_y"] = [digest(t) for t in model.encode_y(y_in)]
            queue()
            api["get_y"] = [digest(t) for t in model.get_y(inputs[..., :3])]     # (the 3-channel image: model.py:272-275 does not slice)
        G = {"source": f"{REF_MODEL} (unmodified) + {REF_TRAINER} lines {lines[0]}-{lines[1]} and 11-38, executed against the torch float64 "
                       f"stand-in for tensorflow defined in scripts/make_reference_model_golden.py",
             "case": {"model": kind, "H": H, "B": B, "patch": patch, "beta": beta, "alpha": alpha, "seed_base": seed_base},
             "inputs_sum": float(np.asarray(b["inputs"], np.float64).sum()),
             "outputs": {n: digest(t) for n, t in zip(names, tup)},
             "output_order": names, "api": api,
             "scalars": {k: float(v) for k, v in sc.items()}}
        path = os.path.join(ROOT, "tests", "golden", f"reference_model_{kind}.json")
        with open(path, "w") as f:
            json.dump(G, f, indent=1)
        print("wrote", path, G["scalars"])


if __name__ == "__main__":
    main()
