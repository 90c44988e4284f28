"""Drop-in mirrors of the reference's model classes (vae/model.py:174-275) on top of libsplitvae.

Same constructor arguments, attributes, call / encode / decode / encode_y / get_y signatures and
output-tuple order as the TF classes; tensors are CUDA torch tensors (NHWC, float32).  Like a
Keras model the device engine is built on first call, when the batch size becomes known.
Returned tensors are views into the engine's device buffers and are overwritten by the next call.
"""
from __future__ import annotations

import torch

from .engine import Engine


def keras_weight_names(kind):
    """Attribute-path variable name of this build -> the name Keras gives the same variable in the reference's `save_weights` HDF5
    file (group = the sub-model layer, dataset = `<model>/<layer>/<sublayer>/<kernel|bias>:0`).  DERIVED from Keras' auto-naming rules
    (snake-cased class names, one global counter per layer class, creation order = vae/model.py:34-79,152-156,182-186,230-234,284-285)
    because TensorFlow cannot be installed here to read a real file; scripts/convert_checkpoint.py uses it on a machine with h5py."""
    model_name = {"lgvae": "lg_vae", "lggmvae": "lggm_vae", "gmvae": "gm_vae"}[kind]
    counters = {}

    def auto(cls):
        n = counters.get(cls, 0)
        counters[cls] = n + 1
        return cls if n == 0 else f"{cls}_{n}"

    names = {}

    def conv_encoder(attr):
        layer = auto("encoder")
        for v in ("e1", "e2", "e3"):
            names[f"{attr}.{v}"] = f"{model_name}/{layer}/{auto('conv2d')}"
        for v in ("e4_mean", "e4_sd"):
            names[f"{attr}.{v}"] = f"{model_name}/{layer}/{auto('dense')}"

    def gm_encoder(attr):
        layer = auto("encoder")
        seq = auto("sequential")
        for i in range(3):
            names[f"{attr}.h_block.{i}"] = f"{model_name}/{layer}/{seq}/{auto('conv2d')}"
        seq = auto("sequential")
        names[f"{attr}.y_block.0"] = f"{model_name}/{layer}/{seq}/{auto('dense')}"
        names[f"{attr}.y_block.2"] = f"{model_name}/{layer}/{seq}/{auto('dense')}"
        names[f"{attr}.y_dense"] = f"{model_name}/{layer}/y_dense"
        names[f"{attr}.h_top_dense"] = f"{model_name}/{layer}/{auto('dense')}"
        names[f"{attr}.z_prior_mean"] = f"{model_name}/{layer}/z_prior_mean"
        names[f"{attr}.z_prior_sig"] = f"{model_name}/{layer}/z_prior_sig"
        for v in ("e1", "z_mean", "z_sig"):
            names[f"{attr}.{v}"] = f"{model_name}/{layer}/{auto('dense')}"

    def decoder(attr):
        layer = auto("decoder")
        names[f"{attr}.d1"] = f"{model_name}/{layer}/{auto('dense')}"
        for v in ("d2", "d3", "d4", "d5"):
            names[f"{attr}.{v}"] = f"{model_name}/{layer}/{auto('conv2d')}"

    if kind == "lgvae":
        conv_encoder("encoder_x")
    else:
        gm_encoder("encoder_x")
    if kind != "gmvae":
        conv_encoder("encoder_x_hat")
    decoder("decoder_x")
    if kind != "gmvae":
        decoder("decoder_x_hat")
    out = {}
    for k, v in names.items():
        out[k + ".kernel"] = v + "/kernel:0"
        out[k + ".bias"] = v + "/bias:0"
    return out


class _SplitModel:
    _kind = None

    def __init__(self, global_latent_dims, local_latent_dims, image_shape=None, y_size=30, tau=0.4,
                 variational=True, precision="bf16x3", **engine_kwargs):
        if not variational:
            raise NotImplementedError("Determiistic LG-AE not implemented")  # vae/model.py:202,250
        self.global_latent_dims = global_latent_dims
        self.local_latent_dims = local_latent_dims
        self.variational = variational
        self.image_shape = image_shape
        self.y_size = y_size
        self.tau = tau
        self.precision = precision
        self._engine_kwargs = dict(engine_kwargs)
        self.engine = None
        self._eval_engines = {}
        self._pending_params = None
        self._pending_optimizer = None      # (adam_m, adam_v, iterations) of a checkpoint loaded before the engine exists

    # -- engine management -------------------------------------------------------------------
    def build(self, batch_size, **overrides):
        kw = dict(self._engine_kwargs)
        kw.update(overrides)
        H, W = int(self.image_shape[1]), int(self.image_shape[2])
        if self.global_latent_dims != 128 or (self.local_latent_dims or 128) != 128:
            raise NotImplementedError("libsplitvae supports the reference default latent sizes (128/128) only")
        old = self.engine
        self.engine = Engine(model=self._kind, height=H, width=W, batch=int(batch_size), y_size=self.y_size or 30,
                             tau=self.tau or 0.4, precision=self.precision, **kw)
        e = self.engine
        if old is not None:        # a new TRAINING engine (another batch size / loss weights): the whole optimizer state moves over
            torch.cuda.synchronize()
            e.params.copy_(old.params); e.adam_m.copy_(old.adam_m); e.adam_v.copy_(old.adam_v)
            e.params_updated()
            e.iterations = old.iterations
        elif self._pending_params is not None:
            e.load_params(self._pending_params)
            if self._pending_optimizer is not None:
                self._apply_optimizer_state(*self._pending_optimizer)
                self._pending_optimizer = None
        else:
            e.init_params(seed=5)
        self._eval_engines = {}
        return e

    def _ensure(self, batch_size):
        """Engine for a forward-only call (model(x), encode, decode, encode_y, get_y).  The training engine serves its own batch size;
        any other batch size gets a separate forward engine fed with the current weights - the training engine, its Adam state,
        iteration count and captured graph are never replaced by such a call (the reference calls model(tf.zeros([8, ...])) before
        training and the visualiser decodes other batch sizes mid-training)."""
        if self.engine is None:
            return self.build(batch_size)
        if self.engine.B == int(batch_size):
            return self.engine
        return self.eval_engine(batch_size)

    def configure(self, **kw):
        """Loss weights / optimizer settings that live in the reference's `config` and optimizer objects."""
        self._engine_kwargs.update(kw)

    def eval_engine(self, batch_size):
        """A second engine for the evaluation batch size that shares nothing with the training engine but the parameter
        VALUES (copied device to device before every evaluation pass): rebuilding the training engine for a different
        batch would drop its Adam state and captured graph."""
        b = int(batch_size)
        if self.engine is not None and self.engine.B == b:
            return self.engine
        ev = self._eval_engines.get(b)
        if ev is None:
            kw = dict(self._engine_kwargs)
            kw["rng_stream"] = (int(kw.get("rng_stream", 0)) ^ 0x400000) + len(self._eval_engines)     # its own noise stream
            H, W = int(self.image_shape[1]), int(self.image_shape[2])
            ev = Engine(model=self._kind, height=H, width=W, batch=b, y_size=self.y_size or 30, tau=self.tau or 0.4,
                        precision=self.precision, **kw)
            ev._synced = None
            self._eval_engines[b] = ev
        # copy + re-pack the weights once per training iteration, not once per evaluation batch (~400 batches per SVHN pass)
        stamp = (id(self.engine), self.engine.iterations, self._weights_version) if self.engine is not None else ("pending", self._weights_version)
        if ev._synced != stamp:
            if self.engine is not None:
                ev.params.copy_(self.engine.params)
                ev.params_updated()
            elif self._pending_params is not None:
                ev.load_params(self._pending_params)
            ev._synced = stamp
        return ev

    _weights_version = 0

    # -- checkpoints (vae/trainer.py:421 model.save_weights; Keras variable names and layouts) ----
    def _keras_layers(self, named):
        """[(Keras layer name, [(Keras weight name, array), ...])] in Keras' order: model.layers = the sub-models in creation order,
        layer.weights = kernel, bias per sub-layer in creation order (= the order of keras_weight_names)"""
        layers = {}
        for ours, keras in keras_weight_names(self._kind).items():
            layers.setdefault(keras.split("/")[1], []).append((keras, named[ours]))
        return list(layers.items())

    def save_weights(self, path, include_optimizer=False):
        """`model.save_weights('models/<run>.h5')` (vae/trainer.py:421).  A path ending in `.h5` / `.hdf5` / `.keras.h5` writes the
        reference's format - a Keras HDF5 weights file (root attributes `layer_names` / `backend` / `keras_version`, one group per
        sub-model with `weight_names`, datasets `<model>/<layer>/<sublayer>/kernel:0`; variables in Keras layouts: conv HWIO, dense
        [in,out], bias [out]) - through splitvae_b200.hdf5_lite (h5py is not part of this image).  Any other path writes a NumPy `.npz`
        keyed by this build's attribute-path names (`encoder_x.e1.kernel`, ...) plus the `keras_names` map.
        With include_optimizer the Adam moments and the iteration counter are stored too (the reference cannot resume; this can): in the
        HDF5 file they live in a top-level `optimizer_weights` group, which Keras' load_weights ignores."""
        import json

        import numpy as np
        e = self.engine
        params = {name: a for name, a in e.get_params().items()}
        opt = None
        if include_optimizer:
            opt = (e._export(e.adam_m), e._export(e.adam_v), np.asarray(e.iterations, dtype=np.int64))
        if str(path).endswith((".h5", ".hdf5")):
            from . import hdf5_lite
            extra = None
            if opt is not None:
                names = keras_weight_names(self._kind)
                extra = {"optimizer_weights": {"Adam/iterations:0": opt[2]}}
                for slot, blob in (("m", opt[0]), ("v", opt[1])):
                    for ours, a in blob.items():
                        extra["optimizer_weights"][f"Adam/{names[ours][:-2]}/{slot}:0"] = a
            return hdf5_lite.save_keras_weights(str(path), self._keras_layers(params), extra_groups=extra)
        blob = dict(params)
        blob["keras_names"] = np.asarray(json.dumps(keras_weight_names(self._kind)))
        if opt is not None:
            for name, a in opt[0].items():
                blob["adam_m/" + name] = a
            for name, a in opt[1].items():
                blob["adam_v/" + name] = a
            blob["optimizer/iterations"] = opt[2]
        if not str(path).endswith(".npz"):
            path = str(path) + ".npz"
        np.savez(path, **blob)
        return path

    def _load_h5(self, path):
        """Keras HDF5 weights file -> the `.npz`-style blob load_weights works on (by-name lookup, like load_weights(by_name=False) on an
        identically built model: every variable of this architecture must be present with its shape)"""
        import numpy as np

        from . import hdf5_lite
        flat, f = hdf5_lite.load_keras_weights(str(path))
        names = keras_weight_names(self._kind)
        missing = [k for k in names.values() if k not in flat]
        if missing:
            raise KeyError(f"{path}: no dataset {missing[0]} (+{len(missing) - 1} more); the file holds e.g. {sorted(flat)[:3]}")
        blob = {ours: np.asarray(flat[keras], dtype=np.float32) for ours, keras in names.items()}
        if "optimizer_weights" in f.keys():
            g = f["optimizer_weights"]
            blob["optimizer/iterations"] = np.asarray(g["Adam/iterations:0"].read())
            for ours, keras in names.items():
                for slot, pre in (("m", "adam_m/"), ("v", "adam_v/")):
                    try:
                        blob[pre + ours] = np.asarray(g[f"Adam/{keras[:-2]}/{slot}:0"].read(), dtype=np.float32)
                    except KeyError:
                        pass                      # reported as "optimizer state incomplete" below
        return blob

    def load_weights(self, path):
        import numpy as np
        if str(path).endswith((".h5", ".hdf5")):
            blob = self._load_h5(path)
        else:
            with np.load(path) as z:
                blob = {k: z[k] for k in z.files}
        named = {k: v for k, v in blob.items() if "/" not in k and k != "keras_names"}
        self.set_weights_by_name(named)
        if "optimizer/iterations" in blob:       # resume: the Adam moments and the step count (LR schedule, bias correction) come back too
            missing = [p + k for p in ("adam_m/", "adam_v/") for k in named if p + k not in blob]
            if missing:
                raise KeyError(f"{path}: optimizer state incomplete, missing {missing[:3]} ...")
            state = ({k: blob["adam_m/" + k] for k in named}, {k: blob["adam_v/" + k] for k in named}, int(blob["optimizer/iterations"]))
            if self.engine is not None:
                self._apply_optimizer_state(*state)
            else:
                self._pending_optimizer = state      # applied by build(), when the engine exists

    def set_weights_by_name(self, named):
        self._pending_params = named
        self._weights_version += 1
        if self.engine is not None:
            self.engine.load_params(named)

    def _apply_optimizer_state(self, adam_m, adam_v, iterations):
        e = self.engine
        import numpy as np
        for arena, blob in ((e.adam_m, adam_m), (e.adam_v, adam_v)):
            host = arena.cpu()
            for name, shape, off, cnt in e.table:
                host[off:off + cnt] = torch.from_numpy(np.asarray(blob[name], dtype=np.float32).reshape(-1))
            arena.copy_(host)
        e.iterations = int(iterations)

    def get_weights_by_name(self):
        return self.engine.get_params()

    @property
    def trainable_variables(self):
        return list(self.engine.get_params().items())

    # -- reference API -----------------------------------------------------------------------
    def _split_outputs(self, e):
        dx, dxh = e.output("dec_x"), e.output("dec_x_hat")
        return dx[..., :3], dx[..., 3:], dxh[..., :3], dxh[..., 3:]

    def encode(self, inputs, eps_g=None, eps_l=None, u=None):
        """vae/model.py:204-209 / 252-257: returns the sampled (z_x, z_x_hat)."""
        e = self._ensure(inputs.shape[0])
        e.forward(inputs, eps_g, eps_l, u)
        return e.output("z_x"), e.output("z_x_hat")

    def decode(self, z_x, z_x_hat, rescale=True):
        """vae/model.py:211-218 / 259-266."""
        e = self._ensure(z_x.shape[0])
        e.decode(z_x.contiguous().float(), z_x_hat.contiguous().float())
        x_mean, _, x_hat_mean, _ = self._split_outputs(e)
        if rescale:
            return torch.clip((x_mean + 1) * 0.5, 0., 1.), torch.clip((x_hat_mean + 1) * 0.5, 0., 1.)
        return x_mean, x_hat_mean


class GMVae(_SplitModel):
    """Plain GMVAE, vae/model.py:277-320: gmvae encoder on x + ONE decoder fed by z_x.  Constructor signature of the reference
    (no local latent): GMVae(global_latent_dims, image_shape, y_size, tau)."""
    _kind = "gmvae"

    def __init__(self, global_latent_dims, image_shape, y_size, tau, variational=True, type="conv", **kw):
        super().__init__(global_latent_dims, 128, image_shape, y_size=y_size, tau=tau, variational=variational, **kw)
        self.local_latent_dims = None

    def __call__(self, inputs, training=False, eps_g=None, u=None):
        """vae/model.py:288-299: 9-tuple (x_mean, x_log_scale, z_x, z_mean_x, z_sig_x, y, y_logits, z_prior_mean, z_prior_sig)."""
        e = self._ensure(inputs.shape[0])
        e.forward(inputs, eps_g, None, u)
        dx = e.output("dec_x")
        o = e.output
        return (dx[..., :3], dx[..., 3:], o("z_x"), o("z_mean_x"), o("z_sig_x"), o("y"), o("y_logits"), o("z_prior_mean"), o("z_prior_sig"))

    def encode(self, inputs, eps_g=None, u=None):
        """vae/model.py:301-305: the sampled z_x."""
        e = self._ensure(inputs.shape[0])
        e.forward(inputs, eps_g, None, u)
        return e.output("z_x")

    def decode(self, z_x, rescale=True):
        """vae/model.py:307-312."""
        e = self._ensure(z_x.shape[0])
        e.decode(z_x.contiguous().float(), None)
        x_mean = e.output("dec_x")[..., :3]
        return torch.clip((x_mean + 1) * 0.5, 0., 1.) if rescale else x_mean

    def encode_y(self, y, rescale=True):
        """vae/model.py:314-316."""
        e = self._ensure(y.shape[0])
        e.encode_y(y.contiguous().float())
        return e.output("z_prior_mean"), e.output("z_prior_sig")

    def get_y(self, x, u=None):
        """vae/model.py:318-320 (6-channel batch, channels 0-2 used)."""
        e = self._ensure(x.shape[0])
        e.forward(x, None, None, u)
        return e.output("y"), e.output("y_logits")


class LGVae(_SplitModel):
    """SPLIT-VAE, vae/model.py:174-218."""
    _kind = "lgvae"

    def __init__(self, global_latent_dims, local_latent_dims, image_shape=None, variational=True, type="conv", **kw):
        super().__init__(global_latent_dims, local_latent_dims, image_shape, y_size=None, tau=None,
                         variational=variational, **kw)

    def __call__(self, inputs, eps_g=None, eps_l=None):
        """vae/model.py:189-200: 10-tuple (x_mean, x_log_scale, z_x, z_mean_x, z_sig_x, z_x_hat, x_hat_mean,
        x_hat_log_scale, z_mean_x_hat, z_sig_x_hat)."""
        e = self._ensure(inputs.shape[0])
        e.forward(inputs, eps_g, eps_l, None)
        x_mean, x_ls, xh_mean, xh_ls = self._split_outputs(e)
        o = e.output
        return (x_mean, x_ls, o("z_x"), o("z_mean_x"), o("z_sig_x"), o("z_x_hat"), xh_mean, xh_ls,
                o("z_mean_x_hat"), o("z_sig_x_hat"))


class LGGMVae(_SplitModel):
    """SPLIT-GMVAE, vae/model.py:221-275."""
    _kind = "lggmvae"

    def __init__(self, global_latent_dims, local_latent_dims, image_shape, y_size, tau, variational=True, type="conv", **kw):
        super().__init__(global_latent_dims, local_latent_dims, image_shape, y_size=y_size, tau=tau,
                         variational=variational, **kw)

    def __call__(self, inputs, training=False, eps_g=None, eps_l=None, u=None):
        """vae/model.py:237-248: the LGVae 10-tuple + (y, y_logits, z_prior_mean, z_prior_sig).  `training` is
        accepted and, as in the reference (model.py:242), has no effect (all dropout inactive)."""
        e = self._ensure(inputs.shape[0])
        e.forward(inputs, eps_g, eps_l, u)
        x_mean, x_ls, xh_mean, xh_ls = self._split_outputs(e)
        o = e.output
        return (x_mean, x_ls, o("z_x"), o("z_mean_x"), o("z_sig_x"), o("z_x_hat"), xh_mean, xh_ls,
                o("z_mean_x_hat"), o("z_sig_x_hat"), o("y"), o("y_logits"), o("z_prior_mean"), o("z_prior_sig"))

    def encode_y(self, y, rescale=True):
        """vae/model.py:268-270."""
        e = self._ensure(y.shape[0])
        e.encode_y(y.contiguous().float())
        return e.output("z_prior_mean"), e.output("z_prior_sig")

    def get_y(self, x, u=None):
        """vae/model.py:272-275.  The reference hands its 6-channel batch to the 3-channel encoder here
        (visualizer.py:328); this build takes the 6-channel batch and uses channels 0-2."""
        e = self._ensure(x.shape[0])
        e.forward(x, None, None, u)
        return e.output("y"), e.output("y_logits")
