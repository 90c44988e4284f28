"""Drop-in mirrors of the reference's model classes (vae/model.py:174-275) on top of libsplitvae.

Same constructor arguments, attributes, call / encode / decode / encode_y / get_y signatures and
output-tuple order as the TF classes; tensors are CUDA torch tensors (NHWC, float32).  Like a
Keras model the device engine is built on first call, when the batch size becomes known.
Returned tensors are views into the engine's device buffers and are overwritten by the next call.
"""
from __future__ import annotations

import torch

from .engine import Engine


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

    # -- engine management -------------------------------------------------------------------
    def build(self, batch_size, **overrides):
        kw = dict(self._engine_kwargs)
        kw.update(overrides)
        H, W = int(self.image_shape[1]), int(self.image_shape[2])
        if self.global_latent_dims != 128 or (self.local_latent_dims or 128) != 128:
            raise NotImplementedError("libsplitvae supports the reference default latent sizes (128/128) only")
        params = None
        if self.engine is not None:
            params = self.engine.get_params()
        self.engine = Engine(model=self._kind, height=H, width=W, batch=int(batch_size), y_size=self.y_size or 30,
                             tau=self.tau or 0.4, precision=self.precision, **kw)
        if params is not None:
            self.engine.load_params(params)
        elif self._pending_params is not None:
            self.engine.load_params(self._pending_params)
        else:
            self.engine.init_params(seed=kw.get("rng_stream", 0) * 0 + 5)
        return self.engine

    def _ensure(self, batch_size):
        if self.engine is None or self.engine.B != int(batch_size):
            self.build(batch_size)
        return self.engine

    def configure(self, **kw):
        """Loss weights / optimizer settings that live in the reference's `config` and optimizer objects."""
        self._engine_kwargs.update(kw)

    def eval_engine(self, batch_size):
        """A second engine for the evaluation batch size that shares nothing with the training engine but the parameter
        VALUES (copied device to device before every evaluation pass): rebuilding the training engine for a different
        batch would drop its Adam state and captured graph."""
        b = int(batch_size)
        if self.engine is not None and self.engine.B == b and not self._eval_engines:
            return self.engine
        ev = self._eval_engines.get(b)
        if ev is None:
            kw = dict(self._engine_kwargs)
            H, W = int(self.image_shape[1]), int(self.image_shape[2])
            ev = Engine(model=self._kind, height=H, width=W, batch=b, y_size=self.y_size or 30, tau=self.tau or 0.4,
                        precision=self.precision, **kw)
            self._eval_engines[b] = ev
        if self.engine is not None:
            ev.params.copy_(self.engine.params)
        elif self._pending_params is not None:
            ev.load_params(self._pending_params)
        ev.params_updated()
        return ev

    # -- checkpoints (vae/trainer.py:421 model.save_weights; Keras variable names and layouts) ----
    def save_weights(self, path, include_optimizer=False):
        """The reference writes Keras HDF5 (`model.save_weights('models/<run>.h5')`); h5py is not part of this image, so the
        same variables (Keras names, conv HWIO / dense [in,out] layouts) go into a NumPy `.npz`.  With include_optimizer the
        Adam moments and the iteration counter are stored too (the reference cannot resume; this build can)."""
        import numpy as np
        e = self.engine
        blob = {name: a for name, a in e.get_params().items()}
        if include_optimizer:
            for name, a in e._export(e.adam_m).items():
                blob["adam_m/" + name] = a
            for name, a in e._export(e.adam_v).items():
                blob["adam_v/" + name] = a
            blob["optimizer/iterations"] = np.asarray(e.iterations, dtype=np.int64)
        if not str(path).endswith(".npz"):
            path = str(path) + ".npz"
        np.savez(path, **blob)
        return path

    def load_weights(self, path):
        import numpy as np
        with np.load(path) as z:
            blob = {k: z[k] for k in z.files}
        named = {k: v for k, v in blob.items() if "/" not in k}
        self.set_weights_by_name(named)
        if self.engine is not None and "optimizer/iterations" in blob:
            e = self.engine
            for arena, prefix in ((e.adam_m, "adam_m/"), (e.adam_v, "adam_v/")):
                host = arena.cpu()
                for name, shape, off, cnt in e.table:
                    host[off:off + cnt] = torch.from_numpy(np.asarray(blob[prefix + name], dtype=np.float32).reshape(-1))
                arena.copy_(host)
            e.iterations = int(blob["optimizer/iterations"])

    def set_weights_by_name(self, named):
        self._pending_params = named
        if self.engine is not None:
            self.engine.load_params(named)

    def get_weights_by_name(self):
        return self.engine.get_params()

    @property
    def trainable_variables(self):
        return list(self.engine.get_params().items())

    # -- reference API -----------------------------------------------------------------------
    def _split_outputs(self):
        e = self.engine
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
        x_mean, _, x_hat_mean, _ = self._split_outputs()
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
        x_mean, x_ls, xh_mean, xh_ls = self._split_outputs()
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
        x_mean, x_ls, xh_mean, xh_ls = self._split_outputs()
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
