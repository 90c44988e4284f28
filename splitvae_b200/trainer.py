"""Mirror of the reference's trainer surface (vae/trainer.py) for the train hot path.

  kl_divergence / kl_divergence_two_gauss / discretised_logistic_loss   vae/trainer.py:11-38
  train_step_lg_vae / train_step_lg_gm_vae                              vae/trainer.py:120-173
  train_local_global_autoencoder                                        vae/trainer.py:72, 305-311, 417-419

The step itself (forward, fused loss fwd+bwd, backward, Adam) runs inside libsplitvae; this module
adds the CUDA-graph capture of the whole step and the data-parallel gradient all-reduce.
Evaluation, classifier scoring and visualisation (vae/trainer.py:313-416) are out of scope.
"""
from __future__ import annotations

import ctypes as C
import time

import torch

from . import _lib
from .model import LGGMVae, LGVae
from .parallel import BucketReducer, mean_scalars


# ---- module-level loss functions (same names / argument meaning as the reference) -------------
def kl_divergence(z_mean, z_sig):
    """vae/trainer.py:11-15 (tiny; plain device ops - the train step uses the fused kernel instead)."""
    z_log_var = torch.log(torch.square(z_sig))
    return torch.mean(-0.5 * torch.sum(1 + z_log_var - torch.square(z_mean) - torch.exp(z_log_var), dim=1))


def kl_divergence_two_gauss(mean1, sig1, mean2, sig2):
    """vae/trainer.py:17-18."""
    mean2 = torch.as_tensor(mean2, dtype=mean1.dtype, device=mean1.device)
    sig2 = torch.as_tensor(sig2, dtype=mean1.dtype, device=mean1.device)
    return torch.mean(torch.sum(torch.log(sig2) - torch.log(sig1)
                                + (torch.square(sig1) + torch.square(mean1 - mean2)) / (2 * torch.square(sig2)) - 0.5, dim=1))


def discretised_logistic_loss(x, m, log_scales):
    """vae/trainer.py:21-38, elementwise NLL, computed by libsplitvae's kernel."""
    lib = _lib.load()
    if not (x.is_cuda and m.is_cuda and log_scales.is_cuda):
        raise _lib.SplitVaeError("discretised_logistic_loss needs CUDA tensors: there is no CPU fallback")
    x, m, ls = (t.contiguous().float() for t in torch.broadcast_tensors(x, m, log_scales))
    out = torch.empty_like(x)
    _lib.check(lib.sv_discretised_logistic_loss(C.c_void_p(x.data_ptr()), C.c_void_p(m.data_ptr()), C.c_void_p(ls.data_ptr()),
                                                C.c_void_p(out.data_ptr()), x.numel(),
                                                C.c_void_p(torch.cuda.current_stream().cuda_stream)), None,
               "sv_discretised_logistic_loss")
    return out


# ---- optimizer objects (vae/main.py:65-68) ----------------------------------------------------
class ExponentialDecay:
    """tf.optimizers.schedules.ExponentialDecay(lr, 1e6, 0.4, staircase=True) - the only schedule the reference uses."""

    def __init__(self, initial_learning_rate, decay_steps=1000000, decay_rate=0.4, staircase=True):
        if decay_steps != 1000000 or decay_rate != 0.4 or not staircase:
            raise NotImplementedError("libsplitvae implements the reference's schedule (1e6, 0.4, staircase) only")
        self.initial_learning_rate = float(initial_learning_rate)


class Adam:
    """tf.keras.optimizers.Adam(learning_rate) with TF 2.0 defaults; the update runs in libsplitvae."""

    def __init__(self, learning_rate=1e-4):
        self.schedule = learning_rate if isinstance(learning_rate, ExponentialDecay) else None
        self.learning_rate = self.schedule.initial_learning_rate if self.schedule else float(learning_rate)


# ---- one train step, optionally graph-captured and data-parallel -------------------------------
class StepRunner:
    """Runs train_step_* on one GPU.  With use_graph the whole step (forward, loss, backward, bucketed
    all-reduce, Adam) is captured once into a CUDA graph and replayed; inputs/noise are staged into
    static device buffers before each replay."""

    def __init__(self, engine, use_graph=True, group=None, explicit_noise=False):
        self.e = engine
        self.group = group
        self.reducer = BucketReducer(engine.grads, engine.segments, group)
        self.use_graph = use_graph
        self.graph = None
        dev = engine.device
        B = engine.B
        self.inputs = torch.zeros(B, engine.H, engine.W, 6, dtype=torch.float32, device=dev)
        self.explicit_noise = explicit_noise
        self.eps_g = torch.zeros(B, 128, device=dev) if explicit_noise else None
        self.eps_l = torch.zeros(B, 128, device=dev) if explicit_noise else None
        self.u = torch.full((B, engine.y_size), 0.5, device=dev) if explicit_noise and engine.model == "lggmvae" else None

    def _issue(self):
        e = self.e
        if not self.reducer.enabled:
            e.train_step(self.inputs, self.eps_g, self.eps_l, self.u)
            return
        e.forward(self.inputs, self.eps_g, self.eps_l, self.u)
        e.loss_fwd_bwd(self.inputs)
        for s in range(len(e.segments)):
            e.backward_segment(s)
            self.reducer.reduce(s)      # all-reduce of this bucket overlaps the next segment
        self.reducer.wait_all()
        e.adam_step()

    def capture(self, warmup=2):
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            it0 = self.e.iterations
            snap = [t.clone() for t in (self.e.params, self.e.adam_m, self.e.adam_v)]
            for _ in range(warmup):
                self._issue()
            torch.cuda.synchronize()
            for t, s in zip((self.e.params, self.e.adam_m, self.e.adam_v), snap):
                t.copy_(s)
            self.e.params_updated()
            self.e.iterations = it0
        torch.cuda.current_stream().wait_stream(side)
        torch.cuda.synchronize()
        self.graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self.graph):
            self._issue()
        # capture does not execute: state is untouched

    def step(self, inputs=None, eps_g=None, eps_l=None, u=None):
        if inputs is not None:
            self.inputs.copy_(inputs, non_blocking=True)
        if self.explicit_noise:
            if eps_g is not None: self.eps_g.copy_(eps_g, non_blocking=True)
            if eps_l is not None: self.eps_l.copy_(eps_l, non_blocking=True)
            if u is not None and self.u is not None: self.u.copy_(u, non_blocking=True)
        if self.use_graph:
            if self.graph is None:
                self.capture()
            self.graph.replay()
        else:
            self._issue()

    def scalars(self):
        s = self.e.output("scalars")[:6]
        s = mean_scalars(s, self.group).cpu().tolist()
        names = ["recon_x", "recon_x_hat", "kl_x", "kl_x_hat", "y_kl" if self.e.model == "lggmvae" else "total_kl", "total"]
        return dict(zip(names, s))


_RUNNERS = {}


def _runner_for(model, images, optimizer, config=None):
    key = id(model)
    r = _RUNNERS.get(key)
    if r is None or r.e is not model.engine or model.engine.B != images.shape[0]:
        kw = {}
        if config is not None:
            kw["beta"] = float(config.get("beta", 40.0))
            kw["alpha"] = float(config.get("alpha", 40.0) or 40.0)
        if optimizer is not None:
            kw["learning_rate"] = optimizer.learning_rate
        model.configure(**kw)
        model.build(images.shape[0])
        r = StepRunner(model.engine, use_graph=bool(config.get("use_graph", True)) if config is not None else True)
        _RUNNERS[key] = r
    return r


def train_step_lg_vae(model, images, optimizer, config=None):
    """vae/trainer.py:120-144."""
    _runner_for(model, images, optimizer, config).step(images)


def train_step_lg_gm_vae(model, images, optimizer, config=None):
    """vae/trainer.py:146-173."""
    _runner_for(model, images, optimizer, config).step(images)


def train_local_global_autoencoder(model, optimizer, dataset, train_dataset, test_dataset, config):
    """Train loop of vae/trainer.py:72 (hot loop 305-311, stop 417-419).  `train_dataset` yields
    [B,H,W,6] float32 batches (or (images, labels) when config.label).  Every `report_every` steps the
    running means of the step scalars are printed (the reference prints them from its test loop
    every 10 000 steps, trainer.py:354-382; evaluation itself is out of scope)."""
    if isinstance(model, LGVae):
        train_step = train_step_lg_vae
    elif isinstance(model, LGGMVae):
        train_step = train_step_lg_gm_vae
    else:
        raise NotImplementedError(type(model).__name__)
    report_every = int(config.get("report_every", 10000) or 10000)
    start = time.time()
    history = []
    for step, train_data in enumerate(train_dataset):
        images = train_data[0] if config.get("label") else train_data
        if not images.is_cuda:
            images = images.cuda(non_blocking=True)
        train_step(model, images, optimizer, config)
        if step % report_every == 0:
            sc = _RUNNERS[id(model)].scalars()
            history.append((step, sc))
            print("Training time: {:.2f}".format(time.time() - start))
            print("step {}: ".format(step) + ", ".join("{}: {:.4f}".format(k, v) for k, v in sc.items()), flush=True)
            start = time.time()
        if step >= int(config.get("training_steps")):  # vae/trainer.py:417-419
            print('Training done!')
            break
    return history
