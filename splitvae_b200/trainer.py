"""Mirror of the reference's trainer surface (vae/trainer.py) for the train hot path.

  kl_divergence / kl_divergence_two_gauss / discretised_logistic_loss   vae/trainer.py:11-38
  train_step_lg_vae / train_step_lg_gm_vae                              vae/trainer.py:120-173
  train_local_global_autoencoder                                        vae/trainer.py:72, 305-311, 417-419

The step itself (forward, fused loss fwd+bwd, backward, Adam) runs inside libsplitvae; this module
adds the CUDA-graph capture of the whole step and the data-parallel gradient all-reduce.
The evaluation steps (test_step_lg_vae / test_step_lg_gm_vae, vae/trainer.py:199-274) and the periodic report
(313-382) reuse the forward and the fused loss kernel without the backward pass; classifier scoring, cluster accuracy
and visualisation need SVHN labels / the missing classifier blob and stay out of scope.
"""
from __future__ import annotations

import ctypes as C
import time

import torch

from . import _lib
from .model import GMVae, LGGMVae, LGVae
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

    def __init__(self, engine, use_graph=True, group=None, explicit_noise=False, nvls=None, write_reduced_grads=False):
        self.e = engine
        self.group = group
        self.reducer = BucketReducer(engine.grads, engine.segments, group)
        # nvls: parallel.NvlsArenas whose alloc() backs engine.params / engine.grads -> the fused in-switch reduce + Adam + broadcast
        # kernel replaces the NCCL all-reduce and the per-rank full-arena Adam
        self.nvls = nvls
        self.write_reduced_grads = write_reduced_grads
        self.use_graph = use_graph
        self.graph = None
        self.opt_stream = None
        dev = engine.device
        B = engine.B
        self.inputs = torch.zeros(B, engine.H, engine.W, 6, dtype=torch.float32, device=dev)
        self.explicit_noise = explicit_noise
        self.eps_g = torch.zeros(B, 128, device=dev) if explicit_noise else None
        self.eps_l = torch.zeros(B, 128, device=dev) if explicit_noise else None
        self.u = torch.full((B, engine.y_size), 0.5, device=dev) if explicit_noise and engine.model != "lgvae" else None

    def _issue(self):
        """One train step.  Data parallel: the gradients of backward segment k are final - decoders, then the encoders' last layers,
        then their first convs - so bucket k is all-reduced while segment k+1 runs, and its Adam update + operand re-pack follow on
        an optimizer stream as soon as the reduction lands.  Only the last (smallest) bucket and its update are exposed."""
        e = self.e
        if not self.reducer.enabled:
            e.train_step(self.inputs, self.eps_g, self.eps_l, self.u)
            return
        main = torch.cuda.current_stream()
        if self.opt_stream is None:
            # level with the engine's weight-gradient streams, below the capture stream (-3): at the default (lowest) priority the
            # first segment's reduction + update starved until the end of the step and the whole optimizer tail was exposed
            self.opt_stream = torch.cuda.Stream(priority=-2)
        opt = self.opt_stream
        e.forward(self.inputs, self.eps_g, self.eps_l, self.u)
        e.loss_fwd_bwd(self.inputs)
        nseg = len(e.segments)
        if self.nvls is not None:
            nv = self.nvls
            mc_g, mc_p = nv.multicast_ptr(e.grads), nv.multicast_ptr(e.params)
            for s in range(nseg):
                opt.wait_stream(main)
                e.backward_segment(s, done_stream=opt if s + 1 < nseg else None)
                if s + 1 == nseg:
                    opt.wait_stream(main)
                with torch.cuda.stream(opt):
                    nv.barrier(2 * s)                       # every rank's gradients of segment s are written
                    e.nvls_adam_segment(s, mc_g, mc_p, nv.rank, nv.world, self.write_reduced_grads)
                    nv.barrier(2 * s + 1)                   # every shard of the new weights has landed in this rank's arena
                    e.repack_segment(s)
            main.wait_stream(opt)
            return
        for s in range(nseg):
            if s + 1 < nseg:
                opt.wait_stream(main)               # (joins the optimizer stream into the capture)
                e.backward_segment(s, done_stream=opt)      # gradients final on opt; main goes on with segment s+1's chain
                with torch.cuda.stream(opt):
                    works = self.reducer.reduce(s)  # behind the optimizer stream on NCCL's stream
                    for w in works:
                        w.wait()                    # the optimizer stream waits for the reduction, the main stream does not
                    e.adam_segment(s)
            else:
                e.backward_segment(s)
                works = self.reducer.reduce(s)
                for w in works:
                    w.wait()
                main.wait_stream(opt)
                e.adam_segment(s)
        self.reducer._work = []

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
        # captured on a highest-priority stream: the engine's side stream matches it, its weight-gradient streams sit one level
        # below and its optimizer stream at the bottom (kernel nodes keep the priority of the stream they were captured on)
        with torch.cuda.graph(self.graph, stream=torch.cuda.Stream(priority=-3)):
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

    def metric_means(self, reset=True):
        """Means of the step's loss terms over every step since the last reset (accumulated on the device inside the step), averaged
        over the data-parallel ranks, and the number of steps."""
        d, n = self.e.metric_means(reset)
        if self.reducer.enabled:
            names = list(d)
            v = mean_scalars(torch.tensor([d[k] for k in names], device=self.e.device, dtype=torch.float32), self.group).cpu().tolist()
            d = dict(zip(names, v))
        return d, n

    def scalars(self):
        s = self.e.output("scalars")[:6]
        s = mean_scalars(s, self.group).cpu().tolist()
        names = ["recon_x", "recon_x_hat", "kl_x", "kl_x_hat", "y_kl" if self.e.model != "lgvae" else "total_kl", "total"]
        d = dict(zip(names, s))
        if self.e.model == "gmvae":
            d.pop("recon_x_hat"), d.pop("kl_x_hat")
        return d


def linear_assignment(labels, pred):
    """vae/trainer.py:40-67: majority-vote mapping of clusters to classes.  labels [n, num_class] one-hot, pred [n, num_cluster]
    scores; every sample of cluster i is assigned the most frequent true class inside that cluster (ties: the class met first,
    as tf.unique_with_counts + tf.argmax do); returns the assignment one-hot [n, num_class].  Host-side metric code: runs on
    whatever device the tensors live on (it is not part of the train hot path)."""
    num_class, num_cluster = labels.shape[1], pred.shape[1]
    lab = torch.argmax(labels, dim=1)
    cluster = torch.argmax(pred, dim=1)
    cluster_pred = torch.zeros_like(lab)
    for i in range(num_cluster):
        members = lab[cluster == i]
        if members.numel() == 0:           # skip if the cluster does not exist
            continue
        seen, counts = [], {}
        for v in members.tolist():         # unique values in order of first occurrence
            if v not in counts:
                seen.append(v)
                counts[v] = 0
            counts[v] += 1
        best = max(range(len(seen)), key=lambda j: (counts[seen[j]], -j))
        cluster_pred = torch.where(cluster == i, torch.full_like(cluster_pred, seen[best]), cluster_pred)
    return torch.nn.functional.one_hot(cluster_pred, num_class).to(labels.dtype)


class CategoricalAccuracy:
    """tf.keras.metrics.CategoricalAccuracy (vae/trainer.py:114-118): running fraction of rows whose argmax matches."""

    def __init__(self, name=None):
        self.name = name
        self.correct, self.count = 0, 0

    def __call__(self, y_true, y_pred):
        self.correct += int((torch.argmax(y_true, dim=1) == torch.argmax(y_pred, dim=1)).sum())
        self.count += int(y_true.shape[0])

    update_state = __call__

    def result(self):
        return self.correct / self.count if self.count else 0.0

    def reset_states(self):
        self.correct, self.count = 0, 0


class Mean:
    """tf.keras.metrics.Mean: running mean with result() / reset_states() (vae/trainer.py:99-113).  result() of an empty
    metric is 0, as in Keras."""

    def __init__(self, name=None):
        self.name = name
        self.total, self.count = 0.0, 0

    def __call__(self, value, count=1):
        """`count` > 1: `value` is already the mean of that many samples (device-side running sums read at report time)."""
        self.total += float(value) * count
        self.count += count

    update_state = __call__

    def result(self):
        return self.total / self.count if self.count else 0.0

    def reset_states(self):
        self.total, self.count = 0.0, 0


METRIC_NAMES = ("x_recon", "x_kl", "x_hat_recon", "x_hat_kl", "total_kl", "y_kl")


def make_metrics():
    """The train/test Keras Mean metrics of vae/trainer.py:99-113 (classifier accuracies stay at Keras' empty value 0)."""
    return {f"{n}_{split}_loss": Mean(f"{n}_{split}_loss") for split in ("train", "test") for n in METRIC_NAMES}


def _update_metrics(metrics, split, sc, gm, count=1):
    metrics[f"x_recon_{split}_loss"](sc["recon_x"], count)
    metrics[f"x_kl_{split}_loss"](sc["kl_x"], count)
    if "recon_x_hat" in sc:                 # (plain GMVAE updates x_recon / x_kl / y_kl only, vae/trainer.py:193-195)
        metrics[f"x_hat_recon_{split}_loss"](sc["recon_x_hat"], count)
        metrics[f"x_hat_kl_{split}_loss"](sc["kl_x_hat"], count)
    if gm:
        metrics[f"y_kl_{split}_loss"](sc["y_kl"], count)
    else:
        metrics[f"total_kl_{split}_loss"](sc["total_kl"], count)


def _test_step(model, images, config=None, eps_g=None, eps_l=None, u=None):
    """Forward + loss terms, no backward, no parameter update.  Runs sv_forward and the fused loss kernel (whose gradient
    outputs land in workspace buffers nobody reads here).  Returns (scalars dict, engine)."""
    kw = {}
    if config is not None:
        kw["beta"] = float(config.get("beta", 40.0))
        kw["alpha"] = float(config.get("alpha", 40.0) or 40.0)
        model.configure(**kw)
    if model.engine is None and not model._eval_engines:
        model.build(images.shape[0])
    e = model.eval_engine(images.shape[0])
    images = images.contiguous().float()
    e.forward(images, eps_g, eps_l, u)
    e.loss_fwd_bwd(images)
    return e.scalars(), e


def test_step_lg_vae(model, images, labels=None, config=None, metrics=None, eps_g=None, eps_l=None):
    """vae/trainer.py:199-232 without the classifier block (labels are accepted and ignored: the SVHN classifier weights
    are not part of the reference checkout).  Returns {recon_x, recon_x_hat, kl_x, kl_x_hat, total_kl, total}."""
    sc, _ = _test_step(model, images, config, eps_g, eps_l)
    if metrics is not None:
        _update_metrics(metrics, "test", sc, gm=False)
    return sc


def test_step_lg_gm_vae(model, images, labels=None, config=None, metrics=None, eps_g=None, eps_l=None, u=None):
    """vae/trainer.py:235-274: updates the test metrics and returns the model's 14-tuple like the reference."""
    sc, e = _test_step(model, images, config, eps_g, eps_l, u)
    if metrics is not None:
        _update_metrics(metrics, "test", sc, gm=True)
    o = e.output
    dx, dxh = o("dec_x"), o("dec_x_hat")
    test_step_lg_gm_vae.last_scalars = sc
    return (dx[..., :3], dx[..., 3:], o("z_x"), o("z_mean_x"), o("z_sig_x"), o("z_x_hat"), dxh[..., :3], dxh[..., 3:],
            o("z_mean_x_hat"), o("z_sig_x_hat"), o("y"), o("y_logits"), o("z_prior_mean"), o("z_prior_sig"))


def test_step_gm_vae(model, images, labels=None, config=None, metrics=None, eps_g=None, u=None):
    """vae/trainer.py:276-292: updates x_recon / x_kl / y_kl test metrics and returns the model's 9-tuple."""
    sc, e = _test_step(model, images, config, eps_g, None, u)
    if metrics is not None:
        _update_metrics(metrics, "test", sc, gm=True)
    o = e.output
    dx = o("dec_x")
    test_step_gm_vae.last_scalars = sc
    return (dx[..., :3], dx[..., 3:], o("z_x"), o("z_mean_x"), o("z_sig_x"), o("y"), o("y_logits"), o("z_prior_mean"), o("z_prior_sig"))


test_step_lg_vae.__test__ = False       # not pytest tests, whatever their names
test_step_lg_gm_vae.__test__ = False
test_step_gm_vae.__test__ = False

REPORT_TEMPLATE = ('Training step {}\n'
                   '            X Recon Loss: {:.4f}, X KLD loss: {:.4f}, Total X loss: {:.4f} \n'
                   '            X hat Recon Loss: {:.4f}, X hat KLD loss: {:.4f}, Total X hat loss: {:.4f} \n'
                   '            Test X Recon Loss: {:.4f}, Test X KLD loss: {:.4f}, Test Total X loss: {:.4f} \n'
                   '            Test X hat Recon Loss: {:.4f}, Test X hat KLD loss: {:.4f}, Test Total X hat loss: {:.4f}\n'
                   '            Total KL train loss: {:.4f}, Total KL test loss: {:.4f}\n'
                   '            Classifier recon acc: {:.4f}, Classifier random z_g acc: {:.4f}, Classifier random z_l acc: {:.4f}\n'
                   '            Classifier cluster acc: {:.4f}\n'
                   '            Y KL train loss: {:.4f}, Y KL test loss: {:.4f}')


def format_report(step, m, cluster_acc=0.0):
    """The stdout report of vae/trainer.py:354-382 (same template, same argument order).  The three classifier accuracies need the
    SVHN classifier (weights blob missing upstream) and print as Keras' empty-metric value 0."""
    r = lambda k: m[k].result()
    return REPORT_TEMPLATE.format(
        step,
        r("x_recon_train_loss"), r("x_kl_train_loss"), r("x_recon_train_loss") + r("x_kl_train_loss"),
        r("x_hat_recon_train_loss"), r("x_hat_kl_train_loss"), r("x_hat_recon_train_loss") + r("x_hat_kl_train_loss"),
        r("x_recon_test_loss"), r("x_kl_test_loss"), r("x_recon_test_loss") + r("x_kl_test_loss"),
        r("x_hat_recon_test_loss"), r("x_hat_kl_test_loss"), r("x_hat_recon_test_loss") + r("x_hat_kl_test_loss"),
        r("total_kl_train_loss"), r("total_kl_test_loss"),
        0.0, 0.0, 0.0, float(cluster_acc),
        r("y_kl_train_loss"), r("y_kl_test_loss"))


def reset_reported(m):
    """Only the metrics the reference resets after a report (vae/trainer.py:405-414): x_hat_* and y_kl_* keep running."""
    for k in ("x_recon_train_loss", "x_kl_train_loss", "x_recon_test_loss", "x_kl_test_loss", "total_kl_test_loss",
              "total_kl_train_loss"):
        m[k].reset_states()


_RUNNERS = {}


def _runner_for(model, images, optimizer, config=None):
    key = id(model)
    r = _RUNNERS.get(key)
    if r is None or r.e is not model.engine or model.engine.B != images.shape[0]:
        kw = {}
        if config is not None:
            kw["beta"] = float(config.get("beta", 40.0))
            kw["alpha"] = float(config.get("alpha", 40.0) or 40.0)
            if config.get("seed") is not None:      # per-run, per-rank Philox stream of the in-kernel noise (the reference never seeds)
                import os
                kw["rng_stream"] = (int(config.get("seed")) * 8191 + int(os.environ.get("RANK", "0")) + 1) & 0x3FFFFF
        if optimizer is not None:
            kw["learning_rate"] = optimizer.learning_rate
        model.configure(**kw)
        model.build(images.shape[0])            # (keeps weights, Adam moments and the iteration count of a previous engine)
        model.engine.output("scalar_sums").zero_()
        r = StepRunner(model.engine, use_graph=bool(config.get("use_graph", True)) if config is not None else True)
        _RUNNERS[key] = r
    return r


def train_step_lg_vae(model, images, optimizer, config=None):
    """vae/trainer.py:120-144."""
    _runner_for(model, images, optimizer, config).step(images)


def train_step_lg_gm_vae(model, images, optimizer, config=None):
    """vae/trainer.py:146-173."""
    _runner_for(model, images, optimizer, config).step(images)


def train_step_gm_vae(model, images, optimizer, config=None):
    """vae/trainer.py:175-195: recon_x + beta * KL(q(z|x) || p(z|y)) + alpha * KL(q(y|x) || U)."""
    _runner_for(model, images, optimizer, config).step(images)


def train_local_global_autoencoder(model, optimizer, dataset, train_dataset, test_dataset, config):
    """Train loop of vae/trainer.py:72 (hot loop 305-311, stop 417-419).  `train_dataset` yields
    [B,H,W,6] float32 batches (or (images, labels) when config.label).  Every `report_every` steps the
    running means of the step scalars are printed (the reference prints them from its test loop
    every 10 000 steps, trainer.py:354-382; evaluation itself is out of scope)."""
    if isinstance(model, LGVae):                       # vae/trainer.py:294-302
        train_step, test_step = train_step_lg_vae, test_step_lg_vae
    elif isinstance(model, LGGMVae):
        train_step, test_step = train_step_lg_gm_vae, test_step_lg_gm_vae
    elif isinstance(model, GMVae):
        train_step, test_step = train_step_gm_vae, test_step_gm_vae
    else:
        raise NotImplementedError(type(model).__name__)
    gm = isinstance(model, (LGGMVae, GMVae))
    y_logits_index = 11 if isinstance(model, LGGMVae) else 6
    report_every = int(config.get("report_every", 10000) or 10000)
    metrics = make_metrics()
    start = time.time()
    history = []
    for step, train_data in enumerate(train_dataset):
        images = train_data[0] if config.get("label") else train_data
        if not images.is_cuda:
            images = images.cuda(non_blocking=True)
        train_step(model, images, optimizer, config)
        if step % report_every == 0:
            # the reference updates its Keras Mean train metrics every step inside the tf.function (trainer.py:140-144); here the
            # step accumulates the running sums on the device and the host reads + clears them at the report steps only
            runner = _RUNNERS[id(model)]
            means, n_steps = runner.metric_means(reset=True)
            _update_metrics(metrics, "train", means, gm, count=max(n_steps, 1))
            sc = runner.scalars()
            history.append((step, sc))
            print("Training time: {:.2f}".format(time.time() - start))
            start = time.time()
            cluster_acc = CategoricalAccuracy("classifier_cluster_acc")
            if test_dataset is not None:                      # evaluation pass, vae/trainer.py:316-352
                all_labels, all_pred = [], []
                for test_data in test_dataset:
                    test_images = test_data[0] if config.get("label") else test_data
                    if not test_images.is_cuda:
                        test_images = test_images.cuda(non_blocking=True)
                    outs = test_step(model, test_images, config=config, metrics=metrics)
                    if config.get("label") and gm:            # trainer.py:323-328: labels and y_logits of the whole test set
                        all_labels.append(test_data[1].detach().cpu())
                        all_pred.append(outs[y_logits_index].detach().cpu().clone())
                if all_labels:                                # trainer.py:345-349: majority-vote cluster accuracy
                    labels = torch.cat(all_labels)
                    cluster_acc(labels, linear_assignment(labels, torch.cat(all_pred)))
                print("Testing time: {:.2f}".format(time.time() - start))
            print(format_report(step, metrics, cluster_acc.result()), flush=True)
            reset_reported(metrics)
            start = time.time()
        if step >= int(config.get("training_steps")):  # vae/trainer.py:417-419
            print('Training done!')
            break
    return history
