"""CPU oracle for the SPLIT-VAE / SPLIT-GMVAE train step.

TEST INFRASTRUCTURE ONLY.  Nothing in the product path (``splitvae_b200``) may import this
module; only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s cpu_baseline /
``--impl reference`` legs do, and only as the checker / the timed CPU baseline.

PARITY PARTLY PINNED.  The reference (51616/split-vae) ships no tests, golden vectors, seeds or
fixtures for this path and its arithmetic lives in TensorFlow 2.0.0 /
Keras (requirements.txt:7), which is not installable in this image.  What IS pinned by the
reference itself: its own source, executed here against a stand-in for the few `tf` names it uses
(generating scripts committed, vectors under tests/golden/reference_*.json):
  * the loss functions and the categorical KL (vae/trainer.py:11-38, 160-161), verbatim, in float64
    (scripts/make_reference_loss_golden.py) - the oracle agrees to 1e-10;
  * the model wiring and the loss assembly: vae/model.py imported UNMODIFIED (LGVae / LGGMVae: layer
    graph, concat / slice order, activations, return-tuple order) and the forward + loss lines of
    train_step_lg_vae / train_step_lg_gm_vae (beta / alpha weighting), with noise drawn from a queue
    and weights injected by Keras variable name (scripts/make_reference_model_golden.py) - the oracle's
    model_forward / step_losses agree to 1e-10 on every output tensor and scalar;
  * the patch scramble: augmentation.py imported UNMODIFIED, Augmentator.scramble driven by an injected patch
    permutation (scripts/make_reference_scramble_golden.py) - oracle.scramble agrees bit for bit.
Pinned by a THIRD-PARTY implementation of TensorFlow's semantics (not the reference, not this repository):
  * Conv2D padding='SAME' (strides 1 and 2, even and odd kernels / image sizes) and tf.image.resize (bilinear,
    half-pixel centres; the decoder's x2 and the CelebA 178 -> 64 down-scale): OpenCV's TensorFlow-graph importer
    executing a hand-encoded TF GraphDef (scripts/make_opencv_primitive_golden.py ->
    tests/golden/opencv_tf_primitives.npz, tests/test_oracle_tf_primitives.py) - conv2d_same / resize2x agree to
    fp32 rounding (3e-5 / 1e-6) on the fixture and on a live random sweep where cv2 is importable.
What stays UNPINNED (restated from the published TF/Keras semantics listed below, checked only by
analytic known-answer tests and fp64 finite differences): activations, Keras Adam / ExponentialDecay - and
autodiff, for which a second, independent numpy backward of the loss block exists (tests/test_oracle.py).

What is restated (reference file:line):
  * Sampling                      vae/model.py:9-13
  * Encoder (conv)                vae/model.py:34-45, 100-114
  * Encoder (gmvae)               vae/model.py:48-79, 116-140
  * Decoder                       vae/model.py:145-169
  * LGVae / LGGMVae               vae/model.py:174-218, 221-275
  * kl_divergence                 vae/trainer.py:11-15
  * kl_divergence_two_gauss       vae/trainer.py:17-18
  * discretised_logistic_loss     vae/trainer.py:21-38
  * train_step_lg_vae             vae/trainer.py:120-144
  * train_step_lg_gm_vae          vae/trainer.py:146-173
  * optimizer / lr schedule       vae/main.py:63-73  (tf.keras.optimizers.Adam, TF 2.0 defaults)
  * scramble augmentation         augmentation.py:43-57

TF/Keras semantics encoded here (not visible in the reference source):
  * Conv2D padding='same': total = max((ceil(in/s)-1)*s + k - in, 0); before = total//2.
  * tf.image.resize default = bilinear, half-pixel centres, no antialias.
  * Keras Dense/Conv2D defaults: glorot_uniform kernel, zero bias; ELU alpha=1.
  * Keras Adam (ResourceApplyAdam): alpha = lr*sqrt(1-b2^t)/(1-b1^t); m += (g-m)(1-b1);
    v += (g*g-v)(1-b2); p -= alpha*m/(sqrt(v)+eps), eps=1e-7, t = iterations+1.
  * ExponentialDecay(lr, 1e6, 0.4, staircase=True) on the 0-based iteration count.
  * All Dropout layers are inactive in the reference train step (SURVEY.md section 5 note).
"""
from __future__ import annotations

import math
from collections import OrderedDict

import numpy as np
import torch
import torch.nn.functional as F

LOG_127_5 = float(np.log(127.5))


# --------------------------------------------------------------------------------------
# variable inventory (Keras layout: conv HWIO, dense [in, out], bias [out])
# --------------------------------------------------------------------------------------
def layer_table(model: str, H: int, W: int, global_latent: int = 128, local_latent: int = 128,
                y_size: int = 30):
    """Ordered list of (name, kind, shape_kernel, bias_init) following the attribute order of
    vae/model.py (Encoder.__init__ 34-79, Decoder.__init__ 152-156, LGVae 182-186, LGGMVae 230-234)."""
    F_ = ((H // 8) * W) // 8 * 128  # operator precedence of vae/model.py:152
    rows = []

    def conv_encoder(prefix, latent):
        rows.append((prefix + ".e1", "conv", (6, 6, 3, 32), 0.0))
        rows.append((prefix + ".e2", "conv", (6, 6, 32, 64), 0.0))
        rows.append((prefix + ".e3", "conv", (4, 4, 64, 128), 0.0))
        rows.append((prefix + ".e4_mean", "dense", (F_, latent), 0.0))
        rows.append((prefix + ".e4_sd", "dense", (F_, latent), 0.0))

    def gm_encoder(prefix, latent):
        rows.append((prefix + ".h_block.0", "conv", (6, 6, 3, 128), 0.0))
        rows.append((prefix + ".h_block.1", "conv", (6, 6, 128, 128), 0.0))
        rows.append((prefix + ".h_block.2", "conv", (4, 4, 128, 128), 0.0))
        rows.append((prefix + ".y_block.0", "dense", (F_, 1024), 0.0))
        rows.append((prefix + ".y_block.2", "dense", (1024, 128), 0.0))
        rows.append((prefix + ".y_dense", "dense", (128, y_size), 0.0))
        rows.append((prefix + ".h_top_dense", "dense", (y_size, 512), 0.0))
        rows.append((prefix + ".z_prior_mean", "dense", (y_size, latent), 0.0))
        rows.append((prefix + ".z_prior_sig", "dense", (y_size, latent), 1.0))  # model.py:68
        rows.append((prefix + ".e1", "dense", (F_, 512), 0.0))
        rows.append((prefix + ".z_mean", "dense", (512, latent), 0.0))
        rows.append((prefix + ".z_sig", "dense", (512, latent), 1.0))          # model.py:76

    def decoder(prefix, latent):
        rows.append((prefix + ".d1", "dense", (latent, F_), 0.0))
        rows.append((prefix + ".d2", "conv", (4, 4, 128, 128), 0.0))
        rows.append((prefix + ".d3", "conv", (4, 4, 128, 64), 0.0))
        rows.append((prefix + ".d4", "conv", (6, 6, 64, 32), 0.0))
        rows.append((prefix + ".d5", "conv", (6, 6, 32, 6), 0.0))

    if model == "gmvae":            # plain GMVAE (vae/model.py:277-286): gm encoder + ONE decoder fed by z only; not built on the
        gm_encoder("encoder_x", global_latent)      # device yet - the oracle is ahead of the product here
        decoder("decoder_x", global_latent)
        return rows
    if model == "lgvae":
        conv_encoder("encoder_x", global_latent)
    elif model == "lggmvae":
        gm_encoder("encoder_x", global_latent)
    else:
        raise NotImplementedError(model)
    conv_encoder("encoder_x_hat", local_latent)
    decoder("decoder_x", global_latent + local_latent)
    decoder("decoder_x_hat", local_latent)
    return rows


def init_params(model: str, H: int, W: int, seed: int = 5, y_size: int = 30,
                global_latent: int = 128, local_latent: int = 128, decoder_ls_bias=None):
    """Glorot-uniform kernels / constant biases in Keras layout, as float32 numpy arrays.
    ``decoder_ls_bias`` (optional) sets the log-scale bias of d5 (channels 3..5) to emulate a
    trained model so the narrow-scale branches of the likelihood are exercised."""
    rng = np.random.default_rng(seed)
    params = OrderedDict()
    for name, kind, shape, bias0 in layer_table(model, H, W, global_latent, local_latent, y_size):
        if kind == "conv":
            kh, kw, ci, co = shape
            fan_in, fan_out = kh * kw * ci, kh * kw * co
        else:
            fan_in, fan_out = shape
            co = shape[1]
        limit = math.sqrt(6.0 / (fan_in + fan_out))
        params[name + ".kernel"] = rng.uniform(-limit, limit, size=shape).astype(np.float32)
        b = np.full((co,), bias0, dtype=np.float32)
        if decoder_ls_bias is not None and name.endswith(".d5"):
            b[3:] = decoder_ls_bias
        params[name + ".bias"] = b
    return params


# --------------------------------------------------------------------------------------
# TF op restatements on torch tensors (NHWC at the interface)
# --------------------------------------------------------------------------------------
def same_pad(in_size: int, k: int, s: int):
    out = -(-in_size // s)
    total = max((out - 1) * s + k - in_size, 0)
    return total // 2, total - total // 2


def conv2d_same(x_nhwc, kernel_hwio, bias, stride):
    """Keras Conv2D(padding='same') on an NHWC tensor."""
    kh, kw = kernel_hwio.shape[0], kernel_hwio.shape[1]
    pt, pb = same_pad(x_nhwc.shape[1], kh, stride)
    pl, pr = same_pad(x_nhwc.shape[2], kw, stride)
    x = x_nhwc.permute(0, 3, 1, 2)
    x = F.pad(x, (pl, pr, pt, pb))
    w = kernel_hwio.permute(3, 2, 0, 1)
    y = F.conv2d(x, w, bias, stride=stride)
    return y.permute(0, 2, 3, 1)


def resize2x(x_nhwc):
    """tf.image.resize(x, [2H, 2W]) (vae/model.py:163-167): bilinear, half-pixel centres."""
    x = x_nhwc.permute(0, 3, 1, 2)
    y = F.interpolate(x, scale_factor=2, mode="bilinear", align_corners=False)
    return y.permute(0, 2, 3, 1)


def dense(x, kernel, bias):
    return x @ kernel + bias


def _act(x, name):
    if name == "relu":
        return torch.relu(x)
    if name == "elu":
        return F.elu(x)
    if name == "softplus":
        return F.softplus(x)
    return x


# --------------------------------------------------------------------------------------
# model forward (vae/model.py)
# --------------------------------------------------------------------------------------
def encoder_conv(P, pre, x, eps):
    """Encoder.call_conv, vae/model.py:100-114."""
    h = _act(conv2d_same(x, P[pre + ".e1.kernel"], P[pre + ".e1.bias"], 2), "relu")
    h = _act(conv2d_same(h, P[pre + ".e2.kernel"], P[pre + ".e2.bias"], 2), "relu")
    h = _act(conv2d_same(h, P[pre + ".e3.kernel"], P[pre + ".e3.bias"], 2), "relu")
    h = h.reshape(h.shape[0], -1)  # Flatten over NHWC
    z_mean = dense(h, P[pre + ".e4_mean.kernel"], P[pre + ".e4_mean.bias"])
    z_sig = _act(dense(h, P[pre + ".e4_sd.kernel"], P[pre + ".e4_sd.bias"]), "softplus")
    z = z_mean + z_sig * eps  # Sampling, vae/model.py:9-13
    return z, z_mean, z_sig


def encoder_gmvae(P, pre, x, eps, u, tau):
    """Encoder.call_gmvae, vae/model.py:116-135 (all dropout inactive)."""
    h = x
    for i in range(3):
        s = 2
        h = _act(conv2d_same(h, P[f"{pre}.h_block.{i}.kernel"], P[f"{pre}.h_block.{i}.bias"], s), "elu")
    h = h.reshape(h.shape[0], -1)
    yh = _act(dense(h, P[pre + ".y_block.0.kernel"], P[pre + ".y_block.0.bias"]), "elu")
    yh = _act(dense(yh, P[pre + ".y_block.2.kernel"], P[pre + ".y_block.2.bias"]), "elu")
    y_logits = dense(yh, P[pre + ".y_dense.kernel"], P[pre + ".y_dense.bias"])
    y = torch.softmax((y_logits - torch.log(-torch.log(u))) / tau, dim=1)  # model.py:123
    z_prior_mean = dense(y, P[pre + ".z_prior_mean.kernel"], P[pre + ".z_prior_mean.bias"])
    z_prior_sig = _act(dense(y, P[pre + ".z_prior_sig.kernel"], P[pre + ".z_prior_sig.bias"]), "softplus")
    h_top = _act(dense(y, P[pre + ".h_top_dense.kernel"], P[pre + ".h_top_dense.bias"]), "elu")
    h = _act(dense(h, P[pre + ".e1.kernel"], P[pre + ".e1.bias"]), "elu")
    h = h + h_top
    z_mean = dense(h, P[pre + ".z_mean.kernel"], P[pre + ".z_mean.bias"])
    z_sig = _act(dense(h, P[pre + ".z_sig.kernel"], P[pre + ".z_sig.bias"]), "softplus")
    z = z_mean + z_sig * eps
    return z, z_mean, z_sig, y, y_logits, z_prior_mean, z_prior_sig


def decoder(P, pre, z, H, W):
    """Decoder.call, vae/model.py:158-169."""
    h = _act(dense(z, P[pre + ".d1.kernel"], P[pre + ".d1.bias"]), "relu")
    h = h.reshape(-1, H // 8, W // 8, 128)
    h = _act(conv2d_same(h, P[pre + ".d2.kernel"], P[pre + ".d2.bias"], 1), "relu")
    h = resize2x(h)
    h = _act(conv2d_same(h, P[pre + ".d3.kernel"], P[pre + ".d3.bias"], 1), "relu")
    h = resize2x(h)
    h = _act(conv2d_same(h, P[pre + ".d4.kernel"], P[pre + ".d4.bias"], 1), "relu")
    h = resize2x(h)
    h = conv2d_same(h, P[pre + ".d5.kernel"], P[pre + ".d5.bias"], 1)
    return h[..., :3], h[..., 3:]


def model_forward(P, model, inputs, eps_g, eps_l, u=None, tau=0.4):
    """LGVae.call (model.py:189-200) / LGGMVae.call (model.py:237-248).  Returns a dict holding
    the reference's output tuple by name."""
    H, W = inputs.shape[1], inputs.shape[2]
    x, x_hat = inputs[..., :3], inputs[..., 3:]
    out = {}
    if model == "gmvae":            # GMVae.call, vae/model.py:288-299 (9-tuple)
        z_x, zm_x, zs_x, y, y_logits, zpm, zps = encoder_gmvae(P, "encoder_x", x, eps_g, u, tau)
        x_mean, x_ls = decoder(P, "decoder_x", z_x, H, W)
        return dict(x_mean=x_mean, x_log_scale=x_ls, z_x=z_x, z_mean_x=zm_x, z_sig_x=zs_x, y=y, y_logits=y_logits,
                    z_prior_mean=zpm, z_prior_sig=zps)
    if model == "lgvae":
        z_x, zm_x, zs_x = encoder_conv(P, "encoder_x", x, eps_g)
    else:
        z_x, zm_x, zs_x, y, y_logits, zpm, zps = encoder_gmvae(P, "encoder_x", x, eps_g, u, tau)
        out.update(y=y, y_logits=y_logits, z_prior_mean=zpm, z_prior_sig=zps)
    z_xh, zm_xh, zs_xh = encoder_conv(P, "encoder_x_hat", x_hat, eps_l)
    x_mean, x_ls = decoder(P, "decoder_x", torch.cat([z_x, z_xh], dim=1), H, W)
    xh_mean, xh_ls = decoder(P, "decoder_x_hat", z_xh, H, W)
    out.update(x_mean=x_mean, x_log_scale=x_ls, z_x=z_x, z_mean_x=zm_x, z_sig_x=zs_x,
               z_x_hat=z_xh, x_hat_mean=xh_mean, x_hat_log_scale=xh_ls,
               z_mean_x_hat=zm_xh, z_sig_x_hat=zs_xh)
    return out


# --------------------------------------------------------------------------------------
# losses (vae/trainer.py:11-38)
# --------------------------------------------------------------------------------------
def kl_divergence(z_mean, z_sig):
    z_log_var = torch.log(torch.square(z_sig))
    return torch.mean(-0.5 * torch.sum(1 + z_log_var - torch.square(z_mean) - torch.exp(z_log_var), dim=1))


def kl_divergence_two_gauss(mean1, sig1, mean2, sig2):
    if not torch.is_tensor(mean2):
        mean2 = torch.as_tensor(mean2, dtype=mean1.dtype)
    if not torch.is_tensor(sig2):
        sig2 = torch.as_tensor(sig2, dtype=mean1.dtype)
    return torch.mean(torch.sum(torch.log(sig2) - torch.log(sig1)
                                + (torch.square(sig1) + torch.square(mean1 - mean2)) / (2 * torch.square(sig2))
                                - 0.5, dim=1))


def discretised_logistic_loss(x, m, log_scales):
    centered_x = x - m
    inv_stdv = torch.exp(-log_scales)
    plus_in = inv_stdv * (centered_x + 1. / 255.)
    min_in = inv_stdv * (centered_x - 1. / 255.)
    cdf_plus = torch.sigmoid(plus_in)
    cdf_min = torch.sigmoid(min_in)
    cdf_delta = cdf_plus - cdf_min
    mid_in = inv_stdv * centered_x
    log_pdf_mid = mid_in - log_scales - 2. * F.softplus(mid_in)
    log_cdf_plus = plus_in - F.softplus(plus_in)
    log_one_minus_cdf_min = -F.softplus(min_in)
    tiny = torch.as_tensor(1e-12, dtype=x.dtype)
    log_prob = torch.where(
        x < -0.999, log_cdf_plus,
        torch.where(x > 0.999, log_one_minus_cdf_min,
                    torch.where(cdf_delta > 1e-5, torch.log(torch.maximum(cdf_delta, tiny)),
                                log_pdf_mid - LOG_127_5)))
    return -log_prob


def step_losses(out, inputs, model, beta, alpha=40.0, y_size=30):
    """Loss block of train_step_lg_vae (trainer.py:125-135) / train_step_lg_gm_vae (151-164)."""
    x, x_hat = inputs[..., :3], inputs[..., 3:]
    L = {}
    L["recon_x"] = torch.mean(torch.sum(discretised_logistic_loss(x, out["x_mean"], out["x_log_scale"]), dim=[1, 2, 3]))
    if model == "gmvae":            # train_step_gm_vae, vae/trainer.py:175-195
        L["kl_x"] = kl_divergence_two_gauss(out["z_mean_x"], out["z_sig_x"], out["z_prior_mean"], out["z_prior_sig"])
        py = torch.softmax(out["y_logits"], dim=1)
        L["y_kl"] = torch.mean(torch.sum(py * (torch.log(py + 1e-8) - math.log(1.0 / y_size)), dim=1))
        L["total"] = L["recon_x"] + beta * L["kl_x"] + alpha * L["y_kl"]
        return L
    L["recon_x_hat"] = torch.mean(torch.sum(discretised_logistic_loss(x_hat, out["x_hat_mean"], out["x_hat_log_scale"]), dim=[1, 2, 3]))
    if model == "lgvae":
        L["total_kl"] = beta * kl_divergence(torch.cat([out["z_mean_x"], out["z_mean_x_hat"]], dim=1),
                                             torch.cat([out["z_sig_x"], out["z_sig_x_hat"]], dim=1))
        L["kl_x"] = kl_divergence(out["z_mean_x"], out["z_sig_x"])
        L["kl_x_hat"] = kl_divergence(out["z_mean_x_hat"], out["z_sig_x_hat"])
        L["total"] = L["recon_x"] + L["recon_x_hat"] + L["total_kl"]
    else:
        L["kl_x"] = kl_divergence_two_gauss(out["z_mean_x"], out["z_sig_x"], out["z_prior_mean"], out["z_prior_sig"])
        L["kl_x_hat"] = kl_divergence_two_gauss(out["z_mean_x_hat"], out["z_sig_x_hat"], 0., 1.)
        py = torch.softmax(out["y_logits"], dim=1)
        L["y_kl"] = torch.mean(torch.sum(py * (torch.log(py + 1e-8) - math.log(1.0 / y_size)), dim=1))
        L["total"] = L["recon_x"] + L["recon_x_hat"] + beta * (L["kl_x"] + L["kl_x_hat"]) + alpha * L["y_kl"]
    return L


# --------------------------------------------------------------------------------------
# optimizer (vae/main.py:65-68; TF 2.0 Keras Adam / ResourceApplyAdam), fp32 numpy, fixed op order
# --------------------------------------------------------------------------------------
def lr_at(model, base_lr, iteration):
    if model == "lgvae":
        return float(base_lr)
    return float(base_lr) * (0.4 ** math.floor(iteration / 1000000.0))


def adam_alpha(lr, t, beta1=0.9, beta2=0.999):
    """Scalar step size for step t = iterations+1, computed in double then rounded to fp32."""
    return np.float32(lr * math.sqrt(1.0 - beta2 ** t) / (1.0 - beta1 ** t))


def keras_adam_update(p, g, m, v, alpha, beta1=0.9, beta2=0.999, eps=1e-7):
    """One ResourceApplyAdam on float32 arrays, each operation rounded to fp32 in this order."""
    f = np.float32
    one_minus_b1 = f(1.0) - f(beta1)
    one_minus_b2 = f(1.0) - f(beta2)
    m = (m + (g - m) * one_minus_b1).astype(np.float32)
    v = (v + (g * g - v) * one_minus_b2).astype(np.float32)
    p = (p - (m * f(alpha)) / (np.sqrt(v) + f(eps))).astype(np.float32)
    return p, m, v


# --------------------------------------------------------------------------------------
# whole train step
# --------------------------------------------------------------------------------------
def to_torch(params, dtype, requires_grad=True):
    P = OrderedDict()
    for k, a in params.items():
        t = torch.tensor(np.asarray(a), dtype=dtype)
        t.requires_grad_(requires_grad)
        P[k] = t
    return P


def forward_backward(params, model, inputs, eps_g, eps_l, u=None, *, beta, alpha=40.0, tau=0.4,
                     y_size=30, dtype=torch.float32, want_outputs=False):
    """Forward + losses + autograd gradients.  Returns (scalars dict, grads dict[, outputs])."""
    P = to_torch(params, dtype)
    tin = torch.tensor(np.asarray(inputs), dtype=dtype)
    te_g = torch.tensor(np.asarray(eps_g), dtype=dtype)
    te_l = torch.tensor(np.asarray(eps_l), dtype=dtype)
    tu = None if u is None else torch.tensor(np.asarray(u), dtype=dtype)
    out = model_forward(P, model, tin, te_g, te_l, tu, tau)
    L = step_losses(out, tin, model, beta, alpha, y_size)
    L["total"].backward()
    scalars = {k: float(v.detach()) for k, v in L.items()}
    grads = OrderedDict((k, (t.grad.detach().numpy().copy() if t.grad is not None
                             else np.zeros(tuple(t.shape), dtype=np.float64 if dtype == torch.float64 else np.float32)))
                        for k, t in P.items())
    if want_outputs:
        return scalars, grads, {k: v.detach().numpy() for k, v in out.items()}
    return scalars, grads


class TrainState:
    """Parameters + Adam slots + iteration counter, all float32 numpy in Keras layout."""

    def __init__(self, params):
        self.params = OrderedDict((k, np.asarray(v, dtype=np.float32).copy()) for k, v in params.items())
        self.m = OrderedDict((k, np.zeros_like(v)) for k, v in self.params.items())
        self.v = OrderedDict((k, np.zeros_like(v)) for k, v in self.params.items())
        self.iterations = 0


def train_step(state: TrainState, model, inputs, eps_g, eps_l, u=None, *, beta, alpha=40.0, tau=0.4,
               y_size=30, lr=1e-4, dtype=torch.float32):
    """train_step_lg_vae / train_step_lg_gm_vae: forward, loss, gradients, Adam.  Mutates state."""
    scalars, grads = forward_backward(state.params, model, inputs, eps_g, eps_l, u, beta=beta, alpha=alpha,
                                      tau=tau, y_size=y_size, dtype=dtype)
    t = state.iterations + 1
    a = adam_alpha(lr_at(model, lr, state.iterations), t)
    for k in state.params:
        g = grads[k].astype(np.float32)
        state.params[k], state.m[k], state.v[k] = keras_adam_update(state.params[k], g, state.m[k], state.v[k], a)
    state.iterations = t
    return scalars, grads


# --------------------------------------------------------------------------------------
# scramble augmentation (augmentation.py:43-57) and synthetic inputs
# --------------------------------------------------------------------------------------
def scramble(x_hwc, p, perm):
    """x_hat built from the p x p patches of x, permuted by ``perm`` (patch q of the output is patch
    perm[q] of the input, both numbered row-major), returns concat([x, x_hat], axis=2)."""
    Hh, Ww, C = x_hwc.shape
    G = Ww // p
    n_patch = (Hh // p) * G
    patches = x_hwc.reshape(Hh // p, p, G, p, C).transpose(0, 2, 1, 3, 4).reshape(n_patch, p, p, C)
    patches = patches[np.asarray(perm)]
    rows = [np.concatenate(list(patches[g * G:(g + 1) * G]), axis=1) for g in range(n_patch // G)]
    x_aug = np.concatenate(rows, axis=0)
    return np.concatenate([x_hwc, x_aug], axis=2)


def celeba_preprocess(u8_hwc, size=64, crop=178):
    """vae/data.py:82-87 on one decoded image [Hs,Ws,3] uint8: tf.image.resize_with_crop_or_pad(image, 178, 178) (centre crop, offset
    (Hs-178)//2 - zero padding when smaller is not needed for the 218x178 aligned CelebA files), tf.image.resize(image, [64, 64])
    (TF2 default: bilinear, half-pixel centres, NO antialiasing == F.interpolate(align_corners=False, antialias=False)), /255*2-1."""
    a = np.asarray(u8_hwc)
    Hs, Ws = a.shape[:2]
    cy, cx = (Hs - crop) // 2, (Ws - crop) // 2
    t = torch.tensor(a[cy:cy + crop, cx:cx + crop].astype(np.float64)).permute(2, 0, 1)[None]
    r = F.interpolate(t, size=(size, size), mode="bilinear", align_corners=False, antialias=False)[0].permute(1, 2, 0)
    return (r / 255.0 * 2 - 1).numpy().astype(np.float32)


def synthetic_batch(B, H, p, y_size=30, seed_base=0):
    """Inputs / noise of BASELINE.md section 5 (seeds 0..4 offset by seed_base)."""
    k = np.random.default_rng(seed_base + 0).integers(0, 256, size=(B, H, H, 3), dtype=np.uint8)
    x = (k / 255.0 * 2 - 1).astype(np.float32)  # vae/data.py:52
    rng_p = np.random.default_rng(seed_base + 1)
    n_patch = (H // p) * (H // p)
    perms = np.stack([rng_p.permutation(n_patch) for _ in range(B)]).astype(np.int32)
    inputs = np.stack([scramble(x[b], p, perms[b]) for b in range(B)]).astype(np.float32)
    eps_g = np.random.default_rng(seed_base + 2).standard_normal((B, 128)).astype(np.float32)
    eps_l = np.random.default_rng(seed_base + 3).standard_normal((B, 128)).astype(np.float32)
    u = np.random.default_rng(seed_base + 4).uniform(1e-6, 1 - 1e-6, size=(B, y_size)).astype(np.float32)
    return dict(u8=k, perms=perms, inputs=inputs, eps_g=eps_g, eps_l=eps_l, u=u)


# --------------------------------------------------------------------------------------
# second, independent restatement: hand-derived numpy forward+backward of the loss block
# (SURVEY.md 9.2).  Used to cross-check the autograd oracle and to check the fused CUDA kernel.
# --------------------------------------------------------------------------------------
def _sigmoid(a):
    with np.errstate(over="ignore"):
        return 1.0 / (1.0 + np.exp(-a))


def _softplus(a):
    return np.logaddexp(0.0, a)


def dll_fwd_bwd_numpy(x, m, ls):
    """Per-element discretised-logistic NLL and its derivatives w.r.t. m and ls (float64)."""
    x = np.asarray(x, np.float64); m = np.asarray(m, np.float64); ls = np.asarray(ls, np.float64)
    s = np.exp(-ls)
    c = x - m
    plus = s * (c + 1. / 255.)
    mn = s * (c - 1. / 255.)
    mid = s * c
    sp, sm = _sigmoid(plus), _sigmoid(mn)
    delta = sp - sm
    dsp, dsm = sp * (1 - sp), sm * (1 - sm)
    b1 = x < -0.999
    b2 = (~b1) & (x > 0.999)
    b3 = (~b1) & (~b2) & (delta > 1e-5)
    b4 = ~(b1 | b2 | b3)
    lp = np.where(b1, plus - _softplus(plus),
                  np.where(b2, -_softplus(mn),
                           np.where(b3, np.log(np.maximum(delta, 1e-12)),
                                    mid - ls - 2 * _softplus(mid) - LOG_127_5)))
    safe = np.where(b3, delta, 1.0)
    dm = np.where(b1, -s * _sigmoid(-plus),
                  np.where(b2, s * sm,
                           np.where(b3, -s * (dsp - dsm) / safe,
                                    -s * (1 - 2 * _sigmoid(mid)))))
    dls = np.where(b1, -plus * _sigmoid(-plus),
                   np.where(b2, mn * sm,
                            np.where(b3, -(plus * dsp - mn * dsm) / safe,
                                     -mid * (1 - 2 * _sigmoid(mid)) - 1.0)))
    branch = np.where(b1, 0, np.where(b2, 1, np.where(b3, 2, 3)))
    # loss = -log_prob
    return -lp, -dm, -dls, branch


def latent_fwd_bwd_numpy(model, zm_g, zs_g, zm_l, zs_l, beta, alpha=40.0, y_logits=None, zpm=None, zps=None):
    """KL terms and their gradients (already scaled by beta|alpha and 1/B), float64."""
    f = np.float64
    zm_g, zs_g, zm_l, zs_l = (np.asarray(a, f) for a in (zm_g, zs_g, zm_l, zs_l))
    B = zm_g.shape[0]
    out = {}
    if model == "lgvae":
        def kl(mu, sg):
            return np.mean(-0.5 * np.sum(1 + np.log(sg * sg) - mu * mu - sg * sg, axis=1))
        out["kl_x"], out["kl_x_hat"] = kl(zm_g, zs_g), kl(zm_l, zs_l)
        out["total_kl"] = beta * (out["kl_x"] + out["kl_x_hat"])
        out["d_zm_g"], out["d_zs_g"] = beta / B * zm_g, beta / B * (zs_g - 1 / zs_g)
        out["d_zm_l"], out["d_zs_l"] = beta / B * zm_l, beta / B * (zs_l - 1 / zs_l)
    else:
        zpm, zps, y_logits = np.asarray(zpm, f), np.asarray(zps, f), np.asarray(y_logits, f)
        def kl2(m1, s1, m2, s2):
            return np.mean(np.sum(np.log(s2) - np.log(s1) + (s1 * s1 + (m1 - m2) ** 2) / (2 * s2 * s2) - 0.5, axis=1))
        out["kl_x"] = kl2(zm_g, zs_g, zpm, zps)
        out["kl_x_hat"] = kl2(zm_l, zs_l, 0.0, 1.0)
        d = zm_g - zpm
        out["d_zm_g"] = beta / B * d / zps ** 2
        out["d_zs_g"] = beta / B * (zs_g / zps ** 2 - 1 / zs_g)
        out["d_zpm"] = -beta / B * d / zps ** 2
        out["d_zps"] = beta / B * (1 / zps - (zs_g ** 2 + d ** 2) / zps ** 3)
        out["d_zm_l"], out["d_zs_l"] = beta / B * zm_l, beta / B * (zs_l - 1 / zs_l)
        K = y_logits.shape[1]
        e = np.exp(y_logits - y_logits.max(axis=1, keepdims=True))
        py = e / e.sum(axis=1, keepdims=True)
        out["y_kl"] = np.mean(np.sum(py * (np.log(py + 1e-8) + np.log(K)), axis=1))
        dfdp = np.log(py + 1e-8) + np.log(K) + py / (py + 1e-8)
        out["d_y_logits"] = alpha / B * py * (dfdp - np.sum(py * dfdp, axis=1, keepdims=True))
    return out
