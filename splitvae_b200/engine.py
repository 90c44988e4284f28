"""Host-side owner of one libsplitvae handle: device arenas (torch tensors used purely as device
memory), parameter import/export in Keras layout, and thin wrappers over the C-ABI calls.
PyTorch is plumbing here (allocation, streams, CUDA-graph capture, torch.distributed); all
arithmetic of the train step happens in libsplitvae.so."""
from __future__ import annotations

import ctypes as C
from collections import OrderedDict

import numpy as np
import torch

from . import _lib
from ._lib import SvConfig, SvParamDesc, check


def _ptr(t):
    return C.c_void_p(0 if t is None else t.data_ptr())


def _stream():
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


class Engine:
    """One model replica on one GPU (vae/main.py:63-74 builds the same objects in TF)."""

    def __init__(self, model="lgvae", height=32, width=32, batch=64, y_size=30, tau=0.4, beta=40.0, alpha=40.0,
                 learning_rate=1e-4, world_size=1, precision="bf16x3", device=None, rng_stream=0, no_tc=False,
                 plan_only=False, arena_alloc=None):
        self.lib = _lib.load()
        self.model, self.H, self.W, self.B = model, int(height), int(width), int(batch)
        self.y_size = int(y_size)
        self.plan_only = plan_only
        flags = (_lib.SV_FLAG_PLAN_ONLY if plan_only else 0) | (_lib.SV_FLAG_NO_TC if no_tc else 0) | (int(rng_stream) << 8)
        self.cfg = SvConfig(_lib.SV_MODEL[model], self.H, self.W, self.B, 128, 128, self.y_size, float(tau), float(beta),
                            float(alpha), float(learning_rate), int(world_size), _lib.SV_PRECISION[precision], flags)
        self.h = C.c_void_p()
        if not plan_only:
            if not torch.cuda.is_available():
                raise _lib.SplitVaeError("no CUDA device: splitvae_b200 has no CPU fallback")
            self.device = torch.device(device if device is not None else f"cuda:{torch.cuda.current_device()}")
            torch.cuda.set_device(self.device)
        check(self.lib.sv_create(C.byref(self.cfg), C.byref(self.h)), None, "sv_create")
        self.table = []
        d = SvParamDesc()
        for i in range(self.lib.sv_param_count(self.h)):
            check(self.lib.sv_param_describe(self.h, i, C.byref(d)), self.h, "sv_param_describe")
            self.table.append((d.name.decode(), tuple(d.shape[:d.ndim]), int(d.offset), int(d.count)))
        self.arena_floats = int(self.lib.sv_arena_floats(self.h))
        self.workspace_bytes = int(self.lib.sv_workspace_bytes(self.h))
        if plan_only:
            return
        dev = self.device
        # arena_alloc(n_floats) -> zeroed fp32 CUDA tensor: lets the data-parallel host place the parameter and gradient arenas in
        # symmetric (NVLS multicast) memory so the fused reduce-scatter + Adam + all-gather kernel can address every rank's copy
        alloc = arena_alloc or (lambda n: torch.zeros(n, dtype=torch.float32, device=dev))
        self.params = alloc(self.arena_floats)
        self.grads = alloc(self.arena_floats)
        self.adam_m = torch.zeros_like(self.params)
        self.adam_v = torch.zeros_like(self.params)
        self.workspace = torch.zeros(self.workspace_bytes + 1024, dtype=torch.uint8, device=dev)
        off = (-self.workspace.data_ptr()) % 1024
        self._ws = self.workspace[off:off + self.workspace_bytes]
        check(self.lib.sv_bind(self.h, _ptr(self.params), _ptr(self.grads), _ptr(self.adam_m), _ptr(self.adam_v),
                               _ptr(self._ws), self.workspace_bytes), self.h, "sv_bind")
        self._outs = {}
        self.segments = []        # per backward segment: its arena ranges [(offset, count), ...] (gradient buckets, in backward order)
        off64, cnt64 = C.c_int64(), C.c_int64()
        for s in range(self.lib.sv_num_segments(self.h)):
            ranges = []
            for i in range(self.lib.sv_segment_num_ranges(self.h, s)):
                check(self.lib.sv_segment_range(self.h, s, i, C.byref(off64), C.byref(cnt64)), self.h, "sv_segment_range")
                ranges.append((int(off64.value), int(cnt64.value)))
            self.segments.append(ranges)

    def __del__(self):
        try:
            if getattr(self, "h", None) and self.h.value:
                self.lib.sv_destroy(self.h)
                self.h = C.c_void_p()
        except Exception:
            pass

    # ---- parameters (Keras layout) ---------------------------------------------------------
    def init_params(self, seed=5):
        """Keras default initialisers: glorot_uniform kernels, zero biases, constant(1) for the two
        softplus heads of the gmvae encoder (vae/model.py:68,76)."""
        gen = torch.Generator(device="cpu").manual_seed(int(seed))
        host = torch.zeros(self.arena_floats, dtype=torch.float32)
        for name, shape, off, cnt in self.table:
            if name.endswith(".kernel"):
                if len(shape) == 4:
                    fan_in, fan_out = shape[0] * shape[1] * shape[2], shape[0] * shape[1] * shape[3]
                else:
                    fan_in, fan_out = shape
                limit = (6.0 / (fan_in + fan_out)) ** 0.5
                host[off:off + cnt] = (torch.rand(cnt, generator=gen) * 2 - 1) * limit
            elif name.endswith("z_prior_sig.bias") or name.endswith("z_sig.bias"):
                host[off:off + cnt] = 1.0
        self.params.copy_(host)
        self.params_updated()

    def load_params(self, named):
        """named: mapping variable name -> array in Keras layout."""
        host = self.params.cpu()
        for name, shape, off, cnt in self.table:
            a = np.asarray(named[name], dtype=np.float32)
            if tuple(a.shape) != tuple(shape):
                raise ValueError(f"{name}: expected shape {shape}, got {a.shape}")
            host[off:off + cnt] = torch.from_numpy(a.reshape(-1))
        self.params.copy_(host)
        self.params_updated()

    def _export(self, arena):
        host = arena.detach().cpu().numpy()
        return OrderedDict((name, host[off:off + cnt].reshape(shape).copy()) for name, shape, off, cnt in self.table)

    def get_params(self):
        return self._export(self.params)

    def get_grads(self):
        return self._export(self.grads)

    def params_updated(self):
        check(self.lib.sv_params_updated(self.h, _stream()), self.h, "sv_params_updated")

    # ---- hot path ---------------------------------------------------------------------------
    def forward(self, inputs, eps_g=None, eps_l=None, u=None):
        self._check_inputs(inputs)
        check(self.lib.sv_forward(self.h, _ptr(inputs), _ptr(eps_g), _ptr(eps_l), _ptr(u), _stream()), self.h, "sv_forward")

    def loss_fwd_bwd(self, inputs):
        check(self.lib.sv_loss_fwd_bwd(self.h, _ptr(inputs), _stream()), self.h, "sv_loss_fwd_bwd")

    def backward_segment(self, seg, done_stream=None):
        """Backward pass of one segment on the current stream.  done_stream (a torch stream already ordered behind the current one):
        the segment's gradients are final THERE and the current stream only carries the activation-gradient chain
        (sv_backward_segment_deferred); not for the last segment."""
        if done_stream is None:
            check(self.lib.sv_backward_segment(self.h, seg, _stream()), self.h, "sv_backward_segment")
        else:
            check(self.lib.sv_backward_segment_deferred(self.h, seg, _stream(), C.c_void_p(done_stream.cuda_stream)), self.h,
                  "sv_backward_segment_deferred")

    def adam_step(self):
        check(self.lib.sv_adam_step(self.h, _stream()), self.h, "sv_adam_step")

    def nvls_adam_segment(self, seg, mc_grads_ptr, mc_params_ptr, rank, world, write_reduced_grads=False):
        """sv_nvls_adam_segment: in-switch gradient reduction + Adam on this rank's shard + multicast of the new weights."""
        check(self.lib.sv_nvls_adam_segment(self.h, seg, C.c_void_p(mc_grads_ptr), C.c_void_p(mc_params_ptr), rank, world,
                                            1 if write_reduced_grads else 0, _stream()), self.h, "sv_nvls_adam_segment")

    def repack_segment(self, seg):
        check(self.lib.sv_repack_segment(self.h, seg, _stream()), self.h, "sv_repack_segment")

    def adam_segment(self, seg):
        """Adam + operand re-pack of one backward segment (in order 0..n-1, once per step each), on the current stream."""
        check(self.lib.sv_adam_segment(self.h, seg, _stream()), self.h, "sv_adam_segment")

    def train_step(self, inputs, eps_g=None, eps_l=None, u=None):
        self._check_inputs(inputs)
        check(self.lib.sv_train_step(self.h, _ptr(inputs), _ptr(eps_g), _ptr(eps_l), _ptr(u), _stream()), self.h, "sv_train_step")

    def capture_graph(self, inputs, eps_g=None, eps_l=None, u=None):
        """Records one train step into the library's own CUDA graph (sv_capture_graph) on the current (non-default) stream; the
        tensors are baked in by address: refill them in place between replays."""
        self._check_inputs(inputs)
        self._graph_refs = (inputs, eps_g, eps_l, u)
        check(self.lib.sv_capture_graph(self.h, _ptr(inputs), _ptr(eps_g), _ptr(eps_l), _ptr(u), _stream()), self.h, "sv_capture_graph")

    def replay(self, n_steps=1):
        check(self.lib.sv_replay(self.h, int(n_steps), _stream()), self.h, "sv_replay")

    def decode(self, z_x, z_x_hat):
        check(self.lib.sv_decode(self.h, _ptr(z_x), _ptr(z_x_hat), _stream()), self.h, "sv_decode")

    def encode_y(self, y):
        check(self.lib.sv_encode_y(self.h, _ptr(y), _stream()), self.h, "sv_encode_y")

    def _check_inputs(self, inputs):
        if inputs.dtype != torch.float32 or not inputs.is_cuda or not inputs.is_contiguous() or \
                tuple(inputs.shape) != (self.B, self.H, self.W, 6):
            raise ValueError(f"inputs must be a contiguous cuda float32 tensor of shape {(self.B, self.H, self.W, 6)}, "
                             f"got {tuple(inputs.shape)} {inputs.dtype} {inputs.device}")

    # ---- outputs ----------------------------------------------------------------------------
    def output(self, name):
        """Zero-copy torch view of a named device result (see SV_OUT_* in include/splitvae.h)."""
        if name in self._outs:
            return self._outs[name]
        which = _lib.OUT_NAMES.index(name)
        p, n = C.c_void_p(), C.c_int64()
        check(self.lib.sv_output_ptr(self.h, which, C.byref(p), C.byref(n)), self.h, "sv_output_ptr")
        base = self._ws.data_ptr()
        off = p.value - base
        t = self._ws[off:off + 4 * n.value].view(torch.float32)
        B = self.B
        if name in ("dec_x", "dec_x_hat"):
            t = t.view(B, self.H, self.W, 6)
        elif name in ("y", "y_logits"):
            t = t.view(B, 32)[:, :self.y_size]
        elif name not in ("scalars", "scalar_sums"):
            t = t.view(B, 128)
        self._outs[name] = t
        return t

    def scalars(self):
        """dict of the step's loss terms (vae/trainer.py:127-135, 153-164); synchronises."""
        s = self.output("scalars").cpu().tolist()
        names = ["recon_x", "recon_x_hat", "kl_x", "kl_x_hat", "y_kl" if self.model != "lgvae" else "total_kl", "total"]
        d = dict(zip(names, s[:6]))
        if self.model == "gmvae":          # one decoder, one Gaussian KL (vae/trainer.py:181-187)
            d.pop("recon_x_hat"), d.pop("kl_x_hat")
        return d

    def metric_means(self, reset=True):
        """Means of the loss terms over every loss pass since the last reset (the reference's Keras Mean metrics, accumulated on the
        device inside the step; vae/trainer.py:140-144) and the number of passes; synchronises."""
        sums = self.output("scalar_sums")
        host = sums.cpu().tolist()
        if reset:
            sums.zero_()
        n = host[8]
        names = ["recon_x", "recon_x_hat", "kl_x", "kl_x_hat", "y_kl" if self.model != "lgvae" else "total_kl", "total"]
        d = {k: (v / n if n else 0.0) for k, v in zip(names, host[:6])}
        if self.model == "gmvae":
            d.pop("recon_x_hat"), d.pop("kl_x_hat")
        return d, int(n)

    @property
    def iterations(self):
        v = C.c_int64()
        check(self.lib.sv_get_iterations(self.h, C.byref(v), _stream()), self.h, "sv_get_iterations")
        return int(v.value)

    @iterations.setter
    def iterations(self, v):
        check(self.lib.sv_set_iterations(self.h, int(v), _stream()), self.h, "sv_set_iterations")

    @property
    def launch_count(self):
        return int(self.lib.sv_launch_count(self.h))

    # ---- test hooks (sv_debug_*) ---------------------------------------------------------------
    def debug_layers(self):
        infos = []
        for i in range(self.lib.sv_debug_layer_count(self.h)):
            info = _lib.SvLayerInfo()
            check(self.lib.sv_debug_layer_info(self.h, i, C.byref(info)), self.h, "sv_debug_layer_info")
            infos.append(info)
        return infos

    def debug_view(self, ptr, elems, dt):
        """torch view of `elems` elements of a workspace buffer (dt: 0 fp32, 1 bf16)."""
        off = ptr - self._ws.data_ptr()
        nbytes = elems * (4 if dt == 0 else 2)
        return self._ws[off:off + nbytes].view(torch.float32 if dt == 0 else torch.bfloat16)

    def debug_pixel_loss(self, inputs):
        check(self.lib.sv_debug_pixel_loss(self.h, _ptr(inputs), _stream()), self.h, "sv_debug_pixel_loss")

    def debug_run_layer(self, index, pass_, impl, inputs=None):
        check(self.lib.sv_debug_run_layer(self.h, index, pass_, impl, _ptr(inputs), _stream()), self.h, "sv_debug_run_layer")
