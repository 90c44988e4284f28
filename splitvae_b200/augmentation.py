"""Augmentator('scramble', size) of augmentation.py:12-57 on the device.

The reference scrambles one image at a time inside a tf.data map (8 CPU threads, vae/main.py:57-61);
here a whole uint8 batch is scaled to [-1,1] (vae/data.py:52) and scrambled by one kernel
(sv_stage_scramble), given one uniform patch permutation per image.  Only 'scramble' (the CLI
default and the only mode any README command uses) is implemented."""
from __future__ import annotations

import ctypes as C

import torch

from . import _lib


class Augmentator(object):
    def __init__(self, type, size=1, mean=0, std=1):
        if type != "scramble":
            raise NotImplementedError(f"augmentation '{type}' is outside the hot path of this build")
        self.size = int(size)
        self.augment = self.scramble

    def draw_permutations(self, batch, height, width, generator=None, device="cuda"):
        """One uniform permutation of the (H/p)*(W/p) patches per image (tf.random.shuffle, augmentation.py:49)."""
        n_patch = (height // self.size) * (width // self.size)
        keys = torch.rand(batch, n_patch, generator=generator, device=device)
        return torch.argsort(keys, dim=1).to(torch.int32)

    def scramble(self, u8_batch, perms=None, out=None):
        """u8_batch: [B,H,W,3] uint8 cuda tensor -> [B,H,W,6] float32 (x | x_hat), augmentation.py:43-57."""
        if not u8_batch.is_cuda:
            raise _lib.SplitVaeError("scramble needs a CUDA tensor: there is no CPU fallback")
        B, H, W, _ = u8_batch.shape
        if perms is None:
            perms = self.draw_permutations(B, H, W, device=u8_batch.device)
        if out is None:
            out = torch.empty(B, H, W, 6, dtype=torch.float32, device=u8_batch.device)
        lib = _lib.load()
        _lib.check(lib.sv_stage_scramble(C.c_void_p(u8_batch.data_ptr()), C.c_void_p(perms.data_ptr()), C.c_void_p(out.data_ptr()),
                                         B, H, W, self.size, C.c_void_p(torch.cuda.current_stream().cuda_stream)), None,
                   "sv_stage_scramble")
        return out
