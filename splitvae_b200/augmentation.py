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
    def __init__(self, type, size=1, mean=0, std=1, seed=0):
        if type != "scramble":
            raise NotImplementedError(f"augmentation '{type}' is outside the hot path of this build")
        self.size = int(size)
        self.augment = self.scramble
        self.seed, self._draws = int(seed), 0

    def draw_permutations(self, batch, height, width, generator=None, device="cuda"):
        """One uniform permutation of the (H/p)*(W/p) patches per image (tf.random.shuffle, augmentation.py:49), drawn by a device
        kernel (sv_draw_permutations: Philox keys + an in-block sort); a torch `generator` selects the older torch.rand + argsort path."""
        n_patch = (height // self.size) * (width // self.size)
        if generator is not None or n_patch > 4096:
            keys = torch.rand(batch, n_patch, generator=generator, device=device)
            return torch.argsort(keys, dim=1).to(torch.int32)
        perm = torch.empty(batch, n_patch, dtype=torch.int32, device=device)
        lib = _lib.load()
        with torch.cuda.device(perm.device):
            _lib.check(lib.sv_draw_permutations(C.c_void_p(perm.data_ptr()), batch, n_patch, C.c_uint64(self.seed & (2 ** 64 - 1)),
                                                C.c_uint64(self._draws), C.c_void_p(torch.cuda.current_stream().cuda_stream)), None,
                       "sv_draw_permutations")
        self._draws += 1
        return perm

    def scramble_resized(self, u8_batch, height, width, crop=178, perms=None, out=None):
        """CelebA path (vae/data.py:82-87 + augmentation.py:43-57): [B,Hs,Ws,3] decoded uint8 images -> centre crop `crop` x `crop` ->
        bilinear resize to height x width -> /255*2-1 -> scramble, one kernel.  Returns [B,height,width,6] float32."""
        if not u8_batch.is_cuda:
            raise _lib.SplitVaeError("scramble_resized needs a CUDA tensor: there is no CPU fallback")
        B, Hs, Ws, _ = u8_batch.shape
        if Hs < crop or Ws < crop:
            raise ValueError(f"images of {Hs}x{Ws} are smaller than the {crop}x{crop} crop")
        if perms is None:
            perms = self.draw_permutations(B, height, width, device=u8_batch.device)
        if out is None:
            out = torch.empty(B, height, width, 6, dtype=torch.float32, device=u8_batch.device)
        lib = _lib.load()
        _lib.check(lib.sv_stage_resize_scramble(C.c_void_p(u8_batch.data_ptr()), C.c_void_p(perms.data_ptr()), C.c_void_p(out.data_ptr()),
                                                B, Hs, Ws, (Hs - crop) // 2, (Ws - crop) // 2, crop, crop, height, width, self.size,
                                                C.c_void_p(torch.cuda.current_stream().cuda_stream)), None, "sv_stage_resize_scramble")
        return out

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
