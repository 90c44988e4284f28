"""Synthetic stand-ins for vae/data.py:get_dataset (the real SVHN / CelebA readers need the network
and the datasets).  Shapes and value grid follow the reference:
uint8 pixels mapped k/255*2-1 (vae/data.py:52); `svhn*` -> 32x32x3, `celeba64` -> 64x64x3."""
from __future__ import annotations

import torch

IMAGE_SIZE = {"svhn": 32, "svhn_no_extra": 32, "celeba64": 64, "celeba128": 128}


def image_shape(dataset):
    if dataset not in IMAGE_SIZE:
        raise NotImplementedError(dataset)  # vae/data.py:21
    s = IMAGE_SIZE[dataset]
    return [-1, s, s, 3]


class SyntheticBatches:
    """Infinite iterator of scrambled batches [B,H,W,6] built on the device from pinned uint8 host
    images: the host->device copy and the scramble kernel are what an input pipeline must do per step."""

    def __init__(self, dataset, batch_size, augmentor, seed=0, pool=8, device="cuda", length=None):
        self.shape = image_shape(dataset)
        self.B, self.S = int(batch_size), self.shape[1]
        self.aug = augmentor
        g = torch.Generator().manual_seed(seed)
        self.pool = [torch.randint(0, 256, (self.B, self.S, self.S, 3), dtype=torch.uint8, generator=g).pin_memory()
                     for _ in range(pool)]
        self.device = device
        self.length = length          # None: infinite (train, `.repeat()` at vae/main.py:57); n: one pass of n batches (test)
        self.i = 0
        self.u8 = torch.empty(self.B, self.S, self.S, 3, dtype=torch.uint8, device=device)

    def __iter__(self):
        if self.length is not None:
            self.i = 0
        return self

    def __next__(self):
        if self.length is not None and self.i >= self.length:
            raise StopIteration
        self.u8.copy_(self.pool[self.i % len(self.pool)], non_blocking=True)
        self.i += 1
        return self.aug.scramble(self.u8)


def get_dataset(dataset, get_label=False, batch_size=64, augmentor=None, seed=0, test_batches=4):
    """Same return contract as vae/data.py:11-21: (train_dataset, test_dataset, image_shape).  The test set is a finite
    pass of `test_batches` synthetic batches of the same batch size (vae/main.py:58-61 batches the test split the same way)."""
    shp = image_shape(dataset)
    train = SyntheticBatches(dataset, batch_size, augmentor, seed=seed)
    test = SyntheticBatches(dataset, batch_size, augmentor, seed=seed + 7919, pool=max(1, test_batches), length=test_batches) \
        if test_batches else None
    return train, test, shp
