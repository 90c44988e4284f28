"""vae/data.py:get_dataset for this build: synthetic batches by default (the datasets need the network, SURVEY.md section 2
row 5), plus readers for local copies of the real data (`data_root=`): the SVHN `.mat` files and the aligned CelebA images
(JPEGs -> centre crop 178 -> bilinear resize 64, data.py:77-134, done on the device).  Shapes and value grid follow the reference:
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


class SvhnBatches:
    """The SVHN cropped-digits `.mat` files of vae/data.py:23-75 (`train_32x32.mat`, `extra_32x32.mat`, `test_32x32.mat` with
    X [32,32,3,N] uint8 and y [N,1], digit 0 stored as label 10), read from `root` (no download: there is no network here).
    Images stay uint8 in pinned host memory; a batch is gathered on the host, copied to the device and handed to the scramble
    kernel, which applies the reference's /255*2-1 scaling (data.py:52) and builds x_hat (augmentation.py:43-57).
      train: shuffled, repeating, full batches only (`shuffle(20000).repeat()...batch(B)`, vae/main.py:57)
      test : one pass, the last batch may be partial (`batch(B)` without drop_remainder, vae/main.py:58)
    With get_label the iterator yields (images, one_hot(y - 1, 10)) like the reference ("0 one-hot at the last index")."""

    def __init__(self, root, split, batch_size, augmentor, get_label=False, extra=True, seed=0, device="cuda"):
        import os

        import numpy as np
        from scipy.io import loadmat
        files = {"train": ["train_32x32.mat"] + (["extra_32x32.mat"] if extra else []), "test": ["test_32x32.mat"]}[split]
        xs, ys = [], []
        for f in files:
            path = os.path.join(root, f)
            if not os.path.exists(path):
                raise FileNotFoundError(f"{path} (the SVHN files are not downloaded by this build; see vae/data.py:35-43 for the URLs)")
            m = loadmat(path)
            xs.append(np.ascontiguousarray(m["X"].transpose((3, 0, 1, 2))))      # data.py:46
            ys.append(np.asarray(m["y"]).reshape(-1).astype(np.int64))
        self.x = torch.from_numpy(np.concatenate(xs))
        self.y = torch.from_numpy(np.concatenate(ys))
        if torch.cuda.is_available():
            self.x = self.x.pin_memory()
        self.split, self.B, self.aug, self.get_label, self.device = split, int(batch_size), augmentor, get_label, device
        self.gen = torch.Generator().manual_seed(seed)
        self.shape = [-1, 32, 32, 3]
        self._order, self._pos = None, 0

    def __len__(self):
        return self.x.shape[0]

    def host_batch(self):
        """(uint8 images [b,32,32,3], one-hot labels [b,10]) of the next batch; raises StopIteration at the end of a test pass."""
        n = self.x.shape[0]
        if self.split == "train":
            if self._order is None or self._pos + self.B > n:      # a new shuffled epoch (full batches only)
                self._order, self._pos = torch.randperm(n, generator=self.gen), 0
            idx = self._order[self._pos:self._pos + self.B]
        else:
            if self._pos >= n:
                raise StopIteration
            idx = torch.arange(self._pos, min(self._pos + self.B, n))
        self._pos += idx.numel()
        labels = torch.nn.functional.one_hot((self.y[idx] - 1) % 10, 10).float()   # data.py:56: digit 0 (label 10) -> index 9
        return self.x[idx], labels

    def __iter__(self):
        if self.split != "train":
            self._pos = 0
        return self

    def __next__(self):
        u8, labels = self.host_batch()
        images = self.aug.scramble(u8.to(self.device, non_blocking=True))
        return (images, labels) if self.get_label else images


class CelebaBatches:
    """The aligned CelebA images of vae/data.py:77-134 read from `root/img_align_celeba/*` (the directory the reference unzips;
    no download here).  The reference decodes every JPEG once, centre-crops 178x178, resizes to 64x64 (bilinear) and caches float
    tensors in a TFRecord; here the DECODED uint8 images are cached once in `root/celeba_u8.npy` (memory-mapped afterwards) and the crop,
    resize, /255*2-1 scaling and the scramble run in one device kernel per batch (sv_stage_resize_scramble).
      split: the first tenth of the (sorted) file list is the test set, the rest the train set (data.py:90-92)
      train: shuffled, repeating, full batches (vae/main.py:57); test: one pass, the last batch may be partial (vae/main.py:58).
    CelebA carries no labels in the reference (README runs it with -no_label)."""

    def __init__(self, root, split, batch_size, augmentor, size=64, seed=0, device="cuda", threads=8):
        import glob
        import os

        import numpy as np
        self.size, self.B, self.aug, self.split, self.device = int(size), int(batch_size), augmentor, split, device
        cache = os.path.join(root, "celeba_u8.npy")
        if not os.path.exists(cache):
            files = sorted(glob.glob(os.path.join(root, "img_align_celeba", "*")))
            if not files:
                raise FileNotFoundError(f"{root}/img_align_celeba/* (CelebA is not downloaded by this build; vae/data.py:113 has the source)")
            from concurrent.futures import ThreadPoolExecutor

            from PIL import Image

            def load(path):
                with Image.open(path) as im:
                    return np.asarray(im.convert("RGB"), dtype=np.uint8)
            first = load(files[0])
            arr = np.lib.format.open_memmap(cache + ".tmp", mode="w+", dtype=np.uint8, shape=(len(files),) + first.shape)
            with ThreadPoolExecutor(max_workers=threads) as ex:     # (the reference maps its decode over 8 threads, vae/main.py:57)
                for i, a in enumerate(ex.map(load, files)):
                    if a.shape != first.shape:
                        raise ValueError(f"{files[i]}: {a.shape} differs from {first.shape}")
                    arr[i] = a
            arr.flush()
            del arr
            os.replace(cache + ".tmp", cache)
        self.all = np.load(cache, mmap_mode="r")
        n = self.all.shape[0]
        self.index = np.arange(0, n // 10) if split == "test" else np.arange(n // 10, n)
        self.gen = np.random.default_rng(seed)
        self.shape = [-1, self.size, self.size, 3]
        self._order, self._pos = None, 0
        self.host = torch.empty((self.B,) + tuple(self.all.shape[1:]), dtype=torch.uint8)
        if torch.cuda.is_available():
            self.host = self.host.pin_memory()

    def __len__(self):
        return len(self.index)

    def host_batch(self):
        import numpy as np
        n = len(self.index)
        if self.split == "train":
            if self._order is None or self._pos + self.B > n:
                self._order, self._pos = self.gen.permutation(n), 0
            idx = self._order[self._pos:self._pos + self.B]
        else:
            if self._pos >= n:
                raise StopIteration
            idx = np.arange(self._pos, min(self._pos + self.B, n))
        self._pos += len(idx)
        rows = np.sort(self.index[idx])                      # sorted gather out of the memory map
        out = self.host[:len(rows)]
        out.numpy()[...] = self.all[rows]
        return out

    def __iter__(self):
        if self.split != "train":
            self._pos = 0
        return self

    def __next__(self):
        u8 = self.host_batch().to(self.device, non_blocking=True)
        return self.aug.scramble_resized(u8, self.size, self.size)


def get_dataset(dataset, get_label=False, batch_size=64, augmentor=None, seed=0, test_batches=4, data_root=None):
    """Same return contract as vae/data.py:11-21: (train_dataset, test_dataset, image_shape).  The test set is a finite
    pass of `test_batches` synthetic batches of the same batch size (vae/main.py:58-61 batches the test split the same way)."""
    shp = image_shape(dataset)
    if data_root is not None and dataset.lower() in ("svhn", "svhn_no_extra"):     # the real files, when the caller has them
        extra = dataset.lower() == "svhn"
        return (SvhnBatches(data_root, "train", batch_size, augmentor, get_label, extra, seed),
                SvhnBatches(data_root, "test", batch_size, augmentor, get_label, extra, seed), shp)
    if data_root is not None and dataset.lower() in ("celeba64", "celeba128"):     # vae/data.py:16-19
        return (CelebaBatches(data_root, "train", batch_size, augmentor, shp[1], seed),
                CelebaBatches(data_root, "test", batch_size, augmentor, shp[1], seed), shp)
    train = SyntheticBatches(dataset, batch_size, augmentor, seed=seed)
    test = SyntheticBatches(dataset, batch_size, augmentor, seed=seed + 7919, pool=max(1, test_batches), length=test_batches) \
        if test_batches else None
    return train, test, shp
