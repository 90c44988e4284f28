#!/usr/bin/env python
"""Converts weight checkpoints between this build's `.npz` (model.save_weights) and the reference's Keras HDF5 format
(`model.save_weights('models/<run>.h5')`, vae/trainer.py:421).

    python scripts/convert_checkpoint.py npz2h5 --model lgvae weights.npz weights.h5
    python scripts/convert_checkpoint.py h52npz --model lgvae weights.h5 weights.npz

The model classes read and write `.h5` themselves (`model.save_weights('run.h5')` / `load_weights`, through splitvae_b200.hdf5_lite);
this converter is the stand-alone path and the cross-check: with h5py installed it goes through libhdf5 (so a file of either writer
can be pushed through the other's reader), without h5py it falls back to hdf5_lite.  The variable layouts are identical on both sides (conv HWIO, dense [in,out], bias [out]); only the names differ.  The Keras-side names come
from splitvae_b200.model.keras_weight_names (derived from Keras' naming rules, see its docstring): one HDF5 group per sub-model
layer (`encoder`, `encoder_1`, `decoder`, `decoder_1`) with a `weight_names` attribute, datasets named `<model>/<layer>/.../kernel:0`."""
import argparse
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from splitvae_b200.model import keras_weight_names  # noqa: E402


def _h5py():
    try:
        import h5py
        return h5py
    except ImportError:
        return None


def npz2h5(kind, src, dst):
    h5py = _h5py()
    names = keras_weight_names(kind)
    with np.load(src) as z:
        blob = {k: z[k] for k in z.files if k in names}
    missing = [k for k in names if k not in blob]
    if missing:
        raise KeyError(f"{src} lacks {missing[:3]} ...")
    groups = {}
    for ours, keras in names.items():
        groups.setdefault(keras.split("/")[1], []).append((keras, blob[ours]))
    if h5py is None:
        from splitvae_b200 import hdf5_lite
        hdf5_lite.save_keras_weights(dst, list(groups.items()))
        return
    with h5py.File(dst, "w") as f:
        f.attrs["layer_names"] = [g.encode() for g in groups]
        f.attrs["backend"] = b"tensorflow"
        f.attrs["keras_version"] = b"2.2.4-tf"
        for g, ws in groups.items():
            grp = f.create_group(g)
            grp.attrs["weight_names"] = [n.encode() for n, _ in ws]
            for n, a in ws:
                grp.create_dataset(n, data=np.asarray(a, dtype=np.float32))


def h52npz(kind, src, dst):
    h5py = _h5py()
    names = keras_weight_names(kind)
    out = {}
    if h5py is None:
        from splitvae_b200 import hdf5_lite
        flat, _ = hdf5_lite.load_keras_weights(src)
    else:
        with h5py.File(src, "r") as f:
            flat = {}
            for g in f.keys():
                grp = f[g]
                for n in grp.attrs.get("weight_names", []):
                    n = n.decode() if isinstance(n, bytes) else n
                    flat[n] = np.asarray(grp[n])
    for ours, keras in names.items():
        if keras not in flat:
            raise KeyError(f"{src} has no dataset {keras} (found e.g. {list(flat)[:3]})")
        out[ours] = flat[keras].astype(np.float32)
    np.savez(dst, **out)


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("direction", choices=["npz2h5", "h52npz"])
    ap.add_argument("--model", default="lgvae", choices=["lgvae", "lggmvae", "gmvae"])
    ap.add_argument("src")
    ap.add_argument("dst")
    a = ap.parse_args()
    (npz2h5 if a.direction == "npz2h5" else h52npz)(a.model, a.src, a.dst)
    print("wrote", a.dst)
