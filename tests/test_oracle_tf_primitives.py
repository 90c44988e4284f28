"""The oracle's TensorFlow LIBRARY primitives pinned by an implementation of TF's semantics that is not this repository's own reading:
OpenCV's TensorFlow-graph importer executing a real TF GraphDef (Conv2D padding="SAME", ResizeBilinear half_pixel_centers=true),
scripts/make_opencv_primitive_golden.py -> tests/golden/opencv_tf_primitives.npz.

* everywhere: oracle.conv2d_same / resize2x / celeba_preprocess's resize against the committed vectors;
* where cv2 is importable: the same comparison against a live OpenCV run (and the committed vectors against that run), plus a
  random sweep of shapes - so the fixture cannot go stale silently."""
import importlib.util
import os
import sys

import numpy as np
import pytest
import torch
import torch.nn.functional as F

from oracle import splitvae_oracle as O

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLD = os.path.join(ROOT, "tests", "golden", "opencv_tf_primitives.npz")


def _gen():
    spec = importlib.util.spec_from_file_location("make_opencv_primitive_golden", os.path.join(ROOT, "scripts", "make_opencv_primitive_golden.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def _oracle(name, x, blob_get):
    xt = torch.from_numpy(np.asarray(x, dtype=np.float64))
    if name.startswith("conv"):
        return O.conv2d_same(xt, torch.from_numpy(blob_get("w").astype(np.float64)), None, int(blob_get("stride"))).numpy()
    size = tuple(int(v) for v in blob_get("size"))
    if size == (2 * x.shape[1], 2 * x.shape[2]):
        return O.resize2x(xt).numpy()                                            # the decoder's tf.image.resize (vae/model.py:163-167)
    # the CelebA down-scale: the same call celeba_preprocess makes (vae/data.py:85)
    return F.interpolate(xt.permute(0, 3, 1, 2), size=size, mode="bilinear", align_corners=False, antialias=False).permute(0, 2, 3, 1).numpy()


def test_oracle_conv_same_and_resize_match_the_opencv_tf_importer_vectors():
    with np.load(GOLD) as z:
        names = sorted({k.split("/")[0] for k in z.files if "/" in k})
        assert sum(n.startswith("conv") for n in names) == 6 and sum(n.startswith("resize") for n in names) == 4
        for name in names:
            x = z[name + "/x"].astype(np.float32)
            y = z[name + "/y"]
            ref = _oracle(name, x, lambda k: z[name + "/" + k])
            assert ref.shape == y.shape, name
            tol = 3e-5 * max(1.0, float(np.abs(ref).max()))                        # OpenCV computes in fp32, the oracle in fp64
            assert float(np.abs(ref - y).max()) <= tol, (name, float(np.abs(ref - y).max()))


def test_celeba_preprocess_resize_is_the_pinned_call():
    """celeba_preprocess == crop + the pinned down-scale + /255*2-1 (same interpolate arguments as the vector above)"""
    rng = np.random.default_rng(3)
    img = rng.integers(0, 256, size=(218, 178, 3), dtype=np.uint8)
    got = O.celeba_preprocess(img)
    crop = img[20:198].astype(np.float64)
    r = F.interpolate(torch.from_numpy(crop).permute(2, 0, 1)[None], size=(64, 64), mode="bilinear", align_corners=False, antialias=False)
    assert np.allclose(got, (r[0].permute(1, 2, 0).numpy() / 255 * 2 - 1).astype(np.float32), atol=1e-6)


@pytest.mark.skipif(importlib.util.find_spec("cv2") is None, reason="OpenCV not installed")
def test_live_opencv_run_agrees_with_the_fixture_and_a_random_sweep():
    g = _gen()
    cs = g.cases()
    ys = g.opencv_outputs(cs)
    with np.load(GOLD) as z:
        for name, y in ys.items():
            assert np.array_equal(z[name + "/x"].astype(np.float32), cs[name]["x"]), name     # the generator is deterministic
            assert np.allclose(z[name + "/y"], y, atol=1e-6), name
    rng = np.random.default_rng(11)
    for _ in range(12):                                                            # shapes beyond the fixture, incl. odd sizes
        H, W = int(rng.integers(5, 20)), int(rng.integers(5, 20))
        k, s = int(rng.choice([3, 4, 5, 6])), int(rng.choice([1, 2]))
        ci, co = int(rng.integers(1, 6)), int(rng.integers(1, 6))
        x = rng.normal(size=(1, H, W, ci)).astype(np.float32)
        w = rng.normal(size=(k, k, ci, co)).astype(np.float32)
        y = g.run_opencv(g.conv_graph(x.shape, w, s), x)
        ref = O.conv2d_same(torch.from_numpy(x).double(), torch.from_numpy(w).double(), None, s).numpy()
        assert ref.shape == y.shape and float(np.abs(ref - y).max()) <= 3e-5 * max(1.0, float(np.abs(ref).max())), (H, W, k, s)
        y = g.run_opencv(g.resize_graph(x.shape, (2 * H, 2 * W)), x)
        assert float(np.abs(O.resize2x(torch.from_numpy(x).double()).numpy() - y).max()) <= 1e-6
