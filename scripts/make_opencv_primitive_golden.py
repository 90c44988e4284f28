#!/usr/bin/env python
"""Pins the oracle's two TensorFlow LIBRARY primitives - Conv2D(padding='SAME') and tf.image.resize (bilinear, half-pixel centres) -
with an implementation of TensorFlow's semantics that is not this repository's reading of them: OpenCV's TensorFlow-graph importer
(cv2.dnn.readNetFromTensorflow).  The script writes a real TF GraphDef (binary protobuf, hand-encoded: Placeholder -> Conv2D with
padding "SAME" / ResizeBilinear with half_pixel_centers) for each case, lets OpenCV execute it, and stores inputs + outputs as
tests/golden/opencv_tf_primitives.npz.  tests/test_oracle.py checks oracle/splitvae_oracle.py against the stored vectors everywhere
and against a live OpenCV run where cv2 is importable.

    python scripts/make_opencv_primitive_golden.py            # needs cv2 (4.13 in this image)

Cases follow the path's layers: 6x6 / 4x4 kernels at stride 2 (encoders, vae/model.py:36-40) and stride 1 (decoders, model.py:152-156),
even and odd image sizes; resize x2 of the decoder (model.py:163-167) and the CelebA down-scale by 178 / 64 (vae/data.py:85)."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


# ---- a minimal protobuf writer (wire format only) -------------------------------------------------------------
def varint(n):
    out = b""
    n &= (1 << 64) - 1
    while True:
        b = n & 0x7F
        n >>= 7
        if n:
            out += bytes([b | 0x80])
        else:
            return out + bytes([b])


def _key(field, wire):
    return varint((field << 3) | wire)


def f_bytes(field, b):
    return _key(field, 2) + varint(len(b)) + b


def f_str(field, s):
    return f_bytes(field, s.encode())


def f_int(field, i):
    return _key(field, 0) + varint(i)


# ---- tensorflow/core/framework/{graph,node_def,attr_value,tensor,tensor_shape}.proto ---------------------------
DT_FLOAT, DT_INT32 = 1, 3


def shape_proto(dims):                      # TensorShapeProto { repeated Dim dim = 2 { int64 size = 1 } }
    return b"".join(f_bytes(2, f_int(1, d)) for d in dims)


def attr(name, value):                      # NodeDef.attr = 5: map<string, AttrValue> entry { key = 1, value = 2 }
    return f_bytes(5, f_str(1, name) + f_bytes(2, value))


def a_type(t):                              # AttrValue.type = 6
    return f_int(6, t)


def a_s(s):                                 # AttrValue.s = 2
    return f_bytes(2, s.encode())


def a_b(b):                                 # AttrValue.b = 5
    return f_int(5, 1 if b else 0)


def a_list_i(xs):                           # AttrValue.list = 1 { repeated int64 i = 3 [packed] }
    return f_bytes(1, f_bytes(3, b"".join(varint(x) for x in xs)))


def a_shape(dims):                          # AttrValue.shape = 7
    return f_bytes(7, shape_proto(dims))


def a_tensor(arr, dt):                      # AttrValue.tensor = 8: TensorProto { dtype = 1, tensor_shape = 2, tensor_content = 4 }
    return f_bytes(8, f_int(1, dt) + f_bytes(2, shape_proto(arr.shape)) + f_bytes(4, arr.tobytes()))


def node(name, op, inputs, attrs):          # GraphDef.node = 1: NodeDef { name = 1, op = 2, input = 3, attr = 5 }
    return f_bytes(1, f_str(1, name) + f_str(2, op) + b"".join(f_str(3, i) for i in inputs) + b"".join(attrs))


def conv_graph(shape_nhwc, w_hwio, stride):
    g = node("input", "Placeholder", [], [attr("dtype", a_type(DT_FLOAT)), attr("shape", a_shape(shape_nhwc))])
    g += node("w", "Const", [], [attr("dtype", a_type(DT_FLOAT)), attr("value", a_tensor(w_hwio.astype("<f4"), DT_FLOAT))])
    g += node("conv", "Conv2D", ["input", "w"], [attr("T", a_type(DT_FLOAT)), attr("strides", a_list_i([1, stride, stride, 1])),
                                                 attr("padding", a_s("SAME")), attr("data_format", a_s("NHWC")),
                                                 attr("dilations", a_list_i([1, 1, 1, 1]))])
    return g


def resize_graph(shape_nhwc, out_hw):
    g = node("input", "Placeholder", [], [attr("dtype", a_type(DT_FLOAT)), attr("shape", a_shape(shape_nhwc))])
    g += node("size", "Const", [], [attr("dtype", a_type(DT_INT32)), attr("value", a_tensor(np.array(out_hw, dtype="<i4"), DT_INT32))])
    g += node("resize", "ResizeBilinear", ["input", "size"], [attr("T", a_type(DT_FLOAT)), attr("align_corners", a_b(False)),
                                                              attr("half_pixel_centers", a_b(True))])
    return g


def run_opencv(graph_bytes, x_nhwc):
    import cv2
    net = cv2.dnn.readNetFromTensorflow(np.frombuffer(graph_bytes, np.uint8))
    net.setInput(np.ascontiguousarray(x_nhwc.transpose(0, 3, 1, 2)))       # OpenCV's blobs are NCHW; its importer maps the NHWC graph
    return np.ascontiguousarray(net.forward().transpose(0, 2, 3, 1))


CONV_CASES = [(16, 6, 2, 3, 8), (16, 4, 2, 8, 8), (8, 4, 1, 8, 4), (16, 6, 1, 4, 6), (7, 6, 2, 3, 4), (9, 4, 1, 2, 2)]   # H, k, stride, Ci, Co
RESIZE_CASES = [((8, 8, 4), (16, 16)), ((5, 7, 3), (10, 14)), ((16, 16, 2), (32, 32)), ((89, 89, 3), (32, 32))]   # (H, W, C), out; the last one has
#                the CelebA pre-processing's scale (178 / 64 = 89 / 32) and uint8-valued pixels


def cases():
    rng = np.random.default_rng(20260401)
    out = {}
    for i, (H, k, s, ci, co) in enumerate(CONV_CASES):
        x = rng.normal(size=(2, H, H, ci)).astype(np.float32)
        w = rng.normal(size=(k, k, ci, co)).astype(np.float32)
        out[f"conv{i}"] = dict(x=x, w=w, stride=s)
    for i, (hwc, size) in enumerate(RESIZE_CASES):
        if hwc[0] > 64:
            x = rng.integers(0, 256, size=(1,) + hwc).astype(np.float32)
        else:
            x = rng.normal(size=(2,) + hwc).astype(np.float32)
        out[f"resize{i}"] = dict(x=x, size=size)
    return out


def opencv_outputs(cs):
    res = {}
    for name, c in cs.items():
        if name.startswith("conv"):
            res[name] = run_opencv(conv_graph(c["x"].shape, c["w"], c["stride"]), c["x"])
        else:
            res[name] = run_opencv(resize_graph(c["x"].shape, c["size"]), c["x"])
    return res


if __name__ == "__main__":
    import cv2
    cs = cases()
    ys = opencv_outputs(cs)
    blob = {"opencv_version": np.asarray(cv2.__version__)}
    for name, c in cs.items():
        blob[name + "/x"] = c["x"].astype(np.uint8) if c["x"].shape[1] > 64 else c["x"]
        if "w" in c:
            blob[name + "/w"] = c["w"]
            blob[name + "/stride"] = np.asarray(c["stride"])
        else:
            blob[name + "/size"] = np.asarray(c["size"])
        blob[name + "/y"] = ys[name]
    path = os.path.join(ROOT, "tests", "golden", "opencv_tf_primitives.npz")
    np.savez_compressed(path, **blob)
    print("wrote", path, os.path.getsize(path), "bytes, OpenCV", cv2.__version__)
