"""splitvae_b200.hdf5_lite: the Keras HDF5 weights format of `model.save_weights('models/<run>.h5')` (vae/trainer.py:421) without h5py.

* structure-level known answers of the writer against the HDF5 File Format Specification (superblock v0, old-style groups);
* round trips (scalars, int64 / float64, > 8 entries per group = several symbol-table nodes, empty groups, string attributes);
* the reader on a HAND-ASSEMBLED file that uses the variants libhdf5 / h5py emit and the writer does not: superblock v1, an object
  header continuation block, a two-level group B-tree, attribute message v3 with a variable-length string in a global heap,
  a compact dataset, a big-endian dataset;
* the Keras layout of all three models (layer groups, weight_names, nested dataset paths) from the plan-only variable table.
"""
import struct

import numpy as np
import pytest

from splitvae_b200 import hdf5_lite as H
from splitvae_b200.engine import Engine
from splitvae_b200.model import keras_weight_names

UNDEF = H.UNDEF


def test_superblock_and_root_group_layout():
    root = H.GroupSpec()
    root.attrs["backend"] = "tensorflow"
    root.create_dataset("w", np.arange(6, dtype=np.float32).reshape(2, 3))
    blob = H.dumps(root)
    assert blob[:8] == b"\x89HDF\r\n\x1a\n"
    assert blob[8:16] == bytes([0, 0, 0, 0, 0, 8, 8, 0])                    # versions 0, 8-byte offsets and lengths
    leaf_k, internal_k, flags = struct.unpack_from("<HHI", blob, 16)
    assert (leaf_k, internal_k, flags) == (4, 16, 0)
    base, free, eof, drv = struct.unpack_from("<QQQQ", blob, 24)
    assert (base, free, drv) == (0, UNDEF, UNDEF) and eof == len(blob) and eof % 8 == 0
    name_off, oh, cache, _, bt, hp = struct.unpack_from("<QQIIQQ", blob, 56)
    assert name_off == 0 and cache == 1
    assert blob[bt:bt + 4] == b"TREE" and blob[hp:hp + 4] == b"HEAP"
    # root object header v1: first message is the symbol table message pointing at the same B-tree / heap
    ver, _, nmsg, refs, size = struct.unpack_from("<BBHII", blob, oh)
    assert (ver, nmsg, refs) == (1, 2, 1) and oh % 8 == 0
    mtype, msize, mflags = struct.unpack_from("<HHB", blob, oh + 16)
    assert (mtype, msize) == (0x11, 16) and struct.unpack_from("<QQ", blob, oh + 24) == (bt, hp)
    # B-tree node: group type, leaf level, one child, no siblings, key 0 = the empty string, full node size reserved
    ntype, level, used, left, right, key0, child, key1 = struct.unpack_from("<BBHQQQQQ", blob, bt + 4)
    assert (ntype, level, used, left, right, key0) == (0, 0, 1, UNDEF, UNDEF, 0)
    assert blob[child:child + 4] == b"SNOD" and struct.unpack_from("<BBH", blob, child + 4) == (1, 0, 1)
    # heap: the data segment starts with the empty string; the name sits at an 8-byte aligned offset = the entry's and key 1's offset
    hver, seg_size, free_off, seg = struct.unpack_from("<B3xQQQ", blob, hp + 4)
    assert hver == 0 and blob[seg:seg + 8] == b"\0" * 8
    ent_name, ent_oh, ent_cache = struct.unpack_from("<QQI", blob, child + 8)
    assert ent_name == key1 == 8 and blob[seg + 8:seg + 10] == b"w\0" and ent_cache == 0
    nxt, fsize = struct.unpack_from("<QQ", blob, seg + free_off)
    assert nxt == 1 and free_off + fsize == seg_size                         # one free block to the end of the segment
    # the dataset's header: dataspace, datatype, fill value, layout (contiguous, 24 bytes of float32 at an aligned address)
    msgs = {}
    p = ent_oh + 16
    for _ in range(struct.unpack_from("<H", blob, ent_oh + 2)[0]):
        t, s = struct.unpack_from("<HH", blob, p)
        msgs[t] = blob[p + 8:p + 8 + s]
        p += 8 + s
    assert sorted(msgs) == [1, 3, 5, 8]
    assert msgs[1][:8] == bytes([1, 2, 0, 0, 0, 0, 0, 0]) and struct.unpack_from("<QQ", msgs[1], 8) == (2, 3)
    # IEEE float32, little-endian: class 1 v1, sign bit 31, size 4; bit offset 0, precision 32, exponent 23 / 8, mantissa 0 / 23, bias 127
    assert msgs[3][:20] == bytes([0x11, 0x20, 31, 0, 4, 0, 0, 0, 0, 0, 32, 0, 23, 8, 0, 23, 127, 0, 0, 0])
    lver, lclass, addr, size = struct.unpack_from("<BBQQ", msgs[8], 0)
    assert (lver, lclass, size) == (3, 1, 24) and addr % 8 == 0
    assert np.array_equal(np.frombuffer(blob, "<f4", 6, addr), np.arange(6, dtype=np.float32))


def test_fixed_string_attribute_encoding():
    root = H.GroupSpec()
    root.attrs["layer_names"] = ["encoder", "decoder_1"]
    blob = H.dumps(root)
    oh = struct.unpack_from("<Q", blob, 64)[0]
    p = oh + 16 + 8 + 16                                                     # after the symbol table message
    t, s = struct.unpack_from("<HH", blob, p)
    body = blob[p + 8:p + 8 + s]
    assert t == 0x0C and body[0] == 1
    nlen, dlen, slen = struct.unpack_from("<HHH", body, 2)
    assert (nlen, dlen, slen) == (12, 8, 16) and body[8:20] == b"layer_names\0"
    q = 8 + 16
    assert body[q:q + 8] == bytes([0x13, 0x01, 0, 0, 9, 0, 0, 0])            # string class, null-padded ASCII, 9 bytes
    q += 8
    assert body[q:q + 16] == bytes([1, 1, 0, 0, 0, 0, 0, 0]) + struct.pack("<Q", 2)
    q += 16
    assert body[q:q + 18] == b"encoder\0\0decoder_1"


def test_round_trip_types_and_wide_groups(tmp_path):
    rng = np.random.default_rng(0)
    root = H.GroupSpec()
    root.attrs["keras_version"] = b"2.2.4-tf"
    root.attrs["ints"] = np.arange(5, dtype=np.int32)
    root.attrs["pi"] = np.float64(3.25)
    data = {}
    for i in range(37):                                                       # 5 symbol-table nodes in one group
        data[f"g/v_{i:02d}:0"] = rng.normal(size=(i % 4 + 1, 3)).astype(np.float32)
    data["scalars/iterations:0"] = np.asarray(123456789012, dtype=np.int64)
    data["scalars/f64"] = rng.normal(size=(2, 2, 2))
    data["scalars/u8"] = np.arange(7, dtype=np.uint8)
    data["scalars/empty"] = np.zeros((0, 4), dtype=np.float32)
    for k, v in data.items():
        root.create_dataset(k, v)
    root.require_group("nothing_here")
    path = H.write_file(str(tmp_path / "t.h5"), root)
    with H.File(path) as f:
        assert sorted(f.keys()) == ["g", "nothing_here", "scalars"]
        assert f["nothing_here"].keys() == [] and f.attrs["keras_version"] == b"2.2.4-tf"
        assert np.array_equal(f.attrs["ints"], np.arange(5)) and f.attrs["pi"] == 3.25
        assert f["g"].keys() == sorted(k.split("/")[1] for k in data if k.startswith("g/"))
        for k, v in data.items():
            got = f[k].read()
            assert got.shape == v.shape and got.dtype == v.dtype and np.array_equal(got, v), k
        assert "g/v_00:0" in f and "g/nope" not in f
        with pytest.raises(KeyError):
            f["g/nope"]


def _hand_made_file():
    """A file laid out the way libhdf5 does when attributes are added after creation and groups grow: superblock v1, root header with a
    continuation block, two-level B-tree over three symbol-table nodes, attribute v3 (UTF-8 name) holding a variable-length string
    (global heap), one compact int16 dataset and one big-endian float64 dataset."""
    buf = bytearray(100)                                                      # superblock v1 is 100 bytes

    def alloc(b):
        buf.extend(b"\0" * ((-len(buf)) % 8))
        a = len(buf)
        buf.extend(b)
        return a

    def msg(t, body, flags=0):
        body = body + b"\0" * ((-len(body)) % 8)
        return struct.pack("<HHB3x", t, len(body), flags) + body

    def header(msgs, nmsg=None):
        body = b"".join(msgs)
        return alloc(struct.pack("<BBHII4x", 1, 0, nmsg or len(msgs), 1, len(body)) + body)

    f64be = bytes([0x11, 0x21, 63, 0, 8, 0, 0, 0]) + struct.pack("<HHBBBBI", 0, 64, 52, 11, 0, 52, 1023)
    i16 = bytes([0x10, 0x08, 0, 0, 2, 0, 0, 0]) + struct.pack("<HH", 0, 16)
    space1 = lambda n: bytes([1, 1, 0, 0, 0, 0, 0, 0]) + struct.pack("<Q", n)       # noqa: E731
    be = np.array([1.5, -2.25, 1e300], dtype=">f8")
    a_be = alloc(be.tobytes())
    d_be = header([msg(1, space1(3)), msg(3, f64be, 1), msg(8, struct.pack("<BBQQ", 3, 1, a_be, 24))])
    compact = np.array([-3, 7], dtype="<i2").tobytes()
    d_compact = header([msg(1, space1(2)), msg(3, i16, 1), msg(8, struct.pack("<BBH", 3, 0, 4) + compact)])
    # global heap collection with one object (index 1) = the UTF-8 text, then the free-space object 0
    text = "tensorflow-ü".encode("utf8")
    obj = struct.pack("<HH4xQ", 1, 1, len(text)) + text + b"\0" * ((-len(text)) % 8)
    gcol_size = 16 + len(obj) + 16
    gcol = alloc(b"GCOL" + struct.pack("<B3xQ", 1, gcol_size) + obj + struct.pack("<HH4xQ", 0, 0, 0))
    # attribute v3: flags 0, name 'backend' (7 + null), vlen-string datatype (class 9 v1, type 1 = string, pad 0, cset 1 = UTF-8, size 16,
    # base type = 1-byte string), scalar dataspace v2 (type 0 = scalar); no padding between the parts
    vlen = bytes([0x19, 0x01, 0x01, 0, 16, 0, 0, 0]) + bytes([0x13, 0x10, 0, 0, 1, 0, 0, 0])
    sp2 = bytes([2, 0, 0, 0])
    name = b"backend\0"
    attr3 = struct.pack("<BBHHHB", 3, 0, len(name), len(vlen), len(sp2), 1) + name + vlen + sp2 + struct.pack("<IQI", len(text), gcol, 1)
    # group with 18 entries in three symbol-table nodes under a level-1 B-tree with two level-0 children
    names = [f"n{i:02d}" for i in range(17)] + ["zz"]
    targets = {n: d_compact for n in names}
    targets["n03"] = d_be
    seg = bytearray(8)
    off = {}
    for n in names:
        off[n] = len(seg)
        seg += n.encode() + b"\0" * (8 - len(n))
    seg += struct.pack("<QQ", 1, 32) + b"\0" * 16
    seg_addr = alloc(bytes(seg))
    heap = alloc(b"HEAP" + struct.pack("<B3xQQQ", 0, len(seg), len(seg) - 32, seg_addr))

    def snod(chunk):
        node = b"SNOD" + struct.pack("<BBH", 1, 0, len(chunk))
        for n in chunk:
            node += struct.pack("<QQII16x", off[n], targets[n], 0, 0)
        return alloc(node + b"\0" * (328 - len(node)))

    def tree(level, children, last_names, left=UNDEF, right=UNDEF):
        t = b"TREE" + struct.pack("<BBHQQ", 0, level, len(children), left, right) + struct.pack("<Q", 0)
        for c, n in zip(children, last_names):
            t += struct.pack("<QQ", c, off[n])
        return alloc(t + b"\0" * (544 - len(t)))

    chunks = [names[0:8], names[8:16], names[16:18]]
    nodes = [snod(c) for c in chunks]
    t0 = tree(0, nodes[:2], [chunks[0][-1], chunks[1][-1]])
    t1 = tree(0, nodes[2:], [chunks[2][-1]], left=t0)
    top = tree(1, [t0, t1], [chunks[1][-1], chunks[2][-1]])
    # root header: symbol table message + continuation message; the attribute and a null message live in the continuation block
    cont = alloc(msg(0x0C, attr3) + msg(0, b"\0" * 8))
    cont_len = len(msg(0x0C, attr3) + msg(0, b"\0" * 8))
    root = header([msg(0x11, struct.pack("<QQ", top, heap)), msg(0x10, struct.pack("<QQ", cont, cont_len))], nmsg=4)
    buf.extend(b"\0" * ((-len(buf)) % 8))
    sb = b"\x89HDF\r\n\x1a\n" + bytes([1, 0, 0, 0, 0, 8, 8, 0]) + struct.pack("<HHIHH", 4, 16, 0, 32, 0)
    sb += struct.pack("<QQQQ", 0, UNDEF, len(buf), UNDEF) + struct.pack("<QQII", 0, root, 1, 0) + struct.pack("<QQ", top, heap)
    assert len(sb) == 100
    buf[:100] = sb
    return bytes(buf), names, be


def test_reader_handles_libhdf5_variants():
    blob, names, be = _hand_made_file()
    f = H.File(blob)
    assert f.keys() == names                                                  # three symbol-table nodes, two B-tree levels
    assert f.attrs["backend"] == "tensorflow-ü".encode("utf8")               # attribute v3, vlen string from the global heap, continuation block
    assert np.array_equal(f["n00"].read(), np.array([-3, 7], dtype=np.int16))  # compact layout
    got = f["n03"].read()
    assert got.dtype == np.float64 and np.array_equal(got, be.astype(np.float64))   # big-endian data comes back native
    assert np.array_equal(f["zz"][1:], [7])


def test_reader_rejects_what_it_does_not_support():
    with pytest.raises(ValueError):
        H.File(b"not an hdf5 file" * 100)
    blob = bytearray(H.dumps(H.GroupSpec()))
    blob[8] = 2                                                               # superblock v2 (libver='latest')
    with pytest.raises(NotImplementedError):
        H.File(bytes(blob))
    # chunked layout
    root = H.GroupSpec()
    root.create_dataset("w", np.zeros(4, np.float32))
    blob = bytearray(H.dumps(root))
    i = blob.index(struct.pack("<HHB3xBB", 8, 24, 0, 3, 1))
    blob[i + 9] = 2
    with pytest.raises(NotImplementedError):
        H.File(bytes(blob))["w"].read()


@pytest.mark.parametrize("kind,top", [("lgvae", ["encoder", "encoder_1", "decoder", "decoder_1"]),
                                      ("lggmvae", ["encoder", "encoder_1", "decoder", "decoder_1"]),
                                      ("gmvae", ["encoder", "decoder"])])
def test_keras_weight_file_layout(kind, top, tmp_path):
    """save_weights_to_hdf5_group's layout for this architecture, read back through load_weights_from_hdf5_group's traversal"""
    e = Engine(model=kind, height=32, width=32, batch=4, plan_only=True)
    names = keras_weight_names(kind)
    assert sorted(names) == sorted(t[0] for t in e.table)
    rng = np.random.default_rng(1)
    params = {name: rng.normal(size=shape).astype(np.float32) for name, shape, _, _ in e.table}
    layers = {}
    for ours, keras in names.items():
        layers.setdefault(keras.split("/")[1], []).append((keras, params[ours]))
    path = H.save_keras_weights(str(tmp_path / "w.h5"), list(layers.items()))
    flat, f = H.load_keras_weights(path)
    assert [n.decode() for n in f.attrs["layer_names"]] == top
    assert f.attrs["backend"] == b"tensorflow" and f.attrs["keras_version"] == b"2.2.4-tf"
    model = {"lgvae": "lg_vae", "lggmvae": "lggm_vae", "gmvae": "gm_vae"}[kind]
    for layer in top:
        g = f[layer]
        assert g.keys() == [model]                                            # weight names nest as <model>/<layer>/<sublayer>/<var>:0
        wn = [n.decode() for n in g.attrs["weight_names"]]
        assert wn == [k for k in names.values() if k.split("/")[1] == layer]
        assert all(n.startswith(f"{model}/{layer}/") and n.endswith((":0",)) for n in wn)
        for n in wn:
            assert g[n].dtype == np.dtype("<f4")
    for ours, keras in names.items():
        assert np.array_equal(flat[keras], params[ours]), ours
    conv = f["decoder"][f"{model}/decoder"]
    assert any(k.startswith("conv2d") for k in conv.keys()) and any(k.startswith("dense") for k in conv.keys())
