"""A small, dependency-free reader / writer for the subset of HDF5 that Keras weight files use.

Why: the reference ends training with `model.save_weights('models/<run>.h5')` (vae/trainer.py:421), i.e. a Keras HDF5 weights file
written through h5py with the library's default ("earliest") file format.  h5py / libhdf5 are not part of this image and cannot be
installed, so the interop is implemented directly against the HDF5 File Format Specification (version 1 structures):

    superblock v0 / v1                      - FFS III.A "Disk Format: Level 0A"
    B-tree v1 group nodes, SNOD, local heap - FFS III.B / III.C / III.D "Level 1A / 1B / 1D"
    global heap (variable-length strings)   - FFS III.E
    object header v1 + continuation blocks  - FFS IV.A.1.a
    messages: dataspace 0x01 (v1, v2), datatype 0x03 (fixed-point, float, fixed / variable-length string), fill value 0x05,
              data layout 0x08 (v3 compact / contiguous), attribute 0x0C (v1, v2, v3), continuation 0x10, symbol table 0x11

The WRITER emits exactly the layout libhdf5 1.8 / 1.10 produce for such files (old-style groups, contiguous little-endian datasets,
null-padded fixed-length ASCII string attributes the way h5py stores numpy `S` arrays), the READER accepts the variants libhdf5 emits
for them (header continuation blocks, multi-level group B-trees, attribute message versions 1-3, variable-length string attributes).
Not supported (raises): superblock v2+ / object header v2 (`libver='latest'`), chunked or filtered datasets, shared messages.
Pinning: there is no libhdf5 in the image to cross-check against, so the byte layout is covered by structure-level known-answer tests
(tests/test_hdf5_lite.py) and by round trips; scripts/convert_checkpoint.py remains the h5py-side path on a host that has it.

Mapping-like API (a deliberately tiny mirror of h5py's): `File(path)` -> `Group` (`keys()`, `[name]`, `attrs`) / `Dataset` (`[...]` via
`read()`, `shape`, `dtype`, `attrs`); `write_file(path, GroupSpec)` writes a tree built with `require_group` / `create_dataset`.
"""
from __future__ import annotations

import struct

import numpy as np

UNDEF = 0xFFFFFFFFFFFFFFFF
SIGNATURE = b"\x89HDF\r\n\x1a\n"
LEAF_K = 4            # symbol-table nodes hold up to 2 * LEAF_K entries (libhdf5 default)
INTERNAL_K = 16       # group B-tree nodes hold up to 2 * INTERNAL_K children (libhdf5 default)
HEAP_FREE_NULL = 1    # "end of the free list" as libhdf5 stores it (H5HL_FREE_NULL)


def _pad8(n):
    return (n + 7) & ~7


def _contig(a):
    return np.asarray(a, order="C")        # (np.ascontiguousarray would turn a 0-d array into shape (1,))


# ====================================================================================================== writer

class GroupSpec:
    def __init__(self, children=None, attrs=None):
        self.children = dict(children or {})     # name -> GroupSpec | numpy array
        self.attrs = dict(attrs or {})

    def require_group(self, path):
        g = self
        for part in [p for p in path.split("/") if p]:
            nxt = g.children.get(part)
            if nxt is None:
                nxt = g.children[part] = GroupSpec()
            elif not isinstance(nxt, GroupSpec):
                raise ValueError(f"{part} is a dataset")
            g = nxt
        return g

    def create_dataset(self, path, data):
        """h5py semantics: a name with slashes creates the intermediate groups"""
        parts = [p for p in path.split("/") if p]
        g = self.require_group("/".join(parts[:-1]))
        g.children[parts[-1]] = _contig(data)


def _datatype_msg(dt):
    """Datatype message body (FFS IV.A.2.d) for a little-endian numpy dtype"""
    dt = np.dtype(dt)
    if dt.kind == "f" and dt.itemsize in (4, 8):
        if dt.itemsize == 4:
            sign, exp_loc, exp_size, man_size, bias = 31, 23, 8, 23, 127
        else:
            sign, exp_loc, exp_size, man_size, bias = 63, 52, 11, 52, 1023
        # class 1 (floating point) version 1; bit field: little-endian, mantissa normalisation 2 (implied msb), sign bit location
        return struct.pack("<BBBBI", 0x11, 0x20, sign, 0, dt.itemsize) + struct.pack("<HHBBBBI", 0, dt.itemsize * 8, exp_loc, exp_size, 0,
                                                                                      man_size, bias)
    if dt.kind in "iu" and dt.itemsize in (1, 2, 4, 8):
        return struct.pack("<BBBBI", 0x10, 0x08 if dt.kind == "i" else 0x00, 0, 0, dt.itemsize) + struct.pack("<HH", 0, dt.itemsize * 8)
    if dt.kind == "S":
        # class 3 (string): null-padded (1), ASCII (0) - what h5py makes of a numpy bytes array
        return struct.pack("<BBBBI", 0x13, 0x01, 0, 0, max(dt.itemsize, 1))
    raise TypeError(f"unsupported dtype {dt}")


def _dataspace_msg(shape):
    """Dataspace message v1 (FFS IV.A.2.b): rank 0 = scalar"""
    return struct.pack("<BBBB4x", 1, len(shape), 0, 0) + b"".join(struct.pack("<Q", int(d)) for d in shape)


def _as_attr_array(value):
    if isinstance(value, str):
        value = value.encode("utf8")
    if isinstance(value, (bytes, np.bytes_)):
        return np.array(bytes(value), dtype=f"S{max(len(value), 1)}")
    if isinstance(value, (list, tuple)) and value and isinstance(value[0], (str, bytes, np.bytes_)):
        value = [v.encode("utf8") if isinstance(v, str) else bytes(v) for v in value]
        return np.array(value, dtype=f"S{max(max(len(v) for v in value), 1)}")
    a = np.asarray(value)
    if a.dtype.kind == "U":
        a = np.char.encode(a, "utf8")
    if a.dtype.kind not in "fiuS":
        raise TypeError(f"unsupported attribute dtype {a.dtype}")
    if a.dtype.byteorder == ">":
        a = a.astype(a.dtype.newbyteorder("<"))
    return _contig(a)


def _message(mtype, body, flags=0):
    body = body + b"\0" * (_pad8(len(body)) - len(body))
    if len(body) > 0xFFFF:
        raise ValueError("header message larger than 64 KiB (object header v1 limit)")
    return struct.pack("<HHB3x", mtype, len(body), flags) + body


def _attribute_msg(name, value):
    """Attribute message v1 (FFS IV.A.2.m): every part padded to 8 bytes"""
    a = _as_attr_array(value)
    nm = name.encode("utf8") + b"\0"
    dt = _datatype_msg(a.dtype)
    ds = _dataspace_msg(a.shape)
    body = struct.pack("<BBHHH", 1, 0, len(nm), len(dt), len(ds))
    for part in (nm, dt, ds):
        body += part + b"\0" * (_pad8(len(part)) - len(part))
    body += a.tobytes()
    return _message(0x000C, body)


class _Writer:
    def __init__(self):
        self.buf = bytearray(96)           # superblock v0 placeholder

    def alloc(self, data, align=8):
        pad = (-len(self.buf)) % align
        self.buf += b"\0" * pad
        addr = len(self.buf)
        self.buf += data
        return addr

    def object_header(self, messages):
        body = b"".join(messages)
        # version 1, reserved, #messages, reference count 1, size of the message block; 4 bytes of padding put the first message on
        # an 8-byte boundary (FFS IV.A.1.a)
        return self.alloc(struct.pack("<BBHII4x", 1, 0, len(messages), 1, len(body)) + body)

    def dataset(self, arr, attrs):
        arr = _contig(arr)
        if arr.dtype.byteorder == ">":
            arr = arr.astype(arr.dtype.newbyteorder("<"))
        raw = arr.tobytes()
        data_addr = self.alloc(raw) if raw else UNDEF
        msgs = [_message(0x0001, _dataspace_msg(arr.shape)),
                _message(0x0003, _datatype_msg(arr.dtype), flags=1),
                # fill value v2: allocation time late (2), write time if-set (2), defined (1), size 0
                _message(0x0005, struct.pack("<BBBBI", 2, 2, 2, 1, 0), flags=1),
                # data layout v3, class 1 = contiguous: address, size
                _message(0x0008, struct.pack("<BBQQ", 3, 1, data_addr, len(raw)))]
        msgs += [_attribute_msg(k, v) for k, v in attrs.items()]
        return self.object_header(msgs)

    def group(self, spec):
        """old-style group: local heap with the link names, symbol-table nodes of <= 2 LEAF_K sorted entries under one level-0
        B-tree node; returns (object header address, B-tree address, heap address)"""
        child_addr = {}
        child_cache = {}
        for name, child in spec.children.items():
            if isinstance(child, GroupSpec):
                oh, bt, hp = self.group(child)
                child_addr[name], child_cache[name] = oh, (bt, hp)
            else:
                attrs = child.attrs if isinstance(child, DatasetSpec) else {}
                data = child.data if isinstance(child, DatasetSpec) else child
                child_addr[name] = self.dataset(data, attrs)
        names = sorted(child_addr, key=lambda s: s.encode("utf8"))       # strcmp order
        # local heap data segment: offset 0 holds the empty string (key 0 of the B-tree), names null-terminated and 8-byte aligned
        seg = bytearray(8)
        name_off = {}
        for n in names:
            name_off[n] = len(seg)
            b = n.encode("utf8") + b"\0"
            seg += b + b"\0" * (_pad8(len(b)) - len(b))
        free_off = len(seg)
        seg_size = max(_pad8(len(seg) + 16), 88)                          # libhdf5's minimum heap data segment is 88 bytes
        seg += struct.pack("<QQ", HEAP_FREE_NULL, seg_size - free_off)    # one free block: (next = end of list, size)
        seg += b"\0" * (seg_size - len(seg))
        seg_addr = self.alloc(bytes(seg))
        heap_addr = self.alloc(b"HEAP" + struct.pack("<B3xQQQ", 0, seg_size, free_off, seg_addr))
        # symbol table nodes
        per = 2 * LEAF_K
        chunks = [names[i:i + per] for i in range(0, len(names), per)]
        if len(chunks) > 2 * INTERNAL_K:
            raise ValueError(f"group with {len(names)} entries needs a multi-level B-tree (not written by this module)")
        snods = []
        for chunk in chunks:
            node = b"SNOD" + struct.pack("<BBH", 1, 0, len(chunk))
            for n in chunk:
                if n in child_cache:
                    node += struct.pack("<QQII", name_off[n], child_addr[n], 1, 0) + struct.pack("<QQ", *child_cache[n])
                else:
                    node += struct.pack("<QQII16x", name_off[n], child_addr[n], 0, 0)
            node += b"\0" * (8 + per * 40 - len(node))
            snods.append(self.alloc(node))
        # B-tree v1 node, type 0 (group), level 0: key[0] = "" (heap offset 0), key[i + 1] = largest name of child i
        tree = b"TREE" + struct.pack("<BBHQQ", 0, 0, len(chunks), UNDEF, UNDEF) + struct.pack("<Q", 0)
        for chunk, addr in zip(chunks, snods):
            tree += struct.pack("<QQ", addr, name_off[chunk[-1]])
        tree += b"\0" * (24 + (2 * INTERNAL_K + 1) * 8 + 2 * INTERNAL_K * 8 - len(tree))
        tree_addr = self.alloc(tree)
        msgs = [_message(0x0011, struct.pack("<QQ", tree_addr, heap_addr))]
        msgs += [_attribute_msg(k, v) for k, v in spec.attrs.items()]
        return self.object_header(msgs), tree_addr, heap_addr

    def finish(self, root):
        oh, bt, hp = self.group(root)
        self.buf += b"\0" * ((-len(self.buf)) % 8)
        sb = SIGNATURE + struct.pack("<BBBBBBBBHHI", 0, 0, 0, 0, 0, 8, 8, 0, LEAF_K, INTERNAL_K, 0)
        sb += struct.pack("<QQQQ", 0, UNDEF, len(self.buf), UNDEF)
        sb += struct.pack("<QQII", 0, oh, 1, 0) + struct.pack("<QQ", bt, hp)      # root symbol-table entry, cache type 1
        assert len(sb) == 96
        self.buf[:96] = sb
        return bytes(self.buf)


class DatasetSpec:
    def __init__(self, data, attrs=None):
        self.data = _contig(data)
        self.attrs = dict(attrs or {})


def dumps(root: GroupSpec) -> bytes:
    return _Writer().finish(root)


def write_file(path, root: GroupSpec):
    blob = dumps(root)
    with open(path, "wb") as f:
        f.write(blob)
    return path


# ====================================================================================================== reader

class _Reader:
    def __init__(self, blob):
        self.b = blob
        start = 0
        while self.b[start:start + 8] != SIGNATURE:      # the superblock may sit at 0, 512, 1024, ... (user block)
            start = 512 if start == 0 else start * 2
            if start + 8 > len(self.b):
                raise ValueError("not an HDF5 file (signature not found)")
        ver = self.b[start + 8]
        if ver > 1:
            raise NotImplementedError(f"superblock version {ver} (libver='latest' files) is not supported; re-save with the default libver")
        so, sl = self.b[start + 13], self.b[start + 14]
        if (so, sl) != (8, 8):
            raise NotImplementedError(f"size of offsets / lengths {so} / {sl} (only 8 / 8)")
        self.leaf_k, self.internal_k = struct.unpack_from("<HH", self.b, start + 16)
        p = start + 24 + (4 if ver == 1 else 0)
        self.base, _free, self.eof, _drv = struct.unpack_from("<QQQQ", self.b, p)
        if self.base == UNDEF:
            self.base = 0            # (every address in the file is relative to the base address)
        p += 32
        _name_off, self.root_header, _cache, _ = struct.unpack_from("<QQII", self.b, p)

    def at(self, addr):
        return self.base + addr

    # ---- object headers ----
    def messages(self, addr):
        """[(type, flags, body)] of a version-1 object header, following continuation messages"""
        p = self.at(addr)
        if self.b[p:p + 4] == b"OHDR":
            raise NotImplementedError("object header version 2 (libver='latest') is not supported")
        ver, _, nmsg, _refs, size = struct.unpack_from("<BBHII", self.b, p)
        if ver != 1:
            raise ValueError(f"object header version {ver} at {addr}")
        out = []
        blocks = [(p + 16, size)]
        while blocks and len(out) < nmsg:
            q, left = blocks.pop(0)
            end = q + left
            while q + 8 <= end and len(out) < nmsg:
                mtype, msize, flags = struct.unpack_from("<HHB", self.b, q)
                body = self.b[q + 8:q + 8 + msize]
                q += 8 + msize
                if mtype == 0x0010:
                    caddr, clen = struct.unpack_from("<QQ", body, 0)
                    blocks.append((self.at(caddr), clen))
                out.append((mtype, flags, body))
        return out

    # ---- groups ----
    def heap_string(self, heap_addr, off):
        p = self.at(heap_addr)
        if self.b[p:p + 4] != b"HEAP":
            raise ValueError("bad local heap signature")
        _size, _free, seg = struct.unpack_from("<QQQ", self.b, p + 8)
        s = self.at(seg) + off
        e = self.b.index(b"\0", s)
        return self.b[s:e].decode("utf8")

    def group_entries(self, btree_addr, heap_addr):
        """name -> object header address, in B-tree (= name) order"""
        out = {}
        p = self.at(btree_addr)
        if self.b[p:p + 4] != b"TREE":
            raise ValueError("bad B-tree signature")
        ntype, level, used = struct.unpack_from("<BBH", self.b, p + 4)
        if ntype != 0:
            raise ValueError("not a group B-tree")
        q = p + 24 + 8                                    # skip key 0
        for _ in range(used):
            child = struct.unpack_from("<Q", self.b, q)[0]
            q += 16                                       # child pointer + next key
            if level > 0:
                out.update(self.group_entries(child, heap_addr))
                continue
            s = self.at(child)
            if self.b[s:s + 4] != b"SNOD":
                raise ValueError("bad symbol table node signature")
            nsym = struct.unpack_from("<H", self.b, s + 6)[0]
            for i in range(nsym):
                name_off, oh = struct.unpack_from("<QQ", self.b, s + 8 + 40 * i)
                out[self.heap_string(heap_addr, name_off)] = oh
        return out

    # ---- datatypes / dataspaces / attributes ----
    def parse_datatype(self, body):
        """-> (numpy dtype | ('vlen_str',) , message length consumed)"""
        cls_ver, b0, b1, _b2, size = struct.unpack_from("<BBBBI", body, 0)
        cls = cls_ver & 0x0F
        if cls == 0:
            order = ">" if b0 & 1 else "<"
            kind = "i" if b0 & 0x08 else "u"
            return np.dtype(f"{order}{kind}{size}"), 12
        if cls == 1:
            order = ">" if b0 & 1 else "<"
            return np.dtype(f"{order}f{size}"), 20
        if cls == 3:
            return np.dtype(f"S{size}"), 8
        if cls == 9:
            if (b0 & 0x0F) != 1:
                raise NotImplementedError("variable-length sequences (only variable-length strings)")
            return ("vlen_str",), None
        raise NotImplementedError(f"datatype class {cls}")

    def parse_dataspace(self, body):
        ver, rank, flags = struct.unpack_from("<BBB", body, 0)
        if ver == 1:
            p = 8
        elif ver == 2:
            if body[3] == 2:        # null dataspace
                return None
            p = 4
        else:
            raise ValueError(f"dataspace version {ver}")
        return tuple(struct.unpack_from("<Q", body, p + 8 * i)[0] for i in range(rank))

    def global_heap_object(self, addr, index):
        p = self.at(addr)
        if self.b[p:p + 4] != b"GCOL":
            raise ValueError("bad global heap signature")
        size = struct.unpack_from("<Q", self.b, p + 8)[0]
        q, end = p + 16, p + size
        while q + 16 <= end:
            idx, _refs, osize = struct.unpack_from("<HH4xQ", self.b, q)
            if idx == index:
                return self.b[q + 16:q + 16 + osize]
            if idx == 0:
                break
            q += 16 + _pad8(osize)
        raise KeyError(f"global heap object {index}")

    def decode_values(self, dtype, shape, raw):
        n = int(np.prod(shape)) if shape else 1
        if isinstance(dtype, tuple):      # variable-length strings: (length u32, collection address u64, object index u32) each
            vals = []
            for i in range(n):
                ln, addr, idx = struct.unpack_from("<IQI", raw, 16 * i)
                vals.append(self.global_heap_object(addr, idx)[:ln] if ln else b"")
            arr = np.array(vals, dtype=object).reshape(shape if shape else ())
            return arr if shape else arr[()]
        arr = np.frombuffer(raw, dtype=dtype, count=n).reshape(shape if shape else ())
        if dtype.kind == "S":
            return arr.copy() if shape else bytes(arr[()])
        arr = arr.astype(dtype.newbyteorder("="))
        return arr if shape else arr[()]

    def parse_attribute(self, body, flags):
        if flags & 2:
            raise NotImplementedError("shared attribute messages")
        ver = body[0]
        if ver == 1:
            nlen, dlen, slen = struct.unpack_from("<HHH", body, 2)
            p = 8
            step = _pad8
        elif ver in (2, 3):
            if body[1] & 3:
                raise NotImplementedError("attribute with a shared datatype / dataspace")
            nlen, dlen, slen = struct.unpack_from("<HHH", body, 2)
            p = 8 + (1 if ver == 3 else 0)
            step = lambda n: n      # noqa: E731  (no padding in versions 2 and 3)
        else:
            raise ValueError(f"attribute message version {ver}")
        name = body[p:p + nlen].split(b"\0")[0].decode("utf8")
        p += step(nlen)
        dtype, _ = self.parse_datatype(body[p:p + dlen])
        p += step(dlen)
        shape = self.parse_dataspace(body[p:p + slen])
        p += step(slen)
        if shape is None:
            return name, None
        return name, self.decode_values(dtype, shape, body[p:])


class _Node:
    def __init__(self, reader, addr, name):
        self._r = reader
        self._addr = addr
        self.name = name
        self._msgs = reader.messages(addr)
        self.attrs = {}
        for mtype, flags, body in self._msgs:
            if mtype == 0x000C:
                k, v = reader.parse_attribute(body, flags)
                self.attrs[k] = v
            elif mtype == 0x0015:
                raise NotImplementedError("attributes in dense storage (attribute info message)")


class Dataset(_Node):
    def __init__(self, reader, addr, name):
        super().__init__(reader, addr, name)
        self.shape = self.dtype = None
        self._layout = None
        for mtype, flags, body in self._msgs:
            if mtype == 0x0001:
                self.shape = reader.parse_dataspace(body)
            elif mtype == 0x0003:
                if flags & 2:
                    raise NotImplementedError("shared (committed) datatype")
                self.dtype, _ = reader.parse_datatype(body)
            elif mtype == 0x0008:
                self._layout = body
            elif mtype == 0x000B:
                raise NotImplementedError(f"{name}: filtered (compressed) dataset")

    def read(self):
        body = self._layout
        if body is None or self.shape is None or self.dtype is None:
            raise ValueError(f"{self.name}: incomplete dataset header")
        if body[0] != 3:
            raise NotImplementedError(f"{self.name}: data layout message version {body[0]}")
        n = int(np.prod(self.shape)) if self.shape else 1
        itemsize = 16 if isinstance(self.dtype, tuple) else self.dtype.itemsize
        if body[1] == 0:                                   # compact: size (2), data
            size = struct.unpack_from("<H", body, 2)[0]
            raw = body[4:4 + size]
        elif body[1] == 1:                                 # contiguous: address (8), size (8)
            addr, size = struct.unpack_from("<QQ", body, 2)
            if addr == UNDEF:
                raw = b"\0" * (n * itemsize)               # never written: fill value 0
            else:
                raw = self._r.b[self._r.at(addr):self._r.at(addr) + size]
        else:
            raise NotImplementedError(f"{self.name}: chunked dataset (Keras weight files are contiguous)")
        if len(raw) < n * itemsize:
            raise ValueError(f"{self.name}: truncated data")
        return self._r.decode_values(self.dtype, self.shape, raw)

    def __array__(self, dtype=None, copy=None):
        a = self.read()
        return a.astype(dtype) if dtype is not None else a

    def __getitem__(self, idx):
        return self.read()[idx]


class Group(_Node):
    def __init__(self, reader, addr, name):
        super().__init__(reader, addr, name)
        self._entries = None
        for mtype, _flags, body in self._msgs:
            if mtype == 0x0011:
                bt, hp = struct.unpack_from("<QQ", body, 0)
                self._entries = reader.group_entries(bt, hp)
            elif mtype in (0x0002, 0x0006):
                raise NotImplementedError("new-style groups (link messages, libver='latest')")
        if self._entries is None:
            raise ValueError(f"{name}: not a group")

    def keys(self):
        return list(self._entries)

    def __contains__(self, name):
        try:
            self[name]
            return True
        except KeyError:
            return False

    def __iter__(self):
        return iter(self._entries)

    def __getitem__(self, path):
        node = self
        for part in [p for p in path.split("/") if p]:
            if not isinstance(node, Group) or part not in node._entries:
                raise KeyError(path)
            addr = node._entries[part]
            child_name = (node.name.rstrip("/") + "/" + part)
            is_group = any(m[0] == 0x0011 for m in node._r.messages(addr))
            node = (Group if is_group else Dataset)(node._r, addr, child_name)
        return node


class File(Group):
    """read-only view of an HDF5 file (the whole file is read into memory: weight files of this model family are < 40 MB)"""

    def __init__(self, path_or_bytes):
        if isinstance(path_or_bytes, (bytes, bytearray, memoryview)):
            blob = bytes(path_or_bytes)
        else:
            with open(path_or_bytes, "rb") as f:
                blob = f.read()
        r = _Reader(blob)
        super().__init__(r, r.root_header, "/")

    def __enter__(self):
        return self

    def __exit__(self, *a):
        return False


# ====================================================================================================== Keras weight files

def _names(v):
    out = []
    for n in np.asarray(v).reshape(-1):
        out.append(n.decode("utf8") if isinstance(n, (bytes, np.bytes_)) else str(n))
    return out


def save_keras_weights(path, layers, extra_groups=None, keras_version="2.2.4-tf", backend="tensorflow"):
    """The file `keras.Model.save_weights(path.h5)` writes (tensorflow/python/keras/saving/hdf5_format.py, save_weights_to_hdf5_group):
    root attributes `layer_names`, `backend`, `keras_version`; one group per layer with the attribute `weight_names` and one dataset per
    weight, named by the variable (`lg_vae/encoder/conv2d/kernel:0` - the slashes make nested groups).
    layers: [(layer name, [(weight name, array), ...]), ...];  extra_groups: {top-level group name: {dataset path: array}} for data
    Keras ignores on load (this build keeps the Adam state there)."""
    root = GroupSpec()
    root.attrs["layer_names"] = [n for n, _ in layers]
    root.attrs["backend"] = backend
    root.attrs["keras_version"] = keras_version
    for lname, weights in layers:
        g = root.require_group(lname)
        g.attrs["weight_names"] = [w for w, _ in weights] if weights else np.zeros((0,), dtype="S1")
        for wname, arr in weights:
            g.create_dataset(wname, np.asarray(arr))
    for gname, items in (extra_groups or {}).items():
        g = root.require_group(gname)
        for dname, arr in items.items():
            g.create_dataset(dname, np.asarray(arr))
    return write_file(path, root)


def load_keras_weights(path):
    """-> ({weight name: array} over all layers in `layer_names` order, File) - load_weights_from_hdf5_group's traversal.  Accepts a
    weights file and a full-model file (`model.save(...)`: the weights then sit under the `model_weights` group)."""
    f = File(path)
    top = f["model_weights"] if ("layer_names" not in f.attrs and "model_weights" in f.keys()) else f
    if "layer_names" not in top.attrs:
        raise KeyError(f"{path}: no `layer_names` attribute - not a Keras weights file")
    flat = {}
    for lname in _names(top.attrs["layer_names"]):
        g = top[lname]
        for wname in _names(g.attrs.get("weight_names", [])):
            flat[wname] = np.asarray(g[wname].read())
    return flat, f
