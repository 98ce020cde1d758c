"""Dependency-free HDF5 subset: a reader (superblock v0/v1, object header v1, contiguous / compact layout) and a writer that
produces files the HDF5 library (h5py) opens -- enough for the two files this package exchanges with jQMC:

* TREXIO inputs written with the oldest on-disk format (``read``), and
* jQMC's restart checkpoints (``jqmc/_checkpoint.py``: groups, contiguous numeric / fixed-length-string datasets, scalar
  attributes), which the replacement drivers write and read through the small h5py-like facade ``File`` below.

When ``h5py`` is importable the drivers use it instead (``open_file``); this module is what runs where it is not (the build
image and the GPU box have no h5py).  Unsupported features raise NotImplementedError -- never a silent mis-read.

On-disk structures follow the HDF5 File Format Specification, version 1.1 layout: superblock v0, version-1 object headers,
version-1 group B-trees with local heaps and symbol-table nodes, version-1 dataspace / attribute messages, version-3
contiguous data layout.
"""

from __future__ import annotations

import struct

import numpy as np

_UNDEF = 0xFFFFFFFFFFFFFFFF


class MiniHDF5:
    def __init__(self, path: str):
        with open(path, "rb") as f:
            self.buf = f.read()
        b = self.buf
        if b[:8] != b"\x89HDF\r\n\x1a\n":
            raise ValueError("not an HDF5 file")
        ver = b[8]
        if ver not in (0, 1):
            raise NotImplementedError(f"superblock version {ver}")
        self.O = b[13]
        self.L = b[14]
        if self.O != 8 or self.L != 8:
            raise NotImplementedError("only 8-byte offsets/lengths")
        p = 24 if ver == 0 else 28
        # base, freespace, eof, driver
        self.base = self._u64(p)
        p += 32
        # root symbol table entry
        self.root = self._read_symbol_entry(p)
        self._gheap_cache = {}

    # -- primitive readers -------------------------------------------------
    def _u16(self, p):
        return struct.unpack_from("<H", self.buf, p)[0]

    def _u32(self, p):
        return struct.unpack_from("<I", self.buf, p)[0]

    def _u64(self, p):
        return struct.unpack_from("<Q", self.buf, p)[0]

    def _read_symbol_entry(self, p):
        name_off = self._u64(p)
        ohdr = self._u64(p + 8)
        cache = self._u32(p + 16)
        btree = heap = None
        if cache == 1:
            btree = self._u64(p + 24)
            heap = self._u64(p + 32)
        return dict(name_off=name_off, ohdr=ohdr, cache=cache, btree=btree, heap=heap)

    # -- object headers ----------------------------------------------------
    def _messages(self, addr):
        b = self.buf
        if b[addr] != 1:
            raise NotImplementedError(f"object header version {b[addr]}")
        nmsg = self._u16(addr + 2)
        size = self._u32(addr + 8)
        blocks = [(addr + 16, size)]
        msgs = []
        while blocks and len(msgs) < nmsg:
            p, sz = blocks.pop(0)
            end = p + sz
            while p + 8 <= end and len(msgs) < nmsg:
                mtype = self._u16(p)
                msize = self._u16(p + 2)
                data = p + 8
                if mtype == 0x10:  # continuation
                    blocks.append((self._u64(data), self._u64(data + 8)))
                msgs.append((mtype, data, msize))
                p = data + msize
        return msgs

    def _group_children(self, ohdr):
        btree = heap = None
        for mtype, data, _ in self._messages(ohdr):
            if mtype == 0x11:
                btree = self._u64(data)
                heap = self._u64(data + 8)
        if btree is None:
            return None
        if self.buf[heap : heap + 4] != b"HEAP":
            raise ValueError("bad local heap")
        heap_data = self._u64(heap + 24)
        out = {}
        self._walk_btree(btree, heap_data, out)
        return out

    def _walk_btree(self, addr, heap_data, out):
        b = self.buf
        if b[addr : addr + 4] != b"TREE":
            raise ValueError("bad btree node")
        level = b[addr + 5]
        n = self._u16(addr + 6)
        p = addr + 8 + 16
        for i in range(n):
            child = self._u64(p + 8)  # key_i at p, child_i at p+8
            p += 16
            if level > 0:
                self._walk_btree(child, heap_data, out)
            else:
                if b[child : child + 4] != b"SNOD":
                    raise ValueError("bad symbol node")
                ns = self._u16(child + 6)
                q = child + 8
                for _ in range(ns):
                    ent = self._read_symbol_entry(q)
                    s = heap_data + ent["name_off"]
                    e = b.index(b"\x00", s)
                    out[b[s:e].decode()] = ent["ohdr"]
                    q += 40

    # -- public API --------------------------------------------------------
    def listdir(self, path="/"):
        ohdr = self._resolve(path)
        ch = self._group_children(ohdr)
        if ch is None:
            raise KeyError(f"{path} is not a group")
        return sorted(ch)

    def _resolve(self, path):
        ohdr = self.root["ohdr"]
        for part in [p for p in path.split("/") if p]:
            ch = self._group_children(ohdr)
            if ch is None or part not in ch:
                raise KeyError(path)
            ohdr = ch[part]
        return ohdr

    def has(self, path):
        try:
            self._resolve(path)
            return True
        except KeyError:
            return False

    def read(self, path):
        ohdr = self._resolve(path)
        shape = None
        dt = None
        layout = None
        for mtype, data, msize in self._messages(ohdr):
            if mtype == 0x01:
                shape = self._dataspace(data)
            elif mtype == 0x03:
                dt = self._datatype(data)
            elif mtype == 0x08:
                layout = self._layout(data)
            elif mtype == 0x0B:
                raise NotImplementedError("filtered dataset")
        if shape is None or dt is None or layout is None:
            raise KeyError(f"{path} is not a dataset")
        count = int(np.prod(shape)) if len(shape) else 1
        kind, info = dt
        if kind == "np":
            nbytes = count * info.itemsize
            raw = self._raw(layout, nbytes)
            arr = np.frombuffer(raw, dtype=info, count=count).reshape(shape)
            return arr.copy()
        if kind == "str":
            raw = self._raw(layout, count * info)
            vals = [raw[i * info : (i + 1) * info].split(b"\x00")[0].decode() for i in range(count)]
            return vals if len(shape) else vals[0]
        if kind == "vlen_str":
            raw = self._raw(layout, count * 16)
            vals = []
            for i in range(count):
                ln, gaddr, gidx = struct.unpack_from("<IQI", raw, i * 16)
                vals.append(self._gheap_obj(gaddr, gidx)[:ln].split(b"\x00")[0].decode())
            return vals if len(shape) else vals[0]
        raise NotImplementedError(kind)

    def _raw(self, layout, nbytes):
        kind, a, sz = layout
        if kind == "compact":
            return self.buf[a : a + nbytes]
        if a == _UNDEF:
            return b"\x00" * nbytes
        a += self.base
        return self.buf[a : a + nbytes]

    def _dataspace(self, p):
        b = self.buf
        ver, rank, flags = b[p], b[p + 1], b[p + 2]
        if ver == 1:
            q = p + 8
        elif ver == 2:
            q = p + 4
        else:
            raise NotImplementedError("dataspace version")
        return tuple(self._u64(q + 8 * i) for i in range(rank))

    def _datatype(self, p):
        b = self.buf
        cls = b[p] & 0x0F
        bits0 = b[p + 1]
        size = self._u32(p + 4)
        if cls == 0:
            signed = bool(bits0 & 0x08)
            if bits0 & 1:
                raise NotImplementedError("big-endian")
            return ("np", np.dtype(("<i" if signed else "<u") + str(size)))
        if cls == 1:
            if bits0 & 1:
                raise NotImplementedError("big-endian")
            return ("np", np.dtype("<f" + str(size)))
        if cls == 3:
            return ("str", size)
        if cls == 9:
            vtype = bits0 & 0x0F
            if vtype == 1:
                return ("vlen_str", None)
            raise NotImplementedError("vlen sequence")
        raise NotImplementedError(f"datatype class {cls}")

    def _layout(self, p):
        b = self.buf
        ver = b[p]
        if ver == 3:
            cls = b[p + 1]
            if cls == 1:
                return ("contiguous", self._u64(p + 2), self._u64(p + 10))
            if cls == 0:
                sz = self._u16(p + 2)
                return ("compact", p + 4, sz)
            raise NotImplementedError("chunked layout")
        if ver in (1, 2):
            rank = b[p + 1]
            cls = b[p + 2]
            if cls == 1:
                return ("contiguous", self._u64(p + 8), 0)
            raise NotImplementedError("layout v1/2 non-contiguous")
        raise NotImplementedError("layout version")

    def _gheap_obj(self, addr, idx):
        if addr not in self._gheap_cache:
            b = self.buf
            a = addr + self.base
            if b[a : a + 4] != b"GCOL":
                raise ValueError("bad global heap")
            size = self._u64(a + 8)
            objs = {}
            p = a + 16
            end = a + size
            while p + 16 <= end:
                oid = self._u16(p)
                osz = self._u64(p + 8)
                if oid == 0:
                    break
                objs[oid] = b[p + 16 : p + 16 + osz]
                p += 16 + ((osz + 7) // 8) * 8
            self._gheap_cache[addr] = objs
        return self._gheap_cache[addr][idx]

    def attrs(self, path):
        """Return {name: value} of the attributes attached to a group or dataset."""
        ohdr = self._resolve(path)
        out = {}
        b = self.buf
        for mtype, p, msize in self._messages(ohdr):
            if mtype != 0x0C:
                continue
            ver = b[p]
            nsz, dsz, ssz = self._u16(p + 2), self._u16(p + 4), self._u16(p + 6)
            if ver == 1:
                q = p + 8
                pad = lambda n: (n + 7) // 8 * 8
            elif ver in (2, 3):
                q = p + 8 + (1 if ver == 3 else 0)
                pad = lambda n: n
            else:
                raise NotImplementedError("attribute version")
            name = b[q : q + nsz].split(b"\x00")[0].decode()
            q += pad(nsz)
            kind, info = self._datatype(q)
            q += pad(dsz)
            shape = self._dataspace(q) if ssz >= 4 else ()
            q += pad(ssz)
            count = int(np.prod(shape)) if len(shape) else 1
            if kind == "np":
                v = np.frombuffer(b, dtype=info, count=count, offset=q).reshape(shape).copy()
                out[name] = v if len(shape) else v.reshape(()).item()
            elif kind == "str":
                out[name] = b[q : q + info].split(b"\x00")[0].decode()
            elif kind == "vlen_str":
                ln, gaddr, gidx = struct.unpack_from("<IQI", b, q)
                out[name] = self._gheap_obj(gaddr, gidx)[:ln].decode()
        return out

    def walk(self, path="/"):
        """Yield (path, is_group) for everything below path."""
        for name in self.listdir(path):
            full = path.rstrip("/") + "/" + name
            ch = self._group_children(self._resolve(full))
            if ch is None:
                yield full, False
            else:
                yield full, True
                yield from self.walk(full)


# =====================================================================================================================
# Writer
# =====================================================================================================================
_LEAF_K = 32  # symbol-table node capacity 2K = 64 entries (declared in the superblock, honoured by the HDF5 library)
_INT_K = 16  # group B-tree node capacity 2K = 32 children


def _pad8(b: bytes) -> bytes:
    return b + b"\x00" * (-len(b) % 8)


def _dt_message(dt) -> bytes:
    """Datatype message body for a numpy dtype (little-endian fixed point / IEEE float / fixed-length string)."""
    dt = np.dtype(dt)
    if dt.kind == "b":
        dt = np.dtype("<i1")
    if dt.kind in "iu":
        bits0 = 0x08 if dt.kind == "i" else 0x00
        return struct.pack("<BBBBIHH", 0x10, bits0, 0, 0, dt.itemsize, 0, 8 * dt.itemsize)
    if dt.kind == "f":
        if dt.itemsize == 8:
            return struct.pack("<BBBBIHHBBBBI", 0x11, 0x20, 63, 0, 8, 0, 64, 52, 11, 0, 52, 1023)
        if dt.itemsize == 4:
            return struct.pack("<BBBBIHHBBBBI", 0x11, 0x20, 31, 0, 4, 0, 32, 23, 8, 0, 23, 127)
        raise NotImplementedError(f"float{8 * dt.itemsize}")
    if dt.kind == "S":
        return struct.pack("<BBBBI", 0x13, 0x11, 0, 0, max(1, dt.itemsize))  # null-padded, UTF-8
    raise NotImplementedError(f"dtype {dt}")


def _ds_message(shape) -> bytes:
    """Dataspace message body, version 1 (scalar: rank 0)."""
    return struct.pack("<BBBBI", 1, len(shape), 0, 0, 0) + b"".join(struct.pack("<Q", int(n)) for n in shape)


def _as_array(value):
    """numpy array (numeric, bool -> int8, str -> fixed-length UTF-8 bytes) of anything the checkpoint stores."""
    if isinstance(value, str):
        return np.array(value.encode("utf-8"))
    if isinstance(value, bytes):
        return np.array(value)
    a = np.asarray(value)
    if a.dtype.kind == "U":
        a = np.char.encode(a, "utf-8")
    elif a.dtype.kind == "O":
        a = np.array([str(x).encode("utf-8") for x in a.reshape(-1)]).reshape(a.shape)
    if a.dtype.kind == "b":
        a = a.astype(np.int8)
    if a.dtype.kind == "S" and a.dtype.itemsize == 0:
        a = a.astype("S1")
    if a.dtype.kind not in "iufS":
        raise NotImplementedError(f"cannot store dtype {a.dtype}")
    shape = a.shape  # (np.ascontiguousarray would turn a scalar into a 1-element vector)
    a = a.astype(a.dtype.newbyteorder("<")) if a.dtype.kind in "iuf" else a
    return np.ascontiguousarray(a).reshape(shape)


def _attr_message(name: str, value) -> bytes:
    a = _as_array(value)
    nm = name.encode("utf-8") + b"\x00"
    dt, ds = _dt_message(a.dtype), _ds_message(a.shape)
    return struct.pack("<BBHHH", 1, 0, len(nm), len(dt), len(ds)) + _pad8(nm) + _pad8(dt) + _pad8(ds) + a.tobytes()


def _object_header(messages) -> bytes:
    """Version-1 object header: messages = [(type, body)]; every body is padded to a multiple of 8 bytes."""
    body = b"".join(struct.pack("<HHBBBB", t, len(_pad8(m)), 0, 0, 0, 0) + _pad8(m) for t, m in messages)
    return struct.pack("<BBHII", 1, 0, len(messages), 1, len(body)) + b"\x00" * 4 + body


class _Node:
    def __init__(self):
        self.attrs = {}


class GroupW(_Node):
    """Group being built (h5py-like subset: create_group, create_dataset, attrs, require_group)."""

    def __init__(self):
        super().__init__()
        self.children = {}

    def create_group(self, name):
        node = self
        for part in [p for p in name.split("/") if p]:
            nxt = node.children.get(part)
            if nxt is None:
                nxt = node.children[part] = GroupW()
            node = nxt
        return node

    require_group = create_group

    def create_dataset(self, name, data=None):
        d = DatasetW(_as_array(data))
        self.children[name] = d
        return d

    def __contains__(self, name):
        return name in self.children

    def __getitem__(self, name):
        node = self
        for part in [p for p in name.split("/") if p]:
            node = node.children[part]
        return node

    def keys(self):
        return self.children.keys()


class DatasetW(_Node):
    def __init__(self, a):
        super().__init__()
        self.a = a


class _Serializer:
    def __init__(self):
        self.buf = bytearray(96)  # superblock (56) + root symbol-table entry (40), filled in last

    def alloc(self, data: bytes) -> int:
        addr = len(self.buf)
        self.buf += _pad8(data)
        return addr

    def put_dataset(self, d: DatasetW) -> int:
        raw = d.a.tobytes()
        data_addr = self.alloc(raw) if raw else _UNDEF
        msgs = [(0x01, _ds_message(d.a.shape)), (0x03, _dt_message(d.a.dtype)),
                (0x08, struct.pack("<BBQQ", 3, 1, data_addr, len(raw)))]  # fmt: skip
        msgs += [(0x0C, _attr_message(k, v)) for k, v in d.attrs.items()]
        return self.alloc(_object_header(msgs))

    def put_group(self, g: GroupW):
        """Returns (object header address, B-tree address, heap address)."""
        entries = []  # (name, header address, cache type, scratch)
        for name in sorted(g.children, key=lambda s: s.encode("utf-8")):  # symbol-table nodes are ordered by strcmp
            c = g.children[name]
            if isinstance(c, GroupW):
                oh, bt, hp = self.put_group(c)
                entries.append((name, oh, 1, struct.pack("<QQ", bt, hp)))
            else:
                entries.append((name, self.put_dataset(c), 0, b"\x00" * 16))
        # local heap: offset 0 holds the empty string (key 0 of the B-tree)
        heap_data = bytearray(8)
        offs = {}
        for name, *_ in entries:
            offs[name] = len(heap_data)
            heap_data += _pad8(name.encode("utf-8") + b"\x00")
        # one free block at the end, as the library leaves it: {offset of the next free block (1 = end of list), block size}
        free_off = len(heap_data)
        heap_data += struct.pack("<QQ", 1, 32) + b"\x00" * 16
        heap_data_addr = self.alloc(bytes(heap_data))
        heap_addr = self.alloc(b"HEAP" + struct.pack("<BBBBQQQ", 0, 0, 0, 0, len(heap_data), free_off, heap_data_addr))
        # symbol-table nodes of up to 2 K_leaf entries, one B-tree leaf node over them
        cap = 2 * _LEAF_K
        chunks = [entries[i : i + cap] for i in range(0, len(entries), cap)] or [[]]
        if len(chunks) > 2 * _INT_K:
            raise NotImplementedError("more than 2048 links in one group")
        snods = []
        for ch in chunks:
            body = b"SNOD" + struct.pack("<BBH", 1, 0, len(ch))
            for name, oh, cache, scratch in ch:
                body += struct.pack("<QQII", offs[name], oh, cache, 0) + scratch
            body += b"\x00" * (40 * (cap - len(ch)))
            snods.append((self.alloc(body), offs[ch[-1][0]] if ch else 0))
        tree = b"TREE" + struct.pack("<BBHQQ", 0, 0, len(snods) if entries else 0, _UNDEF, _UNDEF) + struct.pack("<Q", 0)
        for addr, last_key in snods if entries else []:
            tree += struct.pack("<QQ", addr, last_key)
        tree += b"\x00" * (24 + 8 + 16 * 2 * _INT_K - len(tree))
        bt_addr = self.alloc(tree)
        msgs = [(0x11, struct.pack("<QQ", bt_addr, heap_addr))] + [(0x0C, _attr_message(k, v)) for k, v in g.attrs.items()]
        return self.alloc(_object_header(msgs)), bt_addr, heap_addr

    def finish(self, root: GroupW) -> bytes:
        oh, bt, hp = self.put_group(root)
        sb = b"\x89HDF\r\n\x1a\n" + struct.pack("<BBBBBBBBHHI", 0, 0, 0, 0, 0, 8, 8, 0, _LEAF_K, _INT_K, 0)
        sb += struct.pack("<QQQQ", 0, _UNDEF, len(self.buf), _UNDEF)
        sb += struct.pack("<QQII", 0, oh, 1, 0) + struct.pack("<QQ", bt, hp)
        assert len(sb) == 96
        self.buf[:96] = sb
        return bytes(self.buf)


def write_file(path: str, root: GroupW) -> None:
    data = _Serializer().finish(root)
    with open(path, "wb") as f:
        f.write(data)


# =====================================================================================================================
# h5py-like facade over the reader / writer (only what the checkpoint code uses)
# =====================================================================================================================
class _AttrsR(dict):
    def get(self, k, default=None):
        return dict.get(self, k, default)


class GroupR:
    """Read-side group: keys(), `in`, [] (sub-group or dataset), .attrs."""

    def __init__(self, f: "MiniHDF5", path: str):
        self._f, self._path = f, path.rstrip("/") or "/"
        self.attrs = _AttrsR(f.attrs(self._path))

    def keys(self):
        return self._f.listdir(self._path)

    def __iter__(self):
        return iter(self.keys())

    def __contains__(self, name):
        return self._f.has(self._path.rstrip("/") + "/" + name)

    def __getitem__(self, name):
        full = self._path.rstrip("/") + "/" + name.strip("/")
        if not self._f.has(full):
            raise KeyError(full)
        if self._f._group_children(self._f._resolve(full)) is None:
            return DatasetR(self._f, full)
        return GroupR(self._f, full)


class DatasetR:
    def __init__(self, f, path):
        self._f, self._path = f, path
        self.attrs = _AttrsR(f.attrs(path))

    def __getitem__(self, idx):
        v = self._f.read(self._path)
        if isinstance(v, np.ndarray) and v.shape == ():
            v = v[()]
        if idx == () or idx is Ellipsis:
            return v
        return v[idx]


class File:
    """``with File(path, "w") as f: f.create_group(...)`` / ``with File(path, "r") as f: f["grp"]["x"][()]``."""

    def __init__(self, path, mode="r"):
        self._path, self._mode = path, mode
        if mode == "w":
            self._root = GroupW()
        elif mode == "r":
            self._root = GroupR(MiniHDF5(path), "/")
        else:
            raise ValueError("mode must be 'r' or 'w'")

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()
        return False

    def close(self):
        if self._mode == "w" and self._root is not None:
            write_file(self._path, self._root)
            self._root = None

    def __getattr__(self, name):
        return getattr(self._root, name)

    def __getitem__(self, name):
        return self._root[name]

    def __contains__(self, name):
        return name in self._root

    def copy_group(self, src: GroupR, name: str):
        """Deep copy of a read-side group into this (write-mode) file under `name` (h5py: out.copy(tmp, name))."""
        dst = self._root.create_group(name)

        def rec(s, d):
            for k, v in s.attrs.items():
                d.attrs[k] = v
            for k in s.keys():
                item = s[k]
                if isinstance(item, GroupR):
                    rec(item, d.create_group(k))
                else:
                    ds = d.create_dataset(k, data=_restore_strings(item[()]))
                    for ak, av in item.attrs.items():
                        ds.attrs[ak] = av

        rec(src, dst)
        return dst


def _restore_strings(v):
    if isinstance(v, list):
        return np.array([s.encode("utf-8") for s in v])
    return v


def open_file(path, mode="r"):
    """h5py.File when h5py is installed (byte-compatible with what jQMC writes), else this module's File."""
    try:
        import h5py

        return h5py.File(path, mode)
    except ImportError:
        return File(path, mode)
