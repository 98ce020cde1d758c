"""Minimal read-only HDF5 parser (superblock v0/v1, object header v1, contiguous/compact layout).

Test/fixture infrastructure only: h5py and trexio are not installed in the build image, and the
TREXIO files shipped with the jQMC reference (tests/trexio_example_files/*.h5,
benchmarks/water_ccecp_ccpvqz.h5) are plain, uncompressed, contiguous HDF5 written with the
oldest on-disk format, so ~200 lines of struct unpacking are enough to read them.

Supported: groups via symbol tables (B-tree v1 + local heap), datasets of fixed-point / IEEE
float / fixed-length string / variable-length string, contiguous and compact layouts.
Unsupported features raise NotImplementedError (never silently mis-read).
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
