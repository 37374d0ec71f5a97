"""
TEST INFRASTRUCTURE ONLY -- never imported by the product (cluster_generator_b200/).

Independent pure-Python reader + structural validator for the HDF5 subset the reference's
particle files use (h5py calls at particles.py:239-363, 453-552).  h5py / libhdf5 are not
installed here (third-party dependency absent from /root/reference: h5py, unpinned in
setup.py; libhdf5 1.x), so this restates the published "HDF5 File Format Specification
Version 1.1/2.0" for the default ('earliest') layout: superblock v0/v1 (II.A), v1 group
B-trees (III.A.1), symbol-table nodes (III.C), local heaps (III.D), global heap collections
(III.E), v1 object headers (IV.A.1.a) and the dataspace / datatype / fill / layout /
attribute / continuation / symbol-table messages (IV.A.2).

Pinned against a file written by the real libhdf5: scipy ships
``scipy/io/matlab/tests/data/testhdf5_7.4_GLNX86.mat`` (MATLAB v7.3 = HDF5 behind a 512-byte
user block); tests/test_h5_cpu.py parses it with this module and with the native reader.

Besides decoding, ``load`` enforces the invariants libhdf5 checks when it opens a file
(end-of-file address, heap free-list sentinel, message counts and alignment, sorted symbol
tables, B-tree key bracketing), so a file that passes here is laid out the way libhdf5
expects, not merely the way our own reader tolerates.
"""
from __future__ import annotations

import struct

import numpy as np

UNDEF = 0xFFFFFFFFFFFFFFFF
SIG = b"\x89HDF\r\n\x1a\n"


class H5FormatError(Exception):
    pass


def _req(cond, msg):
    if not cond:
        raise H5FormatError(msg)


class _File:
    def __init__(self, data: bytes):
        self.d = data
        off = 0
        while True:
            if self.d[off:off + 8] == SIG:
                break
            off = 512 if off == 0 else off * 2
            _req(off + 8 <= len(self.d), "no HDF5 signature")
        self.sb = off
        ver = self.d[off + 8]
        _req(ver in (0, 1), f"superblock version {ver}")
        _req(self.d[off + 13] == 8 and self.d[off + 14] == 8, "offsets/lengths must be 8 bytes")
        self.leaf_k, self.internal_k = struct.unpack_from("<HH", self.d, off + 16)
        a = off + (24 if ver == 0 else 28)
        self.base, self.freespace, self.eof, self.driver = struct.unpack_from("<QQQQ", self.d, a)
        _req(self.freespace == UNDEF and self.driver == UNDEF, "free-space / driver blocks unsupported")
        # files behind a user block (base != 0) written by libhdf5 1.6 store the absolute end
        _req(self.eof <= len(self.d) if self.base else self.eof == len(self.d),
             "end-of-file address differs from the file size")
        name_off, self.root_hdr, cache, _, bt, hp = struct.unpack_from("<QQIIQQ", self.d, a + 32)
        self.root_cache = (cache, bt, hp)

    def at(self, addr, n):
        _req(addr != UNDEF and self.base + addr + n <= len(self.d), f"address {addr:#x}+{n} outside the file")
        return self.d[self.base + addr:self.base + addr + n]


def _decode_dtype(b):
    cls, ver = b[0] & 15, b[0] >> 4
    _req(ver in (1, 2, 3), "datatype version")
    size = struct.unpack_from("<I", b, 4)[0]
    if cls == 0:
        _req(not (b[1] & 1), "big-endian integers unsupported")
        off, prec = struct.unpack_from("<HH", b, 8)
        _req(off == 0 and prec == 8 * size, "padded integer")
        return np.dtype(("<i" if b[1] & 8 else "<u") + str(size)), 12
    if cls == 1:
        _req(not (b[1] & 1), "big-endian floats unsupported")
        off, prec, eloc, esz, mloc, msz, bias = struct.unpack_from("<HHBBBBI", b, 8)
        want = {4: (0, 32, 23, 8, 0, 23, 127), 8: (0, 64, 52, 11, 0, 52, 1023)}[size]
        _req((off, prec, eloc, esz, mloc, msz, bias) == want, "not an IEEE float")
        _req(b[2] == 8 * size - 1 and (b[1] >> 4) & 3 == 2, "float sign position / normalisation")
        return np.dtype("<f" + str(size)), 20
    if cls == 3:
        return ("fixed_str", size), 8
    if cls == 9:
        _req(b[1] & 15 == 1, "variable-length sequences unsupported")
        _req(size == 16, "vlen element size")
        _, n = _decode_dtype(b[8:])
        return ("vlen_str", (b[2] & 15)), 8 + n
    if cls == 5:
        return ("opaque", size), 8
    raise H5FormatError(f"datatype class {cls}")


def _decode_space(b):
    ver, rank, flags = b[0], b[1], b[2]
    if ver == 1:
        off = 8
    else:
        _req(ver == 2, "dataspace version")
        off = 4
    dims = struct.unpack_from("<" + "Q" * rank, b, off)
    n = off + 8 * rank * (2 if flags & 1 else 1)
    return tuple(dims), n


def _vlen_string(f, elem):
    slen, addr, idx = struct.unpack_from("<IQI", elem, 0)
    if slen == 0:
        return ""
    h = f.at(addr, 16)
    _req(h[:4] == b"GCOL" and h[4] == 1, "bad global heap collection")
    size = struct.unpack_from("<Q", h, 8)[0]
    _req(size >= 4096, "global heap collection below the minimum size")
    c = f.at(addr, size)
    o, seen0 = 16, False
    found = None
    while o + 16 <= size:
        oi, ref, _, osz = struct.unpack_from("<HHIQ", c, o)
        if oi == 0:
            _req(o + osz == size, "free-space object must span the rest of the collection")
            seen0 = True
            break
        if oi == idx:
            found = c[o + 16:o + 16 + osz][:slen]
        o += 16 + (osz + 7) // 8 * 8
    _req(seen0 or o == size or size - o < 16, "global heap collection not terminated")
    _req(found is not None, f"global heap object {idx} missing")
    return found.decode("utf-8")


def _decode_attr(f, b):
    ver = b[0]
    _req(ver in (1, 2, 3), "attribute version")
    nsz, tsz, ssz = struct.unpack_from("<HHH", b, 2)
    off = 9 if ver == 3 else 8
    adv = (lambda s: (s + 7) // 8 * 8) if ver == 1 else (lambda s: s)
    name = b[off:off + nsz]
    _req(name[-1] == 0, "attribute name not terminated")
    name = name[:-1].decode("utf-8")
    off += adv(nsz)
    dt, dn = _decode_dtype(b[off:off + tsz])
    _req(dn == tsz, "attribute datatype size field")
    off += adv(tsz)
    dims, sn = _decode_space(b[off:off + ssz])
    _req(sn == ssz, "attribute dataspace size field")
    off += adv(ssz)
    count = int(np.prod(dims)) if dims else 1
    if isinstance(dt, tuple):
        kind, arg = dt
        if kind == "fixed_str":
            raw = b[off:off + arg]
            return name, raw.split(b"\0")[0].decode("utf-8")
        if kind == "vlen_str":
            _req(count == 1, "string arrays unsupported")
            _req(len(b) >= off + 16, "attribute data truncated")
            return name, _vlen_string(f, b[off:off + 16])
        return name, None
    _req(len(b) >= off + count * dt.itemsize, "attribute data truncated")
    arr = np.frombuffer(b, dtype=dt, count=count, offset=off).copy()
    return name, (arr[0] if not dims else arr)


def _messages(f, addr, size, out, strict):
    blk = f.at(addr, size)
    o = 0
    while o + 8 <= size:
        mtype, msz, flags = struct.unpack_from("<HHB", blk, o)
        _req(msz % 8 == 0, "v1 message size not a multiple of 8")
        _req(o + 8 + msz <= size, "message crosses the end of its block")
        body = blk[o + 8:o + 8 + msz]
        o += 8 + msz
        out.append((mtype, flags, body))
        if mtype == 0x10:
            caddr, clen = struct.unpack_from("<QQ", body, 0)
            _messages(f, caddr, clen, out, strict)
    _req(o == size or not strict, "gap at the end of an object header block")


def _load_object(f, addr, depth=0):
    _req(depth < 32, "nesting too deep")
    h = f.at(addr, 16)
    _req(h[0] == 1, f"object header version {h[0]}")
    nmsgs, refcnt, hsize = struct.unpack_from("<HII", h, 2)
    _req(refcnt >= 1, "reference count")
    msgs = []
    _messages(f, addr + 16, hsize, msgs, strict=True)
    _req(len(msgs) == nmsgs, f"object header announces {nmsgs} messages, holds {len(msgs)}")
    node = {"attrs": {}}
    space = dtype = layout = None
    for mtype, flags, body in msgs:
        _req(not (flags & 2), "shared messages unsupported")
        if mtype == 0x01:
            space, _ = _decode_space(body)
        elif mtype == 0x03:
            dtype, _ = _decode_dtype(body)
        elif mtype == 0x05:
            _req(body[0] in (1, 2, 3), "fill value version")
        elif mtype == 0x08:
            ver = body[0]
            if ver in (1, 2):
                _req(body[2] == 1, "only contiguous layout")
                layout = (struct.unpack_from("<Q", body, 8)[0], None)
            else:
                _req(ver in (3, 4) and body[1] == 1, "only contiguous layout")
                layout = struct.unpack_from("<QQ", body, 2)
        elif mtype == 0x0C:
            k, v = _decode_attr(f, body)
            _req(k not in node["attrs"], f"duplicate attribute {k}")
            node["attrs"][k] = v
        elif mtype == 0x11:
            node["stab"] = struct.unpack_from("<QQ", body, 0)
    if "stab" in node:
        node["kind"] = "group"
        node["children"] = _load_group(f, *node.pop("stab"), depth)
        return node
    _req(space is not None and dtype is not None and layout is not None, "dataset lacks space/type/layout")
    _req(isinstance(dtype, np.dtype), "dataset type unsupported")
    count = int(np.prod(space)) if space else 1
    nbytes = count * dtype.itemsize
    addr_d, size_d = layout
    _req(size_d is None or size_d == nbytes, "layout size differs from dataspace x datatype")
    if nbytes == 0:
        data = np.empty(space, dtype=dtype)
    else:
        data = np.frombuffer(f.at(addr_d, nbytes), dtype=dtype).reshape(space).copy()
    node.update(kind="dataset", data=data, offset=None if nbytes == 0 else f.base + addr_d)
    return node


def _heap_name(f, seg_addr, seg_size, off):
    _req(off < seg_size, "name offset outside the local heap")
    seg = f.at(seg_addr, seg_size)
    end = seg.index(b"\0", off)
    return seg[off:end].decode("utf-8")


def _load_group(f, btree, heap, depth):
    h = f.at(heap, 32)
    _req(h[:4] == b"HEAP" and h[4] == 0, "bad local heap")
    seg_size, free_head, seg_addr = struct.unpack_from("<QQQ", h, 8)
    _req(seg_size % 8 == 0, "local heap data segment not 8-aligned")
    _req(free_head == 1 or free_head < seg_size, "local heap free-list head (H5HL_FREE_NULL = 1)")
    _req(_heap_name(f, seg_addr, seg_size, 0) == "", "heap offset 0 must hold the empty name")
    children = {}
    order = []

    def walk(addr, lo_key, hi_key, lvl_expected):
        t = f.at(addr, 24)
        _req(t[:4] == b"TREE" and t[4] == 0, "bad group B-tree node")
        level, used = t[5], struct.unpack_from("<H", t, 6)[0]
        _req(lvl_expected is None or level == lvl_expected, "B-tree level")
        _req(used <= 2 * f.internal_k, "B-tree node over-full")
        body = f.at(addr + 24, (2 * f.internal_k + 1) * 8 + 2 * f.internal_k * 8)  # full node is on disk
        keys = [struct.unpack_from("<Q", body, 16 * i)[0] for i in range(used + 1)]
        kids = [struct.unpack_from("<Q", body, 16 * i + 8)[0] for i in range(used)]
        knames = [_heap_name(f, seg_addr, seg_size, k) for k in keys]
        if lo_key is not None:
            _req(knames[0] == lo_key and knames[-1] == hi_key, "child keys do not match the parent's")
        for i, c in enumerate(kids):
            if level > 0:
                walk(c, knames[i], knames[i + 1], level - 1)
                continue
            s = f.at(c, 8 + 2 * f.leaf_k * 40)
            _req(s[:4] == b"SNOD" and s[4] == 1, "bad symbol-table node")
            ns = struct.unpack_from("<H", s, 6)[0]
            _req(1 <= ns <= 2 * f.leaf_k, "symbol-table node entry count")
            names = []
            for e in range(ns):
                noff, oaddr, cache, _, sp0, sp1 = struct.unpack_from("<QQIIQQ", s, 8 + 40 * e)
                nm = _heap_name(f, seg_addr, seg_size, noff)
                names.append(nm)
                child = _load_object(f, oaddr, depth + 1)
                if cache == 1:
                    _req(child["kind"] == "group", "cached symbol-table info on a non-group")
                else:
                    _req(cache == 0, "symbol-table cache type")
                children[nm] = child
                order.append(nm)
            enc = [n.encode() for n in names]
            _req(enc == sorted(enc) and len(set(enc)) == len(enc), "symbol-table node not sorted")
            _req(knames[i].encode() < enc[0] and enc[-1] <= knames[i + 1].encode(),
                 "B-tree keys do not bracket the node")

    walk(btree, None, None, None)
    enc = [n.encode() for n in order]
    _req(enc == sorted(enc), "group entries not in strcmp order")
    return children


def load(path):
    """Parse + validate; returns the root group as nested dicts:
    {"kind": "group", "attrs": {...}, "children": {name: node}} /
    {"kind": "dataset", "attrs": {...}, "data": ndarray, "offset": int}."""
    with open(path, "rb") as fh:
        f = _File(fh.read())
    root = _load_object(f, f.root_hdr)
    _req(root["kind"] == "group", "root object is not a group")
    root["eof"] = f.eof
    return root
