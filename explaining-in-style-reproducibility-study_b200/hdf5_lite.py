"""A minimal HDF5 writer / reader for ``style_change_records.hdf5`` (reference notebook NB:394-417, cell 12) for images that
have no h5py (this one): flat files whose root group holds contiguous little-endian numeric datasets -- exactly what
``h5py.File(...).create_dataset(name, shape, dtype='f')`` produces for the nine AttFind datasets.

Format subset (HDF5 File Format Specification, the "classic" structures h5py / libhdf5 write by default):
superblock version 0 -> root symbol-table entry -> version-1 object headers -> group B-tree (version 1, "TREE") + local heap
("HEAP") + symbol-table node ("SNOD") -> per dataset an object header with dataspace (v1), datatype (fixed / floating point),
fill-value (v2) and data-layout (v3, contiguous or compact) messages.  No chunking, compression, attributes or nested groups
on the write side; the reader skips messages it does not need and follows header continuation blocks.

Pinning: h5py is absent here, so the reader is checked against a GENUINE HDF5 file that ships with scipy's test data
(a MATLAB 7.3 file written by libhdf5: same superblock / B-tree / heap / object-header versions), and the writer against
the reader plus a field-by-field comparison of the structures both files share (tests/test_host_cpu.py).  When h5py is
importable ``attfind.save_records`` uses it instead.
"""
from __future__ import annotations

import struct
from typing import Dict, List, Tuple

import numpy as np

SIGNATURE = b"\x89HDF\r\n\x1a\n"
UNDEF = 0xFFFFFFFFFFFFFFFF
LEAF_K = 8            # symbol-table node holds up to 2 * LEAF_K entries (the library default is 4)
INTERNAL_K = 16
FREE_NULL = 1         # H5HL_FREE_NULL: end of the local heap's free list


def _pad8(n: int) -> int:
    return (n + 7) & ~7


# ---------------------------------------------------------------------------------------------------------------------
# datatype message (0x0003)
# ---------------------------------------------------------------------------------------------------------------------
def _datatype_message(dt: np.dtype) -> bytes:
    dt = np.dtype(dt).newbyteorder("<")
    if dt.kind == "f" and dt.itemsize in (4, 8):
        exp_bits, man_bits, bias = (8, 23, 127) if dt.itemsize == 4 else (11, 52, 1023)
        bits = 8 * dt.itemsize
        # class 1 (floating point), version 1; bit field: little endian, mantissa normalisation 2 (implied msb), sign bit position
        head = struct.pack("<BBBBI", 0x11, 0x20, bits - 1, 0x00, dt.itemsize)
        return head + struct.pack("<HHBBBBI", 0, bits, man_bits, exp_bits, 0, man_bits, bias)
    if dt.kind in "iu" and dt.itemsize in (1, 2, 4, 8):
        head = struct.pack("<BBBBI", 0x10, 0x08 if dt.kind == "i" else 0x00, 0x00, 0x00, dt.itemsize)
        return head + struct.pack("<HH", 0, 8 * dt.itemsize)
    raise TypeError(f"hdf5_lite: unsupported dtype {dt}")


def _parse_datatype(data: bytes) -> np.dtype:
    cls, ver = data[0] & 0x0F, data[0] >> 4
    b0, size = data[1], struct.unpack("<I", data[4:8])[0]
    order = ">" if b0 & 1 else "<"
    if cls == 1:
        return np.dtype(f"{order}f{size}")
    if cls == 0:
        return np.dtype(f"{order}{'i' if b0 & 0x08 else 'u'}{size}")
    raise TypeError(f"hdf5_lite: datatype class {cls} (version {ver}) not supported")


# ---------------------------------------------------------------------------------------------------------------------
# writer
# ---------------------------------------------------------------------------------------------------------------------
def _message(mtype: int, data: bytes, flags: int = 0) -> bytes:
    data = data + b"\0" * (_pad8(len(data)) - len(data))
    return struct.pack("<HHBBBB", mtype, len(data), flags, 0, 0, 0) + data


def _object_header(messages: List[bytes]) -> bytes:
    body = b"".join(messages)
    return struct.pack("<BBHII", 1, 0, len(messages), 1, len(body)) + b"\0" * 4 + body


def write_hdf5(path: str, datasets: Dict[str, np.ndarray]) -> None:
    """Write ``{name: array}`` as contiguous datasets of the root group (names sorted, as the group B-tree requires)."""
    if not datasets:
        raise ValueError("hdf5_lite: nothing to write")
    if len(datasets) > 2 * LEAF_K:
        raise ValueError(f"hdf5_lite: at most {2 * LEAF_K} datasets in the root group")
    names = sorted(datasets, key=lambda s: s.encode())
    arrays = {}
    for n in names:
        if not n or "/" in n or "\0" in n:
            raise ValueError(f"hdf5_lite: bad dataset name {n!r}")
        a = np.asarray(datasets[n])
        a = a if a.ndim == 0 else np.ascontiguousarray(a)          # (ascontiguousarray would turn a scalar into shape (1,))
        arrays[n] = a.astype(a.dtype.newbyteorder("<"), copy=False)

    # local heap data segment: offset 0 = empty string, then the names (null terminated, 8-byte padded), then one free block
    heap = bytearray(8)
    name_off = {}
    for n in names:
        name_off[n] = len(heap)
        raw = n.encode() + b"\0"
        heap += raw + b"\0" * (_pad8(len(raw)) - len(raw))
    free_off = len(heap)
    heap_size = _pad8(max(len(heap) + 16, 128))
    heap += struct.pack("<QQ", FREE_NULL, heap_size - free_off) + b"\0" * (heap_size - free_off - 16)

    # fixed layout (addresses relative to the base address 0)
    sb_size = 56 + 40                                   # superblock v0 incl. the root symbol-table entry
    a_root = _pad8(sb_size)
    root_hdr = _object_header([_message(0x0011, struct.pack("<QQ", 0, 0), flags=1)])     # patched below
    a_heap = a_root + _pad8(len(root_hdr))
    a_heapdata = a_heap + 32
    a_btree = a_heapdata + heap_size
    btree_size = 24 + (2 * INTERNAL_K + 1) * 8 + 2 * INTERNAL_K * 8
    a_snod = a_btree + btree_size
    snod_size = 8 + 2 * LEAF_K * 40
    pos = a_snod + snod_size

    headers, a_hdr, a_data = {}, {}, {}
    for n in names:
        a = arrays[n]
        dims = a.shape if a.ndim else ()
        space = struct.pack("<BBBBI", 1, len(dims), 0, 0, 0) + b"".join(struct.pack("<Q", d) for d in dims)
        msgs = [_message(0x0001, space), _message(0x0003, _datatype_message(a.dtype), flags=1),
                _message(0x0005, struct.pack("<BBBB", 2, 2, 2, 0)),              # fill value v2: late alloc, write-if-set, undefined
                _message(0x0008, struct.pack("<BBQQ", 3, 1, 0, 0))]              # layout v3 contiguous (address patched below)
        headers[n] = msgs
        a_hdr[n] = pos
        pos += _pad8(len(_object_header(msgs)))
    for n in names:
        a_data[n] = pos if arrays[n].nbytes else UNDEF
        pos += _pad8(arrays[n].nbytes)
    eof = pos

    out = bytearray(eof)
    # superblock v0
    sb = SIGNATURE + struct.pack("<BBBBBBBBHHI", 0, 0, 0, 0, 0, 8, 8, 0, LEAF_K, INTERNAL_K, 0)
    sb += struct.pack("<QQQQ", 0, UNDEF, eof, UNDEF)
    sb += struct.pack("<QQII", 0, a_root, 1, 0) + struct.pack("<QQ", a_btree, a_heap)   # root entry, cached symbol table
    out[0:len(sb)] = sb
    root_hdr = _object_header([_message(0x0011, struct.pack("<QQ", a_btree, a_heap), flags=1)])
    out[a_root:a_root + len(root_hdr)] = root_hdr
    out[a_heap:a_heap + 32] = b"HEAP" + struct.pack("<BBBBQQQ", 0, 0, 0, 0, heap_size, free_off, a_heapdata)
    out[a_heapdata:a_heapdata + heap_size] = heap
    # group B-tree: one leaf-level node with one child (the symbol-table node); key0 = "", key1 = the largest name
    bt = b"TREE" + struct.pack("<BBHQQ", 0, 0, 1, UNDEF, UNDEF) + struct.pack("<QQQ", 0, a_snod, name_off[names[-1]])
    out[a_btree:a_btree + len(bt)] = bt
    sn = b"SNOD" + struct.pack("<BBH", 1, 0, len(names))
    for n in names:
        sn += struct.pack("<QQII", name_off[n], a_hdr[n], 0, 0) + b"\0" * 16
    out[a_snod:a_snod + len(sn)] = sn
    for n in names:
        msgs = headers[n][:3] + [_message(0x0008, struct.pack("<BBQQ", 3, 1, a_data[n], arrays[n].nbytes))]
        h = _object_header(msgs)
        out[a_hdr[n]:a_hdr[n] + len(h)] = h
        if arrays[n].nbytes:
            out[a_data[n]:a_data[n] + arrays[n].nbytes] = arrays[n].tobytes()
    with open(path, "wb") as f:
        f.write(bytes(out))


# ---------------------------------------------------------------------------------------------------------------------
# reader
# ---------------------------------------------------------------------------------------------------------------------
class _File:
    def __init__(self, buf: bytes):
        self.buf = buf
        off = 0
        while True:                                   # the superblock may sit behind a user block: 0, 512, 1024, ...
            if buf[off:off + 8] == SIGNATURE:
                break
            off = 512 if off == 0 else off * 2
            if off + 8 > len(buf):
                raise ValueError("hdf5_lite: not an HDF5 file")
        ver = buf[off + 8]
        if ver not in (0, 1) or buf[off + 13] != 8 or buf[off + 14] != 8:
            raise ValueError(f"hdf5_lite: superblock version {ver} / offset size {buf[off + 13]} not supported")
        self.leaf_k, self.internal_k = struct.unpack("<HH", buf[off + 16:off + 20])
        p = off + 24 + (4 if ver == 1 else 0)
        self.base, _, self.eof, _ = struct.unpack("<QQQQ", buf[p:p + 32])
        self.base = off if self.base == UNDEF else self.base
        self.root_entry = buf[p + 32:p + 72]

    def at(self, addr: int, n: int) -> bytes:
        a = self.base + addr
        if addr == UNDEF or a + n > len(self.buf):
            raise ValueError("hdf5_lite: address outside the file")
        return self.buf[a:a + n]

    def messages(self, addr: int) -> List[Tuple[int, bytes]]:
        ver, _, nmsg, _, size = struct.unpack("<BBHII", self.at(addr, 12))
        if ver != 1:
            raise ValueError(f"hdf5_lite: object header version {ver} not supported")
        blocks, out = [(addr + 16, size)], []
        while blocks and len(out) < nmsg:
            off, length = blocks.pop(0)
            end = off + length
            while off + 8 <= end and len(out) < nmsg:
                mtype, msize, _ = struct.unpack("<HHB", self.at(off, 5))
                data = self.at(off + 8, msize)
                out.append((mtype, data))
                if mtype == 0x0010:                      # continuation block
                    blocks.append(struct.unpack("<QQ", data[:16]))
                off += 8 + msize
        return out

    def group_entries(self, btree: int, heap: int) -> List[Tuple[str, int]]:
        if self.at(heap, 4) != b"HEAP":
            raise ValueError("hdf5_lite: bad local heap")
        _, _, hdata = struct.unpack("<QQQ", self.at(heap + 8, 24))
        out: List[Tuple[str, int]] = []

        def name(o: int) -> str:
            raw = self.at(hdata + o, 256 if hdata + o + 256 <= self.eof else self.eof - hdata - o)
            return raw.split(b"\0", 1)[0].decode()

        def walk(node: int):
            sig = self.at(node, 4)
            if sig == b"TREE":
                ntype, level, used = struct.unpack("<BBH", self.at(node + 4, 4))
                if ntype != 0:
                    raise ValueError("hdf5_lite: not a group B-tree")
                body = self.at(node + 24, (2 * used + 1) * 8)
                for i in range(used):
                    walk(struct.unpack("<Q", body[(2 * i + 1) * 8:(2 * i + 2) * 8])[0])
            elif sig == b"SNOD":
                n = struct.unpack("<H", self.at(node + 6, 2))[0]
                for i in range(n):
                    e = self.at(node + 8 + 40 * i, 40)
                    no, oh = struct.unpack("<QQ", e[:16])
                    out.append((name(no), oh))
            else:
                raise ValueError(f"hdf5_lite: unexpected node {sig!r}")

        walk(btree)
        return out


def read_hdf5(path: str, skipped: Dict[str, str] = None) -> Dict[str, np.ndarray]:
    """Read every contiguous / compact numeric dataset of the root group.  Objects this reader does not parse (chunked or
    compressed datasets, non-numeric types, sub-groups) are left out and, when ``skipped`` is a dict, recorded there as
    ``{name: reason}`` so that callers can say WHY a dataset they need is absent."""
    f = _File(open(path, "rb").read())
    _, root_hdr, cache, _ = struct.unpack("<QQII", f.root_entry[:24])
    btree = heap = None
    if cache == 1:
        btree, heap = struct.unpack("<QQ", f.root_entry[24:40])
    for mtype, data in f.messages(root_hdr):
        if mtype == 0x0011:
            btree, heap = struct.unpack("<QQ", data[:16])
    if btree is None:
        raise ValueError("hdf5_lite: root group has no symbol table")
    out: Dict[str, np.ndarray] = {}
    for name, hdr in f.group_entries(btree, heap):
        shape = dtype = layout = None
        try:
            for mtype, data in f.messages(hdr):
                if mtype == 0x0001:
                    ver, rank = data[0], data[1]
                    p = 8 if ver == 1 else 4
                    shape = tuple(struct.unpack("<Q", data[p + 8 * i:p + 8 * i + 8])[0] for i in range(rank))
                elif mtype == 0x0003:
                    dtype = _parse_datatype(data)
                elif mtype == 0x0008:
                    layout = data
        except (TypeError, ValueError) as e:
            if skipped is not None:
                skipped[name] = f"unsupported object header ({e})"
            continue
        if shape is None or dtype is None or layout is None or layout[0] not in (1, 2, 3):
            if skipped is not None:
                skipped[name] = "not a numeric dataset with a version 1-3 data-layout message (a sub-group?)"
            continue
        count = int(np.prod(shape)) if shape else 1
        if layout[0] in (1, 2):                           # layout versions 1 / 2 (older libraries): class at byte 2
            if layout[2] != 1:
                if skipped is not None:
                    skipped[name] = "chunked / external storage (layout class %d)" % layout[2]
                continue
            addr = struct.unpack("<Q", layout[8:16])[0]
            raw = f.at(addr, count * dtype.itemsize) if count and addr != UNDEF else b""
        elif layout[1] == 1:                              # contiguous
            addr, size = struct.unpack("<QQ", layout[2:18])
            raw = f.at(addr, count * dtype.itemsize) if count and addr != UNDEF else b""
        elif layout[1] == 0:                              # compact
            size = struct.unpack("<H", layout[2:4])[0]
            raw = layout[4:4 + size]
        else:                                             # chunked (what compression / resizable datasets use)
            if skipped is not None:
                skipped[name] = "chunked storage (layout class %d): written with chunks= / compression=" % layout[1]
            continue
        if len(raw) < count * dtype.itemsize:
            out[name] = np.zeros(shape, dtype)            # never written: the fill value (zero)
        else:
            out[name] = np.frombuffer(raw, dtype=dtype, count=count).reshape(shape).copy()
    return out
