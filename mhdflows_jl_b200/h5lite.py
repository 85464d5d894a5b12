"""Minimal HDF5 reader / writer for the reference's checkpoint files (`savefile`, `Restart!`, `readMHDFlows`:
src/integrator.jl:208-288, src/utils/IC.jl:245-257) -- no HDF5 library can be installed in this environment.

Scope = what HDF5.jl's `write(f, name, array_or_scalar)` produces with default properties and what `read(f, name)` needs:
  * superblock version 0, 8-byte offsets and lengths;
  * "old style" groups: symbol-table message -> version-1 B-tree (node type 0) + local heap + symbol-table nodes;
  * version-1 object headers (continuation blocks followed when reading);
  * dataspace messages version 1 / 2 (simple and scalar), datatype classes 0 (fixed point) and 1 (floating point),
    little or big endian;
  * data layout message version 3, contiguous or compact storage (chunked / filtered datasets are refused); versions
    1 / 2 (contiguous) are read as well, for files from HDF5 <= 1.6.
The writer emits one root group with up to 2*K_leaf = 8 datasets (the reference writes at most 8: six fields,
`gas_density` / `dye_density`, `time`), all contiguous, names sorted as libhdf5 keeps them.

Structures follow the HDF5 File Format Specification version 2.0 (sections II.A superblock, III.A B-trees, III.B/C symbol
table nodes and entries, III.D local heaps, IV.A object headers, IV.A.2 messages 0x0001, 0x0003, 0x0005, 0x0008, 0x0010,
0x0011).  The reader is pinned against a file written by the real library (tests/test_h5lite.py reads the MATLAB-7.3
fixture that ships inside SciPy); the writer is checked through the reader and byte-level invariants.
Host-side only; nothing here touches the hot path.
"""
from __future__ import annotations

import struct

import numpy as np

SIG = b"\x89HDF\r\n\x1a\n"
UNDEF = 0xFFFFFFFFFFFFFFFF
K_LEAF, K_INTERNAL = 4, 16
HEAP_FREE_NULL = 1          # libhdf5's H5HL_FREE_NULL: "end of free list" marker on disk


class H5Error(RuntimeError):
    pass


def _pad8(n: int) -> int:
    return (n + 7) & ~7


# ------------------------------------------------------------------------------------------------------------------
# reader
# ------------------------------------------------------------------------------------------------------------------
class File:
    """Read-only view of an HDF5 file: `names()`, `read(name)` (NumPy array in C order = HDF5 dimension order, or a NumPy
    scalar), nested groups addressed as "a/b"."""

    def __init__(self, path):
        import mmap
        with open(path, "rb") as f:       # mapped, not read: a restart of a large grid touches only the datasets it reads
            try:
                self.buf = mmap.mmap(f.fileno(), 0, access=mmap.ACCESS_READ)
            except (ValueError, OSError):
                self.buf = f.read()
        start = 0
        while True:                                   # the superblock may sit behind a user block of 512, 1024, ... bytes
            if self.buf[start:start + 8] == SIG:
                break
            start = 512 if start == 0 else start * 2
            if start >= len(self.buf):
                raise H5Error("not an HDF5 file (no superblock signature)")
        b = self.buf
        ver = b[start + 8]
        if ver not in (0, 1):
            raise H5Error(f"superblock version {ver} is not supported (0 and 1 are)")
        self.so, self.sl = b[start + 13], b[start + 14]
        if (self.so, self.sl) != (8, 8):
            raise H5Error("only 8-byte offsets and lengths are supported")
        self.k_leaf, self.k_int = struct.unpack_from("<HH", b, start + 16)
        p = start + 24 + (4 if ver == 1 else 0)
        self.base, _free, self.eof, _drv = struct.unpack_from("<QQQQ", b, p)
        self.root = self._symbol_entry(p + 32)

    # symbol table entry: name offset, header address, cache type, scratch (B-tree, heap)
    def _symbol_entry(self, p):
        name_off, hdr, cache = struct.unpack_from("<QQI", self.buf, p)
        btree, heap = struct.unpack_from("<QQ", self.buf, p + 24)
        return dict(name_off=name_off, header=hdr, cache=cache, btree=btree, heap=heap)

    def _messages(self, addr):
        """(type, data-bytes) of every message of a version-1 object header, continuation blocks included."""
        b = self.buf
        a = self.base + addr
        if b[a] != 1:
            raise H5Error(f"object header version {b[a]} is not supported (version 1 is)")
        nmsg, _refs, size = struct.unpack_from("<HII", b, a + 2)
        blocks = [(a + 16, size)]
        out = []
        while blocks and len(out) < nmsg:
            p, left = blocks.pop(0)
            end = p + left
            while p + 8 <= end and len(out) < nmsg:
                mtype, msize, _flags = struct.unpack_from("<HHB", b, p)
                data = b[p + 8:p + 8 + msize]
                if mtype == 0x0010:                                   # continuation: offset, length
                    off, ln = struct.unpack_from("<QQ", data, 0)
                    blocks.append((self.base + off, ln))
                out.append((mtype, data))
                p += 8 + msize
        return out

    def _group_links(self, btree, heap):
        b = self.buf
        h = self.base + heap
        if b[h:h + 4] != b"HEAP":
            raise H5Error("bad local heap signature")
        seg_addr = struct.unpack_from("<Q", b, h + 24)[0] + self.base

        def name_at(off):
            e = b.find(b"\0", seg_addr + off)
            return b[seg_addr + off:e].decode()

        links = {}

        def walk(addr):
            p = self.base + addr
            if b[p:p + 4] == b"TREE":
                ntype, level, used = struct.unpack_from("<BBH", b, p + 4)
                if ntype != 0:
                    raise H5Error("B-tree node is not a group node")
                q = p + 24
                for i in range(used):
                    child = struct.unpack_from("<Q", b, q + 8 + 16 * i)[0]
                    walk(child)
            elif b[p:p + 4] == b"SNOD":
                n = struct.unpack_from("<H", b, p + 6)[0]
                for i in range(n):
                    e = self._symbol_entry(p + 8 + 40 * i)
                    links[name_at(e["name_off"])] = e
            else:
                raise H5Error("expected a B-tree or symbol-table node")

        walk(btree)
        return links

    def _group_of(self, entry):
        for mtype, data in self._messages(entry["header"]):
            if mtype == 0x0011:
                bt, hp = struct.unpack_from("<QQ", data, 0)
                return self._group_links(bt, hp)
        return None

    def _resolve(self, name):
        entry = self.root
        for part in [s for s in name.split("/") if s]:
            links = self._group_of(entry)
            if links is None or part not in links:
                raise KeyError(name)
            entry = links[part]
        return entry

    def names(self, group=""):
        links = self._group_of(self._resolve(group))
        if links is None:
            raise H5Error(f"{group!r} is not a group")
        return sorted(links)

    def is_group(self, name):
        return self._group_of(self._resolve(name)) is not None

    def read(self, name):
        msgs = self._messages(self._resolve(name)["header"])
        shape = dtype = layout = None
        for mtype, d in msgs:
            if mtype == 0x0001:
                ver, rank, flags = d[0], d[1], d[2]
                if ver == 1:
                    off = 8
                elif ver == 2:
                    off = 4
                    if d[3] == 2:
                        raise H5Error("null dataspace")
                else:
                    raise H5Error(f"dataspace message version {ver}")
                shape = struct.unpack_from(f"<{rank}Q", d, off) if rank else ()
            elif mtype == 0x0003:
                cls, ver = d[0] & 0x0F, d[0] >> 4
                size = struct.unpack_from("<I", d, 4)[0]
                order = ">" if d[1] & 1 else "<"
                if cls == 1:
                    dtype = np.dtype(f"{order}f{size}")
                elif cls == 0:
                    dtype = np.dtype(f"{order}{'i' if d[1] & 8 else 'u'}{size}")
                else:
                    raise H5Error(f"datatype class {cls} is not supported (fixed and floating point are)")
            elif mtype == 0x0008:
                ver = d[0]
                if ver in (1, 2):                                      # files written by HDF5 <= 1.6
                    rank, cls = d[1], d[2]
                    if cls != 1:
                        raise H5Error("only contiguous storage is supported for version-1/2 layout messages")
                    addr = struct.unpack_from("<Q", d, 8)[0]
                    dims = struct.unpack_from(f"<{rank}I", d, 16)      # dataset dimensions, then the element size
                    layout = ("contiguous", addr, int(np.prod(dims, dtype=np.int64)))
                    continue
                if ver != 3:
                    raise H5Error(f"data layout message version {ver} is not supported (1-3 are)")
                if d[1] == 1:
                    layout = ("contiguous",) + struct.unpack_from("<QQ", d, 2)
                elif d[1] == 0:
                    n = struct.unpack_from("<H", d, 2)[0]
                    layout = ("compact", bytes(d[4:4 + n]))
                else:
                    raise H5Error("chunked datasets are not supported")
        if shape is None or dtype is None or layout is None:
            raise H5Error(f"{name!r} is not a dataset this reader understands")
        count = int(np.prod(shape, dtype=np.int64)) if shape else 1
        if layout[0] == "contiguous":
            addr, nbytes = layout[1], layout[2]
            if addr == UNDEF:
                arr = np.zeros(count, dtype)
            else:
                if nbytes < count * dtype.itemsize or self.base + addr + count * dtype.itemsize > len(self.buf):
                    raise H5Error("dataset storage is truncated")
                arr = np.frombuffer(self.buf, dtype, count, self.base + addr)
        else:
            arr = np.frombuffer(layout[1], dtype, count)
        arr = arr.astype(dtype.newbyteorder("="))
        return arr.reshape(shape) if shape else arr[0]


# ------------------------------------------------------------------------------------------------------------------
# writer
# ------------------------------------------------------------------------------------------------------------------
def _msg(mtype, data, flags=0):
    data = data + b"\0" * (_pad8(len(data)) - len(data))
    return struct.pack("<HHB3x", mtype, len(data), flags) + data


def _object_header(msgs):
    body = b"".join(msgs)
    return struct.pack("<BxHII4x", 1, len(msgs), 1, len(body)) + body


def _datatype_msg(dt: np.dtype):
    if dt.kind == "f" and dt.itemsize in (4, 8):
        prec = dt.itemsize * 8
        esize, msize, bias = (8, 23, 127) if dt.itemsize == 4 else (11, 52, 1023)
        head = struct.pack("<BBBBI", 0x11, 0x20, prec - 1, 0, dt.itemsize)       # version 1, class 1; implied mantissa msb; sign bit
        props = struct.pack("<HHBBBBI", 0, prec, msize, esize, 0, msize, bias)
        return head + props
    if dt.kind in "iu" and dt.itemsize in (1, 2, 4, 8):
        head = struct.pack("<BBBBI", 0x10, 0x08 if dt.kind == "i" else 0x00, 0, 0, dt.itemsize)
        return head + struct.pack("<HH", 0, dt.itemsize * 8)
    raise H5Error(f"dtype {dt} cannot be written (float32/64 and integers can)")


def write(path, datasets: dict):
    """Write `datasets` (name -> NumPy array or scalar) as contiguous datasets of the root group."""
    if not 0 < len(datasets) <= 2 * K_LEAF:
        raise H5Error(f"between 1 and {2 * K_LEAF} datasets per file")
    names = sorted(datasets, key=lambda s: s.encode())
    arrays = {}
    for n in names:
        if not n or "/" in n or "\0" in n:
            raise H5Error(f"bad dataset name {n!r}")
        a = np.asarray(datasets[n])
        shape = a.shape                                # ascontiguousarray would turn a 0-d array (scalar dataspace) into 1-d
        a = np.ascontiguousarray(a, dtype=a.dtype.newbyteorder("<") if a.dtype.byteorder == ">" else a.dtype).reshape(shape)
        arrays[n] = a
    # local heap data segment: "" at offset 0, then the names, then one free block up to the end
    seg = bytearray(b"\0" * 8)
    name_off = {}
    for n in names:
        name_off[n] = len(seg)
        raw = n.encode() + b"\0"
        seg += raw + b"\0" * (_pad8(len(raw)) - len(raw))
    free_off = len(seg)
    seg_size = _pad8(max(len(seg) + 16, 88))
    seg += struct.pack("<QQ", HEAP_FREE_NULL, seg_size - free_off) + b"\0" * (seg_size - free_off - 16)
    # file map
    a_root = 96                                        # superblock v0 = 56 + 40 bytes
    root_hdr = _object_header([_msg(0x0011, struct.pack("<QQ", 0, 0))])
    a_btree = a_root + len(root_hdr)
    n_btree = 24 + (2 * K_INTERNAL + 1) * 8 + 2 * K_INTERNAL * 8
    a_heap = a_btree + n_btree
    a_seg = a_heap + 32
    a_snod = a_seg + seg_size
    n_snod = 8 + 2 * K_LEAF * 40
    pos = a_snod + n_snod
    hdr_addr, hdrs = {}, {}
    for n in names:                                    # object headers first (their size does not depend on the addresses)
        a = arrays[n]
        space = struct.pack("<BBB5x", 1, a.ndim, 0) + b"".join(struct.pack("<Q", d) for d in a.shape)
        msgs = [_msg(0x0001, space, 1), _msg(0x0003, _datatype_msg(a.dtype), 1), _msg(0x0005, struct.pack("<BBBBI", 1, 2, 2, 1, 0), 1),
                _msg(0x0008, struct.pack("<BBQQ", 3, 1, 0, 0))]
        hdrs[n] = msgs
        hdr_addr[n] = pos
        pos += len(_object_header(msgs))
    data_addr = {}
    for n in names:
        pos = _pad8(pos)
        data_addr[n] = pos
        pos += arrays[n].nbytes
    eof = pos
    out = bytearray(eof)
    # superblock
    out[0:8] = SIG
    struct.pack_into("<BBBBBBBBHHI", out, 8, 0, 0, 0, 0, 0, 8, 8, 0, K_LEAF, K_INTERNAL, 0)
    struct.pack_into("<QQQQ", out, 24, 0, UNDEF, eof, UNDEF)
    struct.pack_into("<QQII", out, 56, 0, a_root, 1, 0)
    struct.pack_into("<QQ", out, 80, a_btree, a_heap)
    root_hdr = _object_header([_msg(0x0011, struct.pack("<QQ", a_btree, a_heap))])
    out[a_root:a_root + len(root_hdr)] = root_hdr
    # B-tree: one leaf-level node with one child
    out[a_btree:a_btree + 4] = b"TREE"
    struct.pack_into("<BBHQQ", out, a_btree + 4, 0, 0, 1, UNDEF, UNDEF)
    struct.pack_into("<QQQ", out, a_btree + 24, 0, a_snod, name_off[names[-1]])
    # local heap
    out[a_heap:a_heap + 4] = b"HEAP"
    struct.pack_into("<B3xQQQ", out, a_heap + 4, 0, seg_size, free_off, a_seg)
    out[a_seg:a_seg + seg_size] = seg
    # symbol table node
    out[a_snod:a_snod + 4] = b"SNOD"
    struct.pack_into("<BxH", out, a_snod + 4, 1, len(names))
    for i, n in enumerate(names):
        struct.pack_into("<QQII16x", out, a_snod + 8 + 40 * i, name_off[n], hdr_addr[n], 0, 0)
    # datasets
    for n in names:
        a = arrays[n]
        msgs = hdrs[n][:3] + [_msg(0x0008, struct.pack("<BBQQ", 3, 1, data_addr[n], a.nbytes))]
        h = _object_header(msgs)
        out[hdr_addr[n]:hdr_addr[n] + len(h)] = h
        out[data_addr[n]:data_addr[n] + a.nbytes] = a.tobytes()
    with open(path, "wb") as f:
        f.write(out)
    return path
