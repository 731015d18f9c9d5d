"""Minimal pure-Python HDF5 reader (+ writer of the same subset) for the reference's token / mask files.

The reference stores its inputs with h5py (func_vpr.py:647-678):
    tokens  f[img]['ift_dino']                     float32 [1, D_t, dh, dw]      create_dataset(..., chunks=True)
    masks   f[img]['masks'][j]['segmentation']     bool [Hm, Wm] (+ SAM metadata)  create_dataset(name, data=...)
and reads them back through `h5py.File` (func_vpr.py:1079-1080, 746-760).  h5py / libhdf5 are not part of this image, so
this module implements the part of the HDF5 file format those files use, from the published format specification
("HDF5 File Format Specification Version 2.0", the layout h5py writes with its default libver='earliest'):

    superblock version 0 / 1; version-1 object headers (+ continuation blocks); old-style groups (symbol-table message,
    version-1 B-tree of group nodes, local heap, SNOD symbol-table nodes); dataspace message v1 / v2; datatypes
    fixed-point, IEEE float, and enum over a fixed-point base (h5py's bool); data layout message v3 with COMPACT,
    CONTIGUOUS and CHUNKED storage (version-1 B-tree of raw-data chunks, no filters); fill value = 0 for absent chunks.

Anything else (new-style link messages, filter pipelines such as gzip, variable-length types, superblock 2/3) raises
`H5Unsupported` naming the feature -- convert such files with `store.convert_h5` on a machine that has h5py.

PARITY UNPINNED: no HDF5 library and no HDF5 file exist in this offline image, so the reader is checked against files
produced by the writer below (same specification, same author) and against hand-assembled byte layouts in the tests, not
against h5py's own output.

`File(path)` behaves like the read side of `h5py.File`: `f[name]` (also 'a/b/c' paths), `.keys()`, `in`, iteration;
datasets support `[()]`, `[...]`, `.shape`, `.dtype`.  It plugs into every drop-in function that takes a store
(`func_vpr.seg_vlad_gpu_single`, `preload_masks`, `aggFt`).
"""
from __future__ import annotations

import struct
from typing import Dict, Iterator, List, Optional, Tuple

import numpy as np

SIG = b"\x89HDF\r\n\x1a\n"
UNDEF = 0xFFFFFFFFFFFFFFFF


class H5Unsupported(NotImplementedError):
    pass


# ======================================================================================================================
# reader
# ======================================================================================================================
class _Buf:
    def __init__(self, data):
        self.d = data

    def u(self, off: int, n: int) -> int:
        return int.from_bytes(self.d[off:off + n], "little")

    def raw(self, off: int, n: int) -> bytes:
        return bytes(self.d[off:off + n])


class File:
    def __init__(self, path: str, mode: str = "r"):
        if mode != "r":
            raise H5Unsupported("h5min.File is read-only (use h5min.write_file to create files)")
        self._mm = np.memmap(path, dtype=np.uint8, mode="r")
        self.b = _Buf(self._mm)
        base = None
        for off in (0, 512, 1024, 2048, 4096):
            if self.b.raw(off, 8) == SIG:
                base = off
                break
        if base is None:
            raise ValueError(f"{path}: not an HDF5 file (no superblock signature)")
        ver = self.b.u(base + 8, 1)
        if ver not in (0, 1):
            raise H5Unsupported(f"superblock version {ver} (files written with libver='latest'); only 0 / 1 are read")
        self.O = self.b.u(base + 13, 1)
        self.L = self.b.u(base + 14, 1)
        p = base + 24 + (4 if ver == 1 else 0)
        self.base_addr = self.b.u(p, self.O)
        p += 4 * self.O                                   # base, free-space, end-of-file, driver-info addresses
        self.root = Group(self, *self._symtab_entry(p)[1:])

    # symbol table entry -> (name offset, object header address, (btree, heap) or None)
    def _symtab_entry(self, p: int):
        O = self.O
        name_off = self.b.u(p, O)
        ohdr = self.b.u(p + O, O)
        cache = self.b.u(p + 2 * O, 4)
        scratch = p + 2 * O + 8
        bt_heap = (self.b.u(scratch, O), self.b.u(scratch + O, O)) if cache == 1 else None
        return name_off, ohdr, bt_heap

    def _messages(self, addr: int) -> List[Tuple[int, int, int]]:
        """Version-1 object header at `addr` -> [(type, data offset, size)] incl. continuation blocks."""
        b = self.b
        a = addr + self.base_addr
        if b.raw(a, 4) == b"OHDR":
            raise H5Unsupported("version-2 object headers (libver='latest')")
        if b.u(a, 1) != 1:
            raise H5Unsupported(f"object header version {b.u(a, 1)}")
        n_msgs = b.u(a + 2, 2)
        size = b.u(a + 8, 4)
        blocks = [(a + 16, size)]
        out = []
        while blocks and len(out) < n_msgs:
            p, left = blocks.pop(0)
            end = p + left
            while p + 8 <= end and len(out) < n_msgs:
                mtype, msize = b.u(p, 2), b.u(p + 2, 2)
                data = p + 8
                if mtype == 0x10:                          # continuation
                    blocks.append((b.u(data, self.O) + self.base_addr, b.u(data + self.O, self.L)))
                out.append((mtype, data, msize))
                p = data + ((msize + 7) & ~7)
        return out

    # mapping interface (delegated to the root group)
    def __getitem__(self, name):
        return self.root[name]

    def __contains__(self, name):
        return name in self.root

    def keys(self):
        return self.root.keys()

    def __iter__(self):
        return iter(self.root)

    def __len__(self):
        return len(self.root)

    def close(self):
        pass

    def __enter__(self):
        return self

    def __exit__(self, *a):
        return False


class Group:
    def __init__(self, f: File, ohdr: int, bt_heap=None):
        self.f = f
        self.ohdr = ohdr
        self._bt_heap = bt_heap
        self._links: Optional[Dict[str, int]] = None

    def _load(self) -> Dict[str, int]:
        if self._links is not None:
            return self._links
        f, b = self.f, self.f.b
        bt_heap = self._bt_heap
        if bt_heap is None:
            for mtype, data, _ in f._messages(self.ohdr):
                if mtype == 0x11:
                    bt_heap = (b.u(data, f.O), b.u(data + f.O, f.O))
                elif mtype in (0x02, 0x06):
                    raise H5Unsupported("new-style groups (link messages, libver='latest')")
        if bt_heap is None:
            raise ValueError("object is not a group")
        btree, heap = bt_heap[0] + f.base_addr, bt_heap[1] + f.base_addr
        if b.raw(heap, 4) != b"HEAP":
            raise ValueError("bad local heap signature")
        heap_data = b.u(heap + 8 + 2 * f.L, f.O) + f.base_addr
        links: Dict[str, int] = {}

        def name_at(off: int) -> str:
            p = heap_data + off
            e = p
            while b.u(e, 1) != 0:
                e += 1
            return b.raw(p, e - p).decode("utf-8")

        def walk(node: int):
            if b.raw(node, 4) == b"SNOD":
                n = b.u(node + 6, 2)
                p = node + 8
                for _ in range(n):
                    name_off, ohdr, _ = f._symtab_entry(p)
                    links[name_at(name_off)] = ohdr
                    p += 2 * f.O + 24
                return
            if b.raw(node, 4) != b"TREE" or b.u(node + 4, 1) != 0:
                raise ValueError("bad group B-tree node")
            n = b.u(node + 6, 2)
            p = node + 8 + 2 * f.O + f.L                   # skip key 0
            for _ in range(n):
                walk(b.u(p, f.O) + f.base_addr)
                p += f.O + f.L
        walk(btree)
        self._links = links
        return links

    def keys(self):
        return list(self._load().keys())

    def __iter__(self) -> Iterator[str]:
        return iter(self.keys())

    def __len__(self):
        return len(self._load())

    def __contains__(self, name) -> bool:
        try:
            self[name]
            return True
        except KeyError:
            return False

    def __getitem__(self, name):
        parts = [p for p in str(name).split("/") if p]
        obj = self
        for part in parts:
            if not isinstance(obj, Group):
                raise KeyError(name)
            links = obj._load()
            if part not in links:
                raise KeyError(name)
            obj = _open(obj.f, links[part])
        return obj


def _open(f: File, ohdr: int):
    types = {m[0] for m in f._messages(ohdr)}
    if 0x08 in types:
        return Dataset(f, ohdr)
    if 0x11 in types:
        return Group(f, ohdr)
    if types & {0x02, 0x06}:
        raise H5Unsupported("new-style groups (link messages, libver='latest')")
    raise H5Unsupported("object that is neither an old-style group nor a dataset")


def _decode_dtype(b: _Buf, p: int) -> np.dtype:
    cv = b.u(p, 1)
    cls, bits0, size = cv & 0x0F, b.u(p + 1, 1), b.u(p + 4, 4)
    endian = ">" if bits0 & 1 else "<"
    if cls == 0:
        return np.dtype(f"{endian}{'i' if bits0 & 8 else 'u'}{size}")
    if cls == 1:
        if size not in (2, 4, 8):
            raise H5Unsupported(f"{size}-byte floating point")
        return np.dtype(f"{endian}f{size}")
    if cls == 8:                                           # enum: base type at +8; h5py stores numpy bool as enum(int8)
        base = _decode_dtype(b, p + 8)
        return np.dtype(bool) if base.itemsize == 1 else base
    raise H5Unsupported(f"datatype class {cls} (only fixed-point, float and enum are read)")


class Dataset:
    def __init__(self, f: File, ohdr: int):
        self.f = f
        b = f.b
        self.shape: Tuple[int, ...] = ()
        self.dtype = None
        self._layout = None
        for mtype, data, msize in f._messages(ohdr):
            if mtype == 0x01:
                ver, rank = b.u(data, 1), b.u(data + 1, 1)
                dims = data + (8 if ver == 1 else 4)
                if ver not in (1, 2):
                    raise H5Unsupported(f"dataspace message version {ver}")
                self.shape = tuple(b.u(dims + i * f.L, f.L) for i in range(rank))
            elif mtype == 0x03:
                self.dtype = _decode_dtype(b, data)
            elif mtype == 0x0B:
                if b.u(data + 1, 1) > 0:
                    raise H5Unsupported("filter pipeline (compressed / shuffled chunks)")
            elif mtype == 0x08:
                ver = b.u(data, 1)
                if ver != 3:
                    raise H5Unsupported(f"data layout message version {ver}")
                cls = b.u(data + 1, 1)
                if cls == 0:
                    self._layout = ("compact", data + 4, b.u(data + 2, 2))
                elif cls == 1:
                    self._layout = ("contiguous", b.u(data + 2, f.O), b.u(data + 2 + f.O, f.L))
                elif cls == 2:
                    nd = b.u(data + 2, 1)
                    bt = b.u(data + 3, f.O)
                    cdims = tuple(b.u(data + 3 + f.O + 4 * i, 4) for i in range(nd))
                    self._layout = ("chunked", bt, cdims)
                else:
                    raise H5Unsupported(f"data layout class {cls}")
        if self.dtype is None or self._layout is None:
            raise ValueError("dataset without datatype / layout message")

    def _read(self) -> np.ndarray:
        f, b = self.f, self.f.b
        n = int(np.prod(self.shape, dtype=np.int64)) if self.shape else 1
        kind = self._layout[0]
        if kind == "compact":
            return np.frombuffer(b.raw(self._layout[1], n * self.dtype.itemsize), dtype=self.dtype).reshape(self.shape).copy()
        if kind == "contiguous":
            addr = self._layout[1]
            if addr == UNDEF >> (64 - 8 * f.O):
                return np.zeros(self.shape, dtype=self.dtype)
            return np.frombuffer(b.raw(addr + f.base_addr, n * self.dtype.itemsize), dtype=self.dtype).reshape(self.shape).copy()
        _, bt, cdims = self._layout
        rank = len(self.shape)
        chunk = cdims[:rank]
        out = np.zeros(self.shape, dtype=self.dtype)
        if bt == UNDEF >> (64 - 8 * f.O):
            return out
        csize = int(np.prod(chunk, dtype=np.int64)) * self.dtype.itemsize
        key_len = 8 + 8 * (rank + 1)

        def walk(node: int):
            if b.raw(node, 4) != b"TREE" or b.u(node + 4, 1) != 1:
                raise ValueError("bad chunk B-tree node")
            level, n_ent = b.u(node + 5, 1), b.u(node + 6, 2)
            p = node + 8 + 2 * f.O
            for _ in range(n_ent):
                nbytes, fmask = b.u(p, 4), b.u(p + 4, 4)
                offs = tuple(b.u(p + 8 + 8 * i, 8) for i in range(rank))
                child = b.u(p + key_len, f.O) + f.base_addr
                if level > 0:
                    walk(child)
                else:
                    if fmask != 0 or nbytes != csize:
                        raise H5Unsupported("filtered (compressed) chunk")
                    blk = np.frombuffer(b.raw(child, csize), dtype=self.dtype).reshape(chunk)
                    sl_o = tuple(slice(o, min(o + c, s)) for o, c, s in zip(offs, chunk, self.shape))
                    sl_c = tuple(slice(0, s.stop - s.start) for s in sl_o)
                    out[sl_o] = blk[sl_c]
                p += key_len + f.O
        walk(bt + f.base_addr)
        return out

    def __getitem__(self, key):
        arr = self._read()
        if key == () or key is Ellipsis:
            return arr if self.shape else arr.reshape(())[()]
        return arr[key]

    def __array__(self, dtype=None):
        a = self._read()
        return a if dtype is None else a.astype(dtype)


# ======================================================================================================================
# writer (same subset; used by the tests and by store.export_h5 to hand results back to the reference's scripts)
# ======================================================================================================================
class _Out:
    def __init__(self):
        self.buf = bytearray()

    def alloc(self, n: int, align: int = 8) -> int:
        pad = (-len(self.buf)) % align
        self.buf += b"\0" * pad
        off = len(self.buf)
        self.buf += b"\0" * n
        return off

    def put(self, off: int, data: bytes):
        self.buf[off:off + len(data)] = data


def _dtype_msg(dt: np.dtype) -> bytes:
    dt = np.dtype(dt)
    if dt == np.dtype(bool):                               # h5py: enum {FALSE = 0, TRUE = 1} over int8
        base = _dtype_msg(np.dtype("i1"))
        names = b"FALSE\0\0\0" + b"TRUE\0\0\0\0"
        return struct.pack("<BBBBI", 0x18, 2, 0, 0, 1) + base + names + bytes([0, 1])
    if dt.kind in "iu":
        bits0 = 0x08 if dt.kind == "i" else 0x00
        return struct.pack("<BBBBIHH", 0x10, bits0, 0, 0, dt.itemsize, 0, dt.itemsize * 8)
    if dt.kind == "f" and dt.itemsize in (4, 8):
        if dt.itemsize == 4:
            props = struct.pack("<HHBBBBI", 0, 32, 23, 8, 0, 23, 127)
            bits = (0x20, 31, 0)
        else:
            props = struct.pack("<HHBBBBI", 0, 64, 52, 11, 0, 52, 1023)
            bits = (0x20, 63, 0)
        return struct.pack("<BBBBI", 0x11, *bits, dt.itemsize) + props
    raise H5Unsupported(f"writer: dtype {dt}")


def _msg(mtype: int, data: bytes) -> bytes:
    pad = (-len(data)) % 8
    return struct.pack("<HHBBBB", mtype, len(data) + pad, 0, 0, 0, 0) + data + b"\0" * pad


def _ohdr(msgs: List[bytes]) -> bytes:
    body = b"".join(msgs)
    return struct.pack("<BBHII", 1, 0, len(msgs), 1, len(body)) + b"\0" * 4 + body


def _write_dataset(o: _Out, arr: np.ndarray, chunks=None) -> int:
    arr = np.asarray(arr)
    if not arr.flags.c_contiguous:                         # (np.ascontiguousarray would turn a scalar into shape (1,))
        arr = arr.copy(order="C")
    store = arr.view(np.int8) if arr.dtype == np.dtype(bool) else arr
    rank = arr.ndim
    space = struct.pack("<BBBBI", 1, rank, 0, 0, 0) + b"".join(struct.pack("<Q", s) for s in arr.shape)
    if chunks is None and arr.nbytes <= 64:                # compact
        layout = struct.pack("<BBH", 3, 0, arr.nbytes) + store.tobytes()
    elif chunks is None:                                   # contiguous
        addr = o.alloc(arr.nbytes)
        o.put(addr, store.tobytes())
        layout = struct.pack("<BBQQ", 3, 1, addr, arr.nbytes)
    else:                                                  # chunked, one leaf B-tree node
        chunks = tuple(int(c) for c in chunks)
        grid = [range(0, s, c) for s, c in zip(arr.shape, chunks)]
        entries = []
        for offs in np.ndindex(*[len(g) for g in grid]):
            o0 = tuple(g[i] for g, i in zip(grid, offs))
            blk = np.zeros(chunks, dtype=store.dtype)
            sl = tuple(slice(a, min(a + c, s)) for a, c, s in zip(o0, chunks, arr.shape))
            blk[tuple(slice(0, s.stop - s.start) for s in sl)] = store[sl]
            addr = o.alloc(blk.nbytes)
            o.put(addr, blk.tobytes())
            entries.append((o0, addr, blk.nbytes))
        key_len = 8 + 8 * (rank + 1)
        node = o.alloc(8 + 16 + len(entries) * (key_len + 8) + key_len)
        body = b"TREE" + struct.pack("<BBHQQ", 1, 0, len(entries), UNDEF, UNDEF)
        for o0, addr, nb in entries:
            body += struct.pack("<II", nb, 0) + b"".join(struct.pack("<Q", v) for v in o0) + struct.pack("<Q", 0)
            body += struct.pack("<Q", addr)
        body += struct.pack("<II", 0, 0) + b"".join(struct.pack("<Q", s) for s in arr.shape) + struct.pack("<Q", 0)
        o.put(node, body)
        layout = struct.pack("<BBB", 3, 2, rank + 1) + struct.pack("<Q", node) + \
            b"".join(struct.pack("<I", c) for c in chunks) + struct.pack("<I", arr.dtype.itemsize)
    hdr = _ohdr([_msg(0x01, space), _msg(0x03, _dtype_msg(arr.dtype)), _msg(0x08, layout)])
    addr = o.alloc(len(hdr))
    o.put(addr, hdr)
    return addr


def _write_group(o: _Out, tree: dict, chunks_for=None) -> Tuple[int, int, int]:
    """tree: {name: ndarray | dict}.  Returns (object header address, B-tree address, heap address)."""
    children = []
    for name in sorted(tree.keys()):                       # B-tree keys are ordered by name
        v = tree[name]
        if isinstance(v, dict):
            children.append((name, _write_group(o, v, chunks_for)[0]))
        else:
            arr = np.asarray(v)
            ch = chunks_for(name, arr) if chunks_for else None
            children.append((name, _write_dataset(o, arr, ch)))
    # local heap: empty name at offset 0, then the link names (8-byte aligned)
    heap_data = bytearray(b"\0" * 8)
    offs = []
    for name, _ in children:
        offs.append(len(heap_data))
        nb = name.encode("utf-8") + b"\0"
        heap_data += nb + b"\0" * ((-len(nb)) % 8)
    hd_addr = o.alloc(len(heap_data))
    o.put(hd_addr, bytes(heap_data))
    heap = o.alloc(32)
    o.put(heap, b"HEAP" + struct.pack("<BBBBQQQ", 0, 0, 0, 0, len(heap_data), UNDEF, hd_addr))
    # symbol-table nodes of <= 8 entries under one B-tree node
    snods = []
    for i in range(0, max(len(children), 1), 8):
        part = list(zip(children[i:i + 8], offs[i:i + 8]))
        snod = o.alloc(8 + 16 * 40)
        body = b"SNOD" + struct.pack("<BBH", 1, 0, len(part))
        for (name, ohdr), noff in part:
            body += struct.pack("<QQII", noff, ohdr, 0, 0) + b"\0" * 16
        o.put(snod, body)
        snods.append((snod, part[-1][1] if part else 0))
    bt = o.alloc(8 + 16 + len(snods) * 16 + 8)
    body = b"TREE" + struct.pack("<BBHQQ", 0, 0, len(snods), UNDEF, UNDEF) + struct.pack("<Q", 0)
    for snod, last_off in snods:
        body += struct.pack("<QQ", snod, last_off)
    o.put(bt, body)
    hdr = _ohdr([_msg(0x11, struct.pack("<QQ", bt, heap))])
    addr = o.alloc(len(hdr))
    o.put(addr, hdr)
    return addr, bt, heap


def write_file(path: str, tree: dict, chunks_for=None) -> None:
    """Write {name: ndarray | nested dict} as an HDF5 file of the subset documented above (superblock 0, old-style groups).
    chunks_for(name, array) -> chunk shape or None selects chunked storage per dataset (the reference writes its token
    datasets with chunks=True, func_vpr.py:662)."""
    o = _Out()
    sb = o.alloc(96)
    root, bt, heap = _write_group(o, tree, chunks_for)
    eof = len(o.buf)
    head = SIG + struct.pack("<BBBBBBBBHHI", 0, 0, 0, 0, 0, 8, 8, 0, 4, 16, 0)
    head += struct.pack("<QQQQ", 0, UNDEF, eof, UNDEF)
    head += struct.pack("<QQII", 0, root, 1, 0) + struct.pack("<QQ", bt, heap)
    o.put(sb, head)
    with open(path, "wb") as fh:
        fh.write(bytes(o.buf))
