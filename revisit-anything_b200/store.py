"""On-disk formats either side of the hot path.

The reference keeps three kinds of files:
  * token HDF5   `<img>/ift_dino`  [1, D_t, dh, dw] fp32                       (func_vpr.py:661-662, 674-678)
  * mask HDF5    `<img>/masks/<j>/segmentation` bool [Hm, Wm] (+ SAM metadata)   (func_vpr.py:746-760)
  * result pickles `segFtVLAD1/2` tensors and `{'sims','matches'}` [Nq,200]       (place_rec_main.py:62-75, 292-305)
h5py is not part of this image, so the drop-in functions accept ANY mapping with the same indexing
(`store[img]['ift_dino'][()]`, `store[f'{img}/masks/'].keys()`); `DirStore` below is such a mapping over a directory
of .npy files, and `convert_h5` turns the reference's HDF5 files into it on a machine that has h5py.

`BankFile` is the shardable descriptor bank (descriptors + seg->image map + per-image segment counts) the
multi-GPU search loads row ranges from: one flat file, 4 KiB-aligned fp32 rows, read through np.memmap so a rank
touches only its shard (distributed.shard_bounds).
"""
from __future__ import annotations

import json
import os
import pickle
import re
import struct
from typing import Dict, Iterable, List, Optional, Sequence, Tuple

import numpy as np

_MAGIC = b"SEGVBANK"
_ALIGN = 4096


# ------------------------------------------------------------------------------------------------------------------
# shardable bank file
# ------------------------------------------------------------------------------------------------------------------
def save_bank(path: str, descriptors, im_inds, seg_counts: Optional[Sequence[int]] = None, meta: Optional[dict] = None,
              chunk_rows: int = 65536) -> None:
    """descriptors [n, D] (any float dtype / torch or numpy; stored fp32 like faiss ingests them,
    place_rec_main.py:53-60), im_inds [n] segment -> image id (imInds1 of place_rec_main.py:281-283)."""
    n, D = int(descriptors.shape[0]), int(descriptors.shape[1])
    im = np.ascontiguousarray(np.asarray(im_inds), dtype=np.int32)
    if im.shape != (n,):
        raise ValueError("im_inds must have one entry per descriptor row")
    counts = np.ascontiguousarray(np.bincount(im) if seg_counts is None else np.asarray(seg_counts), dtype=np.int32)
    head = json.dumps({"n": n, "D": D, "dtype": "float32", "n_images": int(counts.size), "meta": meta or {}}).encode()
    off_im = _align(len(_MAGIC) + 8 + len(head))
    off_cnt = _align(off_im + im.nbytes)
    off_rows = _align(off_cnt + counts.nbytes)
    tmp = path + ".tmp"
    with open(tmp, "wb") as fh:
        fh.write(_MAGIC + struct.pack("<II", 1, len(head)) + head)
        fh.seek(off_im); fh.write(im.tobytes())
        fh.seek(off_cnt); fh.write(counts.tobytes())
        fh.seek(off_rows)
        for r0 in range(0, n, chunk_rows):
            blk = descriptors[r0:r0 + chunk_rows]
            blk = blk.detach().cpu().numpy() if hasattr(blk, "detach") else np.asarray(blk)
            fh.write(np.ascontiguousarray(blk, dtype=np.float32).tobytes())
        if n == 0:
            fh.truncate(off_rows)
    os.replace(tmp, path)


def _align(x: int) -> int:
    return (x + _ALIGN - 1) // _ALIGN * _ALIGN


class BankFile:
    """Read side of `save_bank`.  `rows(lo, hi)` is a zero-copy memmap view; `shard(rank, world)` returns
    (row_offset, rows view, im_inds) for distributed.sharded_search_and_vote."""

    def __init__(self, path: str):
        self.path = path
        with open(path, "rb") as fh:
            if fh.read(8) != _MAGIC:
                raise ValueError(f"{path}: not a segvlad bank file")
            ver, hl = struct.unpack("<II", fh.read(8))
            if ver != 1:
                raise ValueError(f"{path}: unsupported bank version {ver}")
            h = json.loads(fh.read(hl))
        self.n, self.D, self.n_images, self.meta = h["n"], h["D"], h["n_images"], h["meta"]
        off_im = _align(16 + hl)
        off_cnt = _align(off_im + 4 * self.n)
        self._off_rows = _align(off_cnt + 4 * self.n_images)
        need = self._off_rows + 4 * self.n * self.D
        if os.path.getsize(path) < need:
            raise ValueError(f"{path}: truncated ({os.path.getsize(path)} < {need} bytes)")
        self.im_inds = np.fromfile(path, dtype=np.int32, count=self.n, offset=off_im)
        self.seg_counts = np.fromfile(path, dtype=np.int32, count=self.n_images, offset=off_cnt)

    def rows(self, lo: int = 0, hi: Optional[int] = None) -> np.ndarray:
        hi = self.n if hi is None else hi
        if not (0 <= lo <= hi <= self.n):
            raise IndexError("row range outside the bank")
        if hi == lo:
            return np.zeros((0, self.D), dtype=np.float32)
        return np.memmap(self.path, dtype=np.float32, mode="r", offset=self._off_rows + 4 * lo * self.D,
                         shape=(hi - lo, self.D))

    def shard(self, rank: int, world: int) -> Tuple[int, np.ndarray, np.ndarray]:
        from .distributed import shard_bounds
        lo, hi = shard_bounds(self.n, world)[rank]
        return lo, self.rows(lo, hi), self.im_inds

    def seg_ranges(self) -> List[np.ndarray]:
        """segRange lists of place_rec_main.py:354-355 (contiguous rows per image)."""
        off = np.concatenate([[0], np.cumsum(self.seg_counts)])
        return [np.arange(off[i], off[i + 1]) for i in range(self.n_images)]


# ------------------------------------------------------------------------------------------------------------------
# h5py-shaped directory store (tokens and masks)
# ------------------------------------------------------------------------------------------------------------------
def _natural(s: str):
    return [int(t) if t.isdigit() else t for t in re.split(r"(\d+)", s)]


class _Leaf:
    """`dataset[()]` of h5py: returns the whole array."""

    def __init__(self, path):
        self._path = path

    def __getitem__(self, key):
        a = np.load(self._path, mmap_mode="r")
        return np.asarray(a) if key == () else np.asarray(a[key])

    @property
    def shape(self):
        return np.load(self._path, mmap_mode="r").shape


class DirStore:
    """Directory tree <-> nested groups; `name.npy` files are datasets.  Supports the access patterns of the
    reference: `f.keys()`, `f[img]['ift_dino'][()]`, `f[f'{img}/masks/'].keys()`, `f[f'{img}/masks/{j}']['segmentation'][()]`.
    Group names are stored percent-escaped so image keys containing '/' or '.' survive."""

    def __init__(self, root: str):
        self._root = root
        if not os.path.isdir(root):
            raise FileNotFoundError(root)

    @staticmethod
    def _enc(name: str) -> str:
        return name.replace("%", "%25").replace("/", "%2F")

    @staticmethod
    def _dec(name: str) -> str:
        return name.replace("%2F", "/").replace("%25", "%")

    def keys(self):
        names = []
        for e in os.listdir(self._root):
            names.append(self._dec(e[:-4]) if e.endswith(".npy") and os.path.isfile(os.path.join(self._root, e))
                         else self._dec(e))
        return sorted(names, key=_natural)

    def __contains__(self, key):
        try:
            self[key]
            return True
        except KeyError:
            return False

    def __len__(self):
        return len(os.listdir(self._root))

    def __iter__(self):
        return iter(self.keys())

    def __getitem__(self, key: str):
        node = self._root
        parts = [p for p in key.split("/") if p]
        i = 0
        while i < len(parts):
            # image keys may themselves contain '/': take the longest stored name that matches
            for j in range(len(parts), i, -1):
                cand = os.path.join(node, self._enc("/".join(parts[i:j])))
                if os.path.isdir(cand):
                    node, i = cand, j
                    break
                if os.path.isfile(cand + ".npy") and j == len(parts):
                    return _Leaf(cand + ".npy")
            else:
                raise KeyError(key)
        return DirStore(node)

    # -- writing ----------------------------------------------------------------------------------------------------
    def put(self, group: str, name: str, array) -> None:
        parts = [p for p in group.split("/") if p]
        d = self._root
        for p in parts:
            d = os.path.join(d, self._enc(p))
        os.makedirs(d, exist_ok=True)
        np.save(os.path.join(d, self._enc(name) + ".npy"), np.asarray(array))

    @classmethod
    def create(cls, root: str) -> "DirStore":
        os.makedirs(root, exist_ok=True)
        return cls(root)


def write_tokens(store: DirStore, img_key: str, tokens_1dhw) -> None:
    """One image's DINOv2 tokens [1, D_t, dh, dw] fp32 under `<img>/ift_dino` (func_vpr.py:661-662)."""
    a = np.asarray(tokens_1dhw, dtype=np.float32)
    if a.ndim != 4 or a.shape[0] != 1:
        raise ValueError("tokens must be [1, D_t, dh, dw]")
    store.put(img_key, "ift_dino", a)


def write_masks(store: DirStore, img_key: str, masks: Iterable[dict]) -> None:
    """SAM records (automatic_mask_generator.py:185-191) under `<img>/masks/<j>/<field>` (func_vpr.py:674-678)."""
    for j, rec in enumerate(masks):
        rec = rec if isinstance(rec, dict) else {"segmentation": rec}
        for field, val in rec.items():
            store.put(f"{img_key}/masks/{j}", field, np.asarray(val))


def convert_h5(h5_path: str, out_dir: str) -> DirStore:
    """HDF5 (reference layout) -> DirStore; needs h5py, i.e. runs where the reference's files were produced."""
    try:
        import h5py
    except ImportError as e:
        raise RuntimeError("convert_h5 needs h5py (not installed in this image)") from e
    st = DirStore.create(out_dir)

    def walk(group, prefix):
        for name, item in group.items():
            if isinstance(item, h5py.Dataset):
                st.put(prefix, name, item[()])
            else:
                walk(item, f"{prefix}/{name}" if prefix else name)

    with h5py.File(h5_path, "r") as f:
        walk(f, "")
    return st


# ------------------------------------------------------------------------------------------------------------------
# result pickles (place_rec_main.py:62-75, 292-305, 357-370)
# ------------------------------------------------------------------------------------------------------------------
def save_segment_features(path: str, seg_ft) -> None:
    os.makedirs(os.path.dirname(os.path.abspath(path)), exist_ok=True)
    with open(path, "wb") as fh:
        pickle.dump(seg_ft, fh)


def load_segment_features(path: str):
    with open(path, "rb") as fh:
        return pickle.load(fh)


def save_search_results(path: str, sims, matches) -> None:
    """`{'sims': D [Nq,200], 'matches': I [Nq,200]}` exactly as place_rec_main.py:68-72 pickles them."""
    save_segment_features(path, {"sims": np.asarray(sims), "matches": np.asarray(matches)})
