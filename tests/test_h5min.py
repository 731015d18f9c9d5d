"""f3: minimal HDF5 reader (revisit-anything_b200/h5min.py) on files laid out like the reference's token / mask stores
(func_vpr.py:647-678).  PARITY UNPINNED against h5py (absent offline): the files come from the module's own writer, plus one
hand-assembled byte layout per storage class below."""
import struct

import numpy as np
import pytest

from revisit_anything_b200 import func_vpr, h5min


def _token_tree(rng, names, D=24, dh=5, dw=7):
    return {n: {"ift_dino": rng.randn(1, D, dh, dw).astype(np.float32)} for n in names}


def test_token_file_chunked_like_the_reference(tmp_path):
    rng = np.random.RandomState(0)
    names = [f"img_{i:03d}.jpg" for i in range(11)]          # > 8 links: several symbol-table nodes
    tree = _token_tree(rng, names)
    path = str(tmp_path / "tokens.h5")
    h5min.write_file(path, tree, chunks_for=lambda name, a: (1, 8, 3, 4) if a.ndim == 4 else None)   # ragged edge chunks
    with h5min.File(path) as f:
        assert sorted(f.keys()) == names and len(f) == 11 and "img_003.jpg" in f and "nope" not in f
        for n in names:
            d = f[n]["ift_dino"]
            assert d.shape == (1, 24, 5, 7) and d.dtype == np.float32
            np.testing.assert_array_equal(d[()], tree[n]["ift_dino"])
            np.testing.assert_array_equal(f[f"{n}/ift_dino"][...], tree[n]["ift_dino"])
        with pytest.raises(KeyError):
            f["img_000.jpg"]["missing"]
    # the drop-in opens a path through the same reader when h5py is absent
    store = func_vpr._open_store(path)
    np.testing.assert_array_equal(np.asarray(store[names[2]]["ift_dino"][()]), tree[names[2]]["ift_dino"])


def test_mask_file_layout_and_preload_masks(tmp_path):
    rng = np.random.RandomState(1)
    masks = [rng.rand(30, 40) < 0.3 for _ in range(13)]
    tree = {"frame_7.png": {"masks": {str(j): {"segmentation": m, "area": np.int64(m.sum()),
                                               "bbox": np.array([1, 2, 3, 4], dtype=np.int64),
                                               "predicted_iou": np.float64(0.5 + 0.01 * j),
                                               "point_coords": np.array([[3.5, 4.5]])} for j, m in enumerate(masks)}}}
    path = str(tmp_path / "masks.h5")
    h5min.write_file(path, tree)
    f = h5min.File(path)
    got = func_vpr.preload_masks(f, "frame_7.png")            # natural key order 0, 1, ..., 12 (func_vpr.py:757-759)
    assert len(got) == 13
    for a, b in zip(got, masks):
        assert a.dtype == np.bool_ and np.array_equal(a, b)
    g = f["frame_7.png/masks/12"]
    assert int(g["area"][()]) == int(masks[12].sum()) and g["bbox"][()].tolist() == [1, 2, 3, 4]
    assert float(g["predicted_iou"][()]) == 0.62 and g["point_coords"].shape == (1, 2)


def test_hand_assembled_contiguous_dataset(tmp_path):
    """Bytes laid out by hand from the format specification (superblock 0, root symbol table with ONE contiguous int32
    dataset [2, 3]) -- independent of the module's writer code paths for the object header and the layout message."""
    O = 8
    buf = bytearray(2048)
    def put(off, data):
        buf[off:off + len(data)] = data
    # superblock
    put(0, h5min.SIG + bytes([0, 0, 0, 0, 0, 8, 8, 0]) + struct.pack("<HHI", 4, 16, 0))
    put(24, struct.pack("<QQQQ", 0, h5min.UNDEF, 2048, h5min.UNDEF))
    put(56, struct.pack("<QQII", 0, 96, 1, 0) + struct.pack("<QQ", 136, 680))        # root entry: ohdr 96, btree 136, heap 680
    # root object header: one symbol-table message
    put(96, struct.pack("<BBHII", 1, 0, 1, 1, 24) + b"\0" * 4 + struct.pack("<HHBBBB", 0x11, 16, 0, 0, 0, 0) + struct.pack("<QQ", 136, 680))
    # group B-tree: one SNOD child at 200
    put(136, b"TREE" + struct.pack("<BBHQQ", 0, 0, 1, h5min.UNDEF, h5min.UNDEF) + struct.pack("<Q", 0) + struct.pack("<QQ", 200, 8))
    # symbol table node: one entry "data" -> object header at 800
    put(200, b"SNOD" + struct.pack("<BBH", 1, 0, 1) + struct.pack("<QQII", 8, 800, 0, 0) + b"\0" * 16)
    # local heap + data segment at 720
    put(680, b"HEAP" + struct.pack("<BBBBQQQ", 0, 0, 0, 0, 24, h5min.UNDEF, 720))
    put(720, b"\0" * 8 + b"data\0\0\0\0")
    # dataset object header: dataspace v1 rank 2, datatype int32 LE signed, layout v3 contiguous at 1024
    space = struct.pack("<BBBBI", 1, 2, 0, 0, 0) + struct.pack("<QQ", 2, 3)
    dtype = struct.pack("<BBBBIHH", 0x10, 0x08, 0, 0, 4, 0, 32)
    layout = struct.pack("<BBQQ", 3, 1, 1024, 24) + b"\0" * 6
    msgs = b""
    for t, d in ((1, space), (3, dtype), (8, layout)):
        d = d + b"\0" * ((-len(d)) % 8)
        msgs += struct.pack("<HHBBBB", t, len(d), 0, 0, 0, 0) + d
    put(800, struct.pack("<BBHII", 1, 0, 3, 1, len(msgs)) + b"\0" * 4 + msgs)
    put(1024, np.arange(6, dtype="<i4").tobytes())
    path = tmp_path / "hand.h5"
    path.write_bytes(bytes(buf))
    f = h5min.File(str(path))
    assert f.keys() == ["data"]
    d = f["data"]
    assert d.shape == (2, 3) and d.dtype == np.dtype("<i4")
    np.testing.assert_array_equal(d[()], np.arange(6, dtype=np.int32).reshape(2, 3))


def test_unsupported_features_are_named(tmp_path):
    p = tmp_path / "v2.h5"
    p.write_bytes(h5min.SIG + bytes([2]) + b"\0" * 200)       # superblock version 2 (libver='latest')
    with pytest.raises(h5min.H5Unsupported, match="superblock version 2"):
        h5min.File(str(p))
    q = tmp_path / "junk.bin"
    q.write_bytes(b"not hdf5" * 1000)
    with pytest.raises(ValueError):
        h5min.File(str(q))
