"""f3: on-disk formats -- shardable bank file, h5py-shaped directory store, result pickles (CPU only)."""
import numpy as np
import pytest

from revisit_anything_b200 import func_vpr, store
from revisit_anything_b200.distributed import shard_bounds


def test_bank_roundtrip_and_shards(tmp_path):
    rng = np.random.RandomState(0)
    counts = rng.randint(1, 9, size=37)
    n, D = int(counts.sum()), 24
    desc = rng.randn(n, D)                                  # fp64 in, fp32 on disk (faiss ingest dtype)
    im = np.repeat(np.arange(37), counts)
    p = str(tmp_path / "bank.segv")
    store.save_bank(p, desc, im, meta={"dataset": "17places"}, chunk_rows=50)
    b = store.BankFile(p)
    assert (b.n, b.D, b.n_images, b.meta["dataset"]) == (n, D, 37, "17places")
    np.testing.assert_array_equal(b.rows(), desc.astype(np.float32))
    np.testing.assert_array_equal(b.im_inds, im)
    np.testing.assert_array_equal(b.seg_counts, counts)
    got = []
    for r in range(3):
        lo, rows, imi = b.shard(r, 3)
        assert lo == shard_bounds(n, 3)[r][0] and imi is b.im_inds
        got.append(np.asarray(rows))
    np.testing.assert_array_equal(np.concatenate(got), desc.astype(np.float32))
    rng_list = b.seg_ranges()
    assert len(rng_list) == 37 and rng_list[-1][-1] == n - 1
    with pytest.raises(IndexError):
        b.rows(0, n + 1)
    assert b.rows(5, 5).shape == (0, D)


def test_bank_empty_and_corrupt(tmp_path):
    p = str(tmp_path / "e.segv")
    store.save_bank(p, np.zeros((0, 8)), np.zeros(0, dtype=np.int64))
    b = store.BankFile(p)
    assert b.n == 0 and b.rows().shape == (0, 8) and b.shard(1, 2)[1].shape == (0, 8)
    with open(p, "r+b") as fh:
        fh.write(b"XXXX")
    with pytest.raises(ValueError):
        store.BankFile(p)
    q = str(tmp_path / "t.segv")
    store.save_bank(q, np.ones((10, 8)), np.zeros(10, dtype=np.int64))
    with open(q, "r+b") as fh:
        fh.truncate(4096 * 3 + 16)
    with pytest.raises(ValueError):
        store.BankFile(q)
    with pytest.raises(ValueError):
        store.save_bank(q, np.ones((10, 8)), np.zeros(9))


def test_dirstore_has_the_reference_access_patterns(tmp_path):
    st = store.DirStore.create(str(tmp_path / "toks"))
    rng = np.random.RandomState(1)
    toks = {f"frame{i}.jpg": rng.randn(1, 16, 3, 4).astype(np.float32) for i in (10, 9, 100)}
    for k, v in toks.items():
        store.write_tokens(st, k, v)
    assert list(st.keys()) == ["frame9.jpg", "frame10.jpg", "frame100.jpg"]          # natsorted order
    np.testing.assert_array_equal(st["frame10.jpg"]["ift_dino"][()], toks["frame10.jpg"])
    assert st["frame9.jpg"]["ift_dino"].shape == (1, 16, 3, 4)
    ms = store.DirStore.create(str(tmp_path / "masks"))
    masks = [{"segmentation": rng.rand(6, 8) > 0.5, "area": 7 + j, "bbox": [0, 0, 3, 3]} for j in range(12)]
    store.write_masks(ms, "frame10.jpg", masks)
    got = func_vpr.preload_masks(ms, "frame10.jpg")                                   # func_vpr.py:746-760 access path
    assert len(got) == 12
    for j in range(12):
        np.testing.assert_array_equal(got[j], masks[j]["segmentation"])
    assert int(ms["frame10.jpg/masks/11"]["area"][()]) == 18
    with pytest.raises(KeyError):
        st["nope"]
    with pytest.raises(ValueError):
        store.write_tokens(st, "x", np.zeros((16, 3, 4)))


def test_result_pickles(tmp_path):
    import pickle

    import torch
    p = str(tmp_path / "results" / "global" / "SegLoc" / "17places_segFtVLAD1.pkl")
    t = torch.randn(5, 7, dtype=torch.float64)
    store.save_segment_features(p, t)
    assert torch.equal(store.load_segment_features(p), t)
    q = str(tmp_path / "r.pkl")
    store.save_search_results(q, np.ones((3, 200), np.float32), np.zeros((3, 200), np.int64))
    with open(q, "rb") as fh:
        d = pickle.load(fh)
    assert set(d) == {"sims", "matches"} and d["matches"].dtype == np.int64 and d["sims"].shape == (3, 200)
