"""GPU parity of the AnyLoc whole-image baseline: utilities.VLAD.generate through the SegVLAD
aggregation kernels (one all-ones segment per image), aggFt(..., 'vlad'), get_recall through the search kernel."""
import os

import numpy as np
import pytest
import torch

from oracle import segvlad_oracle as O
from revisit_anything_b200 import func_vpr
from revisit_anything_b200.utilities import VLAD
from gpu_util import assert_desc_close

pytestmark = pytest.mark.gpu


def _vlad(centers):
    v = VLAD(centers.shape[0], desc_dim=None, dist_mode="cosine", vlad_mode="hard", cache_dir=None)
    v.set_centers(torch.as_tensor(centers))
    return v


def test_generate_golden_reference(golden_dir):
    g = np.load(os.path.join(golden_dir, "anyloc_vlad.npz"))
    v = _vlad(g["centers"])
    multi = v.generate_multi(g["tokens"])                       # [B, N, D] batched launch
    assert multi.dtype == torch.float32 and tuple(multi.shape) == g["vlad"].shape
    assert_desc_close(multi.numpy(), g["vlad"])                 # reference sums in fp32, kernel in fp64: 1e-5 rel
    one = v.generate(g["tokens"][2])
    np.testing.assert_array_equal(one.numpy(), multi[2].numpy())


def test_ragged_list_and_real_shape_vs_oracle():
    g = torch.Generator().manual_seed(7)
    K, D = 32, 1536
    centers = 0.5 * torch.nn.functional.normalize(torch.randn(K, D, generator=g), dim=1)
    items = [torch.randn(n, D, generator=g) + 3.0 * centers[torch.randint(0, K, (n,), generator=g)]
             for n in (1530, 77, 1530, 33)]
    v = _vlad(centers)
    got = v.generate_multi(items)
    assert isinstance(got, list) and len(got) == 4
    for it, gd in zip(items, got):
        assert_desc_close(gd.numpy()[None], O.anyloc_vlad_generate(it, centers).numpy()[None])


def test_aggft_store_and_cached_vocabulary(tmp_path):
    g = torch.Generator().manual_seed(9)
    K, D, dh, dw = 8, 64, 5, 7
    centers = 0.5 * torch.nn.functional.normalize(torch.randn(K, D, generator=g), dim=1)
    torch.save(centers, tmp_path / "c_centers.pt")
    v = VLAD(K, desc_dim=None, dist_mode="cosine", vlad_mode="hard", cache_dir=str(tmp_path))
    v.fit(None)
    assert v.desc_dim == D
    store = {f"img{i}.jpg": {"ift_dino": torch.randn(1, D, dh, dw, generator=g).numpy()} for i in (10, 2, 1)}
    fts = func_vpr.aggFt(store, None, None, {"desired_height": 70, "desired_width": 98}, "vlad", v, upsample=True)
    keys = ["img1.jpg", "img2.jpg", "img10.jpg"]                # natural order, as natsorted() in the reference
    for kname, ft in zip(keys, fts):
        tok = torch.from_numpy(store[kname]["ift_dino"]).reshape(D, dh * dw).t()
        assert_desc_close(ft[None], O.anyloc_vlad_generate(tok, centers).numpy()[None])
    with pytest.raises(NotImplementedError):
        func_vpr.aggFt(store, None, None, {}, "avg", v)
    with pytest.raises(ValueError):
        VLAD(K, cache_dir=None).fit(None)


def test_get_recall_golden_reference(golden_dir):
    g = np.load(os.path.join(golden_dir, "anyloc_recall.npz"))
    gt = [[int(x) for x in row if x >= 0] for row in g["gt"]]
    recall, per_query, matches = func_vpr.get_recall(g["db"], g["q"], gt, analysis=True, k=int(g["k"]))
    np.testing.assert_array_equal(np.stack([m["img_id_r"] for m in matches]), g["nbrs"])
    np.testing.assert_allclose(recall, g["recall"], rtol=0, atol=1e-12)
    np.testing.assert_array_equal(per_query, g["per_query"])
    qr = func_vpr.convert_to_queries_results_for_map([list(m["img_id_r"]) for m in matches], gt)
    assert abs(func_vpr.calculate_map(qr) - float(g["map"])) < 1e-15
    r2, m2 = func_vpr.get_recall(g["db"], g["q"], gt, k=int(g["k"]))
    np.testing.assert_array_equal(r2, recall)


def test_recall_anyloc_end_to_end(tmp_path):
    """place_rec_main.py:379-391 over DirStore token files: query i is a noisy copy of reference i."""
    from revisit_anything_b200 import place_rec_main, store
    g = torch.Generator().manual_seed(11)
    K, D, dh, dw, n_img = 8, 64, 4, 6, 12
    centers = 0.5 * torch.nn.functional.normalize(torch.randn(K, D, generator=g), dim=1)
    v = _vlad(centers)
    sr, sq = store.DirStore.create(str(tmp_path / "r")), store.DirStore.create(str(tmp_path / "q"))
    toks_r, toks_q = [], []
    for i in range(n_img):
        t = torch.randn(1, D, dh, dw, generator=g)
        store.write_tokens(sr, f"im{i}", t.numpy())
        tq = t + 0.05 * torch.randn(1, D, dh, dw, generator=g)
        store.write_tokens(sq, f"im{i}", tq.numpy())
        toks_r.append(t.reshape(D, -1).t())
        toks_q.append(tq.reshape(D, -1).t())
    gt = [[i] for i in range(n_img)]
    recall, info, im1, im2 = place_rec_main.recall_anyloc(sr, sq, {"desired_height": 56, "desired_width": 84}, v, gt)
    db = O.normalize_feat(np.stack([O.anyloc_vlad_generate(t, centers).numpy() for t in toks_r]))
    qq = O.normalize_feat(np.stack([O.anyloc_vlad_generate(t, centers).numpy() for t in toks_q]))
    want, _, nbrs = O.get_recall(db, qq, gt, k=5)
    np.testing.assert_allclose(recall, want, rtol=0, atol=1e-12)
    np.testing.assert_array_equal(np.stack([m["img_id_r"] for m in info])[:, 0], nbrs[:, 0])
    assert recall[0] == 100.0 and len(im1) == n_img
