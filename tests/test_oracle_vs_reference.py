"""Oracle vs the UNMODIFIED reference functions imported live (build container only; skipped on the
GPU box where /root/reference does not exist).  Randomised inputs beyond the committed goldens."""
import numpy as np
import pytest
import torch

from oracle import ref_shim
from oracle import segvlad_oracle as O
from revisit_anything_b200 import synth

pytestmark = pytest.mark.skipif(not ref_shim.available(), reason="/root/reference not mounted")


@pytest.fixture(scope="module")
def ref():
    return ref_shim.load()


@pytest.mark.parametrize("seed,S,order", [(1, 11, 3), (2, 5, 1), (3, 2, 2), (4, 9, 0)])
def test_aggregation_vs_reference(ref, seed, S, order):
    D, dh, dw = 40, 5, 7
    centers = synth.make_centers(32, D, seed)
    x = O.normalize_tokens(synth.make_tokens(D, dh, dw, seed, centers).reshape(D, -1))
    masks = synth.make_masks(S, 35, 49, seed)
    member = torch.from_numpy(O.mask_to_patch_membership(masks, 70, 98))
    adj = ref.nbrMasksAGGFastSingle(masks, order) if order else None
    want, lab_ref = ref_shim.vlad_single_cpu(ref, x, centers, member, adj)
    got, lab, _ = O.vlad_single(x, centers, member, adj)
    assert torch.equal(lab, lab_ref)
    np.testing.assert_allclose(got.numpy(), want.numpy(), rtol=0, atol=1e-13)
    if order:
        np.testing.assert_array_equal(O.neighbour_adjacency(masks, order), adj.numpy())


@pytest.mark.parametrize("seed", [0, 1, 2])
def test_vote_vs_reference(ref, seed):
    rng = np.random.RandomState(seed)
    n_qimg, segs, Nr, kv = 7, 9, 300, 50
    Nq = n_qimg * segs
    im_inds_ref = np.sort(rng.randint(0, 25, size=Nr)).astype(np.int64)
    matches = np.stack([rng.choice(Nr, kv, replace=False) for _ in range(Nq)]).astype(np.int64)
    sims = (2 - np.sort(rng.uniform(0, 2, (Nq, kv)).astype(np.float32), axis=1)).astype(np.float32)
    if seed == 2:
        sims = (np.round(sims * 4) / 4).astype(np.float32)
    seg_range = [np.arange(i * segs, (i + 1) * segs) for i in range(n_qimg)]
    gt = [list(range(i, i + 3)) for i in range(n_qimg)]
    want = ref.get_matches(matches, gt, sims, seg_range, im_inds_ref, n=5,
                           method="max_seg_topk_wt_borda_Im")
    got = O.get_matches_wt_borda(matches, n_qimg, sims, seg_range, im_inds_ref, n=5)
    assert [list(map(int, p)) for p in got] == [list(map(int, p)) for p in want]
    assert O.calc_recall(got, gt, 5) == ref.calc_recall(want, gt, 5)


@pytest.mark.parametrize("B,D,H,K,ab,alt", [(2, 64, 7, 16, (8.0, 7.0, 1.0), False), (1, 128, 11, 64, (8.0, 7.0, 1.0), True),
                                            (3, 48, 5, 8, (4.0, 2.5, 0.7), False)])
def test_netvlad_antiburst_vs_reference(B, D, H, K, ab, alt):
    """a9: oracle vs the reference's NetVLAD.forward (aggregation.py:266-361) imported live, both vlad branches."""
    mod = ref_shim.load_netvlad_module()
    g = torch.Generator().manual_seed(B * 100 + D + K)
    x = torch.randn(B, D, H, H, generator=g)
    x[:, :, 1, :4] = x[:, :, 1, :1]                                # a burst
    cent = torch.rand(K, D, generator=g)
    W = 9.0 * cent / cent.norm(dim=1, keepdim=True)
    want = ref_shim.netvlad_reference_forward(mod, x, cent, W, ab, for_loop_alt=alt)
    got = O.netvlad_antiburst(x.reshape(B, D, -1), cent, W, ab)
    np.testing.assert_allclose(got.numpy(), want.numpy(), rtol=0, atol=1e-7)


@pytest.mark.parametrize("seed,n_comp,Din", [(0, 16, 64), (1, 40, 128)])
def test_pca_apply_vs_reference(ref, tmp_path, seed, n_comp, Din):
    """a5: oracle vs the reference's apply_pca_transform_from_pkl (func_vpr.py:1419-1443) on a freshly fitted whitening PCA,
    plus normalizeFeat (:1673-1676).  The installed sklearn (1.9) applies the mean as an fp32 bias; the oracle follows the
    reference's pinned 1.3.2 formula in fp64 -- 2e-5 relative / 5e-8 absolute covers the difference (see test_oracle_golden)."""
    import pickle

    from sklearn.decomposition import PCA
    rng = np.random.RandomState(seed)
    train = (rng.randn(400, Din) @ rng.randn(Din, Din) * 0.05).astype(np.float32)
    pca = PCA(n_components=n_comp, whiten=True, svd_solver="arpack").fit(train)
    path = tmp_path / "pca.pkl"
    with open(path, "wb") as fh:
        pickle.dump(pca, fh)
    X = rng.randn(23, Din) * 0.05
    want = ref.apply_pca_transform_from_pkl(torch.from_numpy(X), str(path)).numpy()
    got = O.pca_apply(X, pca.mean_, pca.components_, pca.explained_variance_)
    np.testing.assert_allclose(got, want, rtol=2e-5, atol=5e-8)
    np.testing.assert_allclose(O.normalize_feat(got), ref.normalizeFeat(want.copy()), rtol=2e-5, atol=5e-8)


@pytest.mark.parametrize("S,order,seed", [(25, 3, 7), (9, 2, 8), (4, 1, 9), (3, 3, 10), (2, 2, 11), (1, 1, 12)])
def test_adjacency_vs_reference(ref, S, order, seed):
    """a4: SuperSegment adjacency (Delaunay neighbours to the given order, func_vpr.py:1309-1347) incl. the S <= 3 special case."""
    masks = synth.make_masks(S, 48, 64, seed)
    np.testing.assert_array_equal(O.neighbour_adjacency(masks, order), ref.nbrMasksAGGFastSingle(masks, order).numpy())


@pytest.mark.parametrize("seed", [0, 1])
@pytest.mark.parametrize("method", ["max_sim", "max_seg_sim"])
def test_host_get_matches_variants_vs_reference(ref, seed, method):
    """a7 boundary: the signature's DEFAULT method "max_sim" (func_vpr.py:87-93) and "max_seg_sim" (:103-117) work on the
    top-1 match per query segment; the drop-in keeps them on the host.  Compared with the unmodified reference, including
    a non-contiguous segRangeQuery."""
    from revisit_anything_b200 import func_vpr as ours
    rng = np.random.RandomState(seed)
    n_qimg, Nr = 9, 400
    lens = rng.randint(3, 70, size=n_qimg)
    Nq = int(lens.sum())
    perm = rng.permutation(Nq) if seed else np.arange(Nq)
    off = np.concatenate([[0], np.cumsum(lens)])
    seg_range = [perm[off[i]:off[i + 1]] for i in range(n_qimg)]
    im_inds_ref = np.sort(rng.randint(0, 30, size=Nr)).astype(np.int64)
    matches = rng.randint(0, Nr, size=Nq).astype(np.int64)
    sims = rng.permutation(Nq).astype(np.float32) / Nq          # distinct values: argsort order is unique
    gt = [[0]] * n_qimg
    want = ref.get_matches(matches, gt, sims, seg_range, im_inds_ref, n=5, method=method)
    got = ours.get_matches(matches, gt, sims, seg_range, im_inds_ref, n=5, method=method)
    assert [list(map(int, p)) for p in got] == [list(map(int, p)) for p in want]
    if method == "max_sim":                                     # the default argument must be this branch
        dflt = ours.get_matches(matches, gt, sims, seg_range, im_inds_ref, n=5)
        assert [list(map(int, p)) for p in dflt] == [list(map(int, p)) for p in want]
