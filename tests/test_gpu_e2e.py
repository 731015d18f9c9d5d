"""End to end on the GPU: descriptors -> recall_segloc (kNN + vote + Recall@k) vs the oracle pipeline; and the
full extract-less pipeline tokens+masks -> aggregate -> match -> vote on a tiny synthetic place-recognition set."""
import numpy as np
import pytest
import torch

from gpu_util import assert_desc_close
from oracle import segvlad_oracle as O
from revisit_anything_b200 import place_rec_main, synth

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("pca", [False, True])
def test_recall_segloc_matches_oracle(tmp_path, pca):
    q, r, imq, imr = synth.make_structured_bank(n_ref_img=60, n_qry_img=40, segs_per_img=25, D=256, seed=1, noise=1.1)
    if pca:                                    # un-normalised fp64 "PCA outputs": recall_segloc must normalise
        q = (q.double() * 3.7)
        r = (r.double() * 0.4)
    seg_range = [np.where(imq == i)[0] for i in range(40)]
    gt = [list(np.arange(i - 1, i + 2)) for i in range(40)]
    cfg = {"pca": pca, "results_pkl_suffix": "t.pkl"}
    rec = place_rec_main.recall_segloc(str(tmp_path), "synth", cfg, "exp", r, q, gt, seg_range, imr, False, "x",
                                       save_results=True)
    rec_o, preds_o, (D2o, Io) = O.recall_segloc(r.numpy(), q.numpy(), gt, seg_range, imr, pca)
    assert rec == rec_o
    import pickle
    saved = pickle.load(open(tmp_path / "results/global/exp/synth_matches_sims_domain_x__t.pkl", "rb"))
    assert saved["sims"].shape == (1000, 200) and saved["matches"].dtype == np.int64
    np.testing.assert_allclose(saved["sims"], D2o, rtol=1e-5, atol=2e-6)
    assert (saved["matches"] == Io).mean() > 0.99
    d2, idx, res = place_rec_main.search_and_vote(r, q, seg_range, imr, 40, pca)
    p = res.preds.cpu().numpy()
    assert [list(p[i][p[i] >= 0]) for i in range(40)] == [[int(x) for x in pr] for pr in preds_o]


def test_full_pipeline_tokens_to_recall():
    # 12 reference + 12 query images (query i = ref i with token noise + jittered masks), 196x266 px, D_t=128, K=32
    D, H, W, K, order = 128, 196, 266, 32, 2
    dh, dw = H // 14, W // 14
    cfg = {"desired_height": H, "desired_width": W}
    centers = synth.make_centers(K, D, 9)
    g = torch.Generator().manual_seed(0)
    toks_r, toks_q, masks_r, masks_q = [], [], [], []
    def margin_ok(t):
        return float(O.assign_labels(O.normalize_tokens(t.reshape(D, -1)), centers)[1].min()) > 1e-4
    for i in range(12):
        seed = 500 + i
        while True:          # assignment near-ties are not reproducible in fp32 (reference included): avoid them
            t = synth.make_tokens(D, dh, dw, seed, centers)
            tq = t + 0.02 * torch.randn(t.shape, generator=g)
            if margin_ok(t) and margin_ok(tq):
                break
            seed += 1000
        m = synth.make_masks(10 + i % 4, H // 2, W // 2, 600 + i)
        toks_r.append(t)
        masks_r.append(m)
        toks_q.append(tq)
        masks_q.append(synth.jitter_masks(m, 700 + i, px=2))
    ref, im_r = place_rec_main.build_segment_descriptors(toks_r, masks_r, centers, cfg, order, desc_dim=D, batch_images=5)
    qry, im_q = place_rec_main.build_segment_descriptors(toks_q, masks_q, centers, cfg, order, desc_dim=D, batch_images=12)
    # oracle descriptors for the same images
    def oracle_desc(toks, masks):
        out = []
        for t, m in zip(toks, masks):
            adj = torch.from_numpy(O.neighbour_adjacency(m, order))
            v, _, margin, _ = O.seg_vlad_single_img(t, m, centers, cfg, adj)
            out.append(v)
        return torch.cat(out)
    ref_o, qry_o = oracle_desc(toks_r, masks_r), oracle_desc(toks_q, masks_q)
    assert_desc_close(ref.cpu().numpy(), ref_o.numpy())
    assert_desc_close(qry.cpu().numpy(), qry_o.numpy())
    seg_range = [np.where(im_q == i)[0] for i in range(12)]
    gt = [[i] for i in range(12)]
    k = min(200, ref.shape[0])
    d2, idx, res = place_rec_main.search_and_vote(ref, qry, seg_range, im_r, 12, pca=False, k_search=k)
    rec_o, preds_o, _ = O.recall_segloc(ref_o.numpy(), qry_o.numpy(), gt, seg_range, im_r, False, k_search=k)
    p = res.preds.cpu().numpy()
    preds = [list(p[i][p[i] >= 0]) for i in range(12)]
    assert preds == [[int(x) for x in pr] for pr in preds_o]
    assert O.calc_recall(preds, gt, 5) == rec_o and rec_o[0] >= 0.9


def test_full_pipeline_with_pca(tmp_path):
    # the reference's default experiment (exp0_global_SegLoc_VLAD_PCA_o3): aggregate -> whitening PCA -> normalizeFeat ->
    # search -> vote, all on the device; PCA fitted here with sklearn on the oracle's reference descriptors
    import pickle

    from sklearn.decomposition import PCA
    D, H, W, K, order = 64, 140, 182, 32, 1
    dh, dw = H // 14, W // 14
    cfg = {"desired_height": H, "desired_width": W}
    centers = synth.make_centers(K, D, 19)
    g = torch.Generator().manual_seed(1)
    toks_r, toks_q, masks_r, masks_q = [], [], [], []
    for i in range(10):
        seed = 900 + i
        while True:
            t = synth.make_tokens(D, dh, dw, seed, centers)
            tq = t + 0.02 * torch.randn(t.shape, generator=g)
            ok = lambda z: float(O.assign_labels(O.normalize_tokens(z.reshape(D, -1)), centers)[1].min()) > 1e-4
            if ok(t) and ok(tq):
                break
            seed += 1000
        m = synth.make_masks(8 + i % 3, H // 2, W // 2, 950 + i)
        toks_r.append(t); masks_r.append(m); toks_q.append(tq); masks_q.append(synth.jitter_masks(m, 970 + i, px=2))

    def oracle_desc(toks, masks):
        out = []
        for t, m in zip(toks, masks):
            adj = torch.from_numpy(O.neighbour_adjacency(m, order))
            out.append(O.seg_vlad_single_img(t, m, centers, cfg, adj)[0])
        return torch.cat(out)
    ref_o, qry_o = oracle_desc(toks_r, masks_r), oracle_desc(toks_q, masks_q)
    pca = PCA(n_components=32, whiten=True, svd_solver="arpack").fit(ref_o.float().numpy())
    path = str(tmp_path / "pca.pkl")
    pickle.dump(pca, open(path, "wb"))
    ref, im_r = place_rec_main.build_segment_descriptors(toks_r, masks_r, centers, cfg, order, desc_dim=D, pca_model_path=path)
    qry, im_q = place_rec_main.build_segment_descriptors(toks_q, masks_q, centers, cfg, order, desc_dim=D, pca_model_path=path)
    assert ref.shape[1] == 32 and ref.is_cuda
    ref_p = O.pca_apply(ref_o.numpy(), pca.mean_, pca.components_, pca.explained_variance_)
    qry_p = O.pca_apply(qry_o.numpy(), pca.mean_, pca.components_, pca.explained_variance_)
    np.testing.assert_allclose(ref.cpu().numpy(), ref_p, rtol=1e-4, atol=1e-6)     # whitening amplifies the 1e-7 input error
    seg_range = [np.where(im_q == i)[0] for i in range(10)]
    gt = [[i] for i in range(10)]
    k = min(200, ref.shape[0])
    d2, idx, res = place_rec_main.search_and_vote(ref, qry, seg_range, im_r, 10, pca=True, k_search=k)
    rec_o, preds_o, _ = O.recall_segloc(ref_p, qry_p, gt, seg_range, im_r, True, k_search=k)
    p = res.preds.cpu().numpy()
    preds = [list(p[i][p[i] >= 0]) for i in range(10)]
    assert [pp[0] for pp in preds] == [int(pr[0]) for pr in preds_o]              # top-1 agrees (descriptors differ by ~1e-6)
    assert O.calc_recall(preds, gt, 5)[0] == rec_o[0]
