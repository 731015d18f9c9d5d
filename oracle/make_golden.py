"""Generate tests/golden/*.npz by RUNNING THE REFERENCE ITSELF in the build container.

    python -m oracle.make_golden            (needs /root/reference; CPU only)

The reference has no tests and no golden vectors, so the known answers are the
outputs of its own functions on small seeded inputs:

* vlad_matmuls_per_cluster, nbrMasksAGGFastSingle, get_matches, weighted_borda_count, calc_recall,
  normalizeFeat: called unmodified (oracle/ref_shim.py import).
* seg_vlad_gpu_single_img / vlad_single hard-code .to('cuda') (func_vpr.py:1082-1096, 1145) and call
  vlad_matmuls_per_cluster with its default device='cuda' (:1181): their
  source text is taken with inspect.getsource, the literal 'cuda' is replaced by 'cpu', and the
  patched functions are executed -- same lines, same arithmetic, CPU device.

The fixtures are small (< 1 MB total) and are what the CUDA path and the oracle are both held to on
the GPU box, where /root/reference does not exist.
"""
from __future__ import annotations

import inspect
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from oracle import ref_shim  # noqa: E402
from revisit_anything_b200 import synth  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden")


def _cpu_patched(ref):
    ns = dict(vars(ref))
    for name in ("vlad_matmuls_per_cluster", "vlad_single", "seg_vlad_gpu_single_img"):
        src = inspect.getsource(getattr(ref, name))
        src = src.replace("'cuda'", "'cpu'").replace('"cuda"', '"cpu"')
        exec(compile(src, f"<reference {name} patched cuda->cpu>", "exec"), ns)
    return ns["seg_vlad_gpu_single_img"], ns["vlad_single"]


def agg_case(ref, seg_vlad_img, name, D, H, W, S, order, seed, centers=None, col_stride=1,
             normalized=True):
    cfg = {"desired_height": H, "desired_width": W}
    dh, dw = H // 14, W // 14
    if centers is None:
        centers = synth.make_centers(32, D, seed)
    tokens = synth.make_tokens(D, dh, dw, seed, centers, normalized=normalized)
    masks = synth.make_masks(S, H // 2, W // 2, seed)
    adj = ref.nbrMasksAGGFastSingle(masks, order) if order else None
    ind = np.empty((H, W, 2), dtype="int32")
    for i in range(H):
        for j in range(W):
            ind[i, j] = (np.clip(i // 14, 0, dh - 1), np.clip(j // 14, 0, dw - 1))
    ind_flat = torch.tensor(np.ravel_multi_index(ind.reshape(-1, 2).T, (dh, dw)))
    gd = seg_vlad_img(ind_flat, ind, tokens.clone(), "img", masks, centers, cfg, desc_dim=D, adj_mat=adj)
    gd = gd.numpy()
    cols = np.arange(0, gd.shape[1], col_stride)
    np.savez_compressed(
        os.path.join(OUT, name + ".npz"),
        tokens=tokens.numpy(), masks=np.asarray(masks), centers=centers.numpy(),
        adj=(adj.numpy() if adj is not None else np.zeros((0, 0), bool)),
        order=order, H=H, W=W, cols=cols, vlad_cols=gd[:, cols],
        row_block_norms=np.linalg.norm(gd.reshape(gd.shape[0], 32, D), axis=2))
    print(name, gd.shape, "->", cols.size, "cols kept")


def vote_case(ref, name, n_qimg, segs, n_rimg, rsegs, seed, kv=50, n=5, dup_sims=False):
    rng = np.random.RandomState(seed)
    Nq, Nr = n_qimg * segs, n_rimg * rsegs
    im_inds_ref = (np.arange(Nr) // rsegs).astype(np.int64)
    matches = np.empty((Nq, kv), dtype=np.int64)
    for q in range(Nq):
        true_img = (q // segs) % n_rimg
        pool_true = np.where(im_inds_ref == true_img)[0]
        m = rng.choice(Nr, size=kv, replace=False)
        hit = rng.rand(kv) < 0.08
        m[hit] = rng.choice(pool_true, size=int(hit.sum()))
        matches[q] = m
    d2 = np.sort(rng.uniform(0.2, 1.9, size=(Nq, kv)).astype(np.float32), axis=1)
    if dup_sims:  # many exactly equal sims -> exercises tie order (insertion order) of the vote
        d2 = (np.round(d2 * 8) / 8).astype(np.float32)
    sims = (2 - d2).astype(np.float32)
    seg_range = [np.arange(i * segs, (i + 1) * segs) for i in range(n_qimg)]
    gt = [list(np.arange(i - 1, i + 2)) if i % 3 else list(np.arange(i + 3, i + 6)) for i in range(n_qimg)]
    preds = ref.get_matches(matches, gt, sims, seg_range, im_inds_ref, n=n,
                            method="max_seg_topk_wt_borda_Im")
    full = []
    lo, hi = np.min(sims), np.max(sims)
    for i in range(n_qimg):
        m_t = matches[seg_range[i]].T.tolist()
        s_t = ((sims[seg_range[i]].T - lo) / (hi - lo)).tolist()
        pairs = [list(zip(im_inds_ref[m_t[k]], s_t[k])) for k in range(len(s_t))]
        full.append(np.asarray(ref.weighted_borda_count(*pairs), dtype=np.int64))
    preds_cnt = ref.get_matches(matches, gt, sims, seg_range, im_inds_ref, n=n, method="max_seg_topk")
    recalls = ref.calc_recall(preds, gt, n)
    pad = lambda lst, w: np.array([list(p) + [-1] * (w - len(p)) for p in lst], dtype=np.int64)
    wmax = max(len(f) for f in full)
    np.savez_compressed(
        os.path.join(OUT, name + ".npz"), matches=matches, sims=sims, segs=segs,
        im_inds_ref=im_inds_ref, n_qimg=n_qimg, n=n, preds=pad(preds, n),
        full_ranking=pad(full, wmax), preds_bincount=pad(preds_cnt, n),
        recalls=np.asarray(recalls), gt=np.asarray(gt))
    print(name, "recalls", recalls)


def adjacency_case(ref, name):
    out = {}
    for S, order, seed in [(12, 1, 1), (12, 2, 1), (12, 3, 1), (40, 3, 2), (3, 3, 3), (2, 1, 4), (1, 2, 5)]:
        masks = synth.make_masks(S, 60, 80, seed)
        adj = ref.nbrMasksAGGFastSingle(masks, order)
        out[f"masks_S{S}_o{order}"] = np.asarray(masks)
        out[f"adj_S{S}_o{order}"] = adj.numpy()
    np.savez_compressed(os.path.join(OUT, name + ".npz"), **out)
    print(name, "ok")


def pca_case(ref, name):
    """Reference apply_pca_transform_from_pkl (func_vpr.py:1419-1443) on a small whitening PCA fitted with sklearn."""
    import pickle
    import tempfile

    from sklearn.decomposition import PCA
    rng = np.random.RandomState(31)
    train = (rng.randn(300, 96) @ rng.randn(96, 96) * 0.05).astype(np.float32)     # fitted on fp32 like place_rec_pca.py:380
    pca = PCA(n_components=24, whiten=True, svd_solver="arpack").fit(train)
    X = rng.randn(11, 96) * 0.05                                                    # fp64 descriptors
    with tempfile.NamedTemporaryFile(suffix=".pkl", delete=False) as fh:
        pickle.dump(pca, fh)
        path = fh.name
    Y = ref.apply_pca_transform_from_pkl(torch.from_numpy(X), path).numpy()
    os.unlink(path)
    np.savez_compressed(os.path.join(OUT, name + ".npz"), X=X, Y=Y, components=pca.components_, mean=pca.mean_,
                        explained_variance=pca.explained_variance_)
    print(name, Y.shape, Y.dtype)


def anyloc_case(ref):
    """Reference utilities.VLAD.generate (hard assignment, whole image) and func_vpr.get_recall / calculate_map.
    fast_pytorch_kmeans is absent offline: VLAD.kmeans is a stub applying the library's published cosine rule."""
    import sys

    import torch
    util = sys.modules["utilities"]

    class _Predict:
        def __init__(self, c):
            self.centroids = c

        def predict(self, x):
            a = x / (x.norm(dim=-1, keepdim=True) + 1e-8)
            b = self.centroids / (self.centroids.norm(dim=-1, keepdim=True) + 1e-8)
            return (a @ b.t()).max(dim=-1)[1]

    g = torch.Generator().manual_seed(41)
    K, D, N, B = 8, 64, 300, 5
    centers = 0.5 * torch.nn.functional.normalize(torch.randn(K, D, generator=g), dim=1)
    centers[5] = 3.0 * centers[5]          # a far centre that attracts few tokens; cluster 7 made empty below
    tokens = torch.randn(B, N, D, generator=g) + 2.0 * centers[torch.randint(0, K - 1, (B, N), generator=g)]
    vlad = util.VLAD(K, desc_dim=None, dist_mode="cosine", vlad_mode="hard", cache_dir=None)
    vlad.c_centers, vlad.kmeans, vlad.desc_dim = centers, _Predict(centers), D
    out = torch.stack([vlad.generate(tokens[b]) for b in range(B)])
    np.savez_compressed(os.path.join(OUT, "anyloc_vlad.npz"), tokens=tokens.numpy(), centers=centers.numpy(),
                        vlad=out.numpy())
    print("anyloc_vlad", out.shape, out.dtype)

    rng = np.random.RandomState(43)
    n_db, n_q, Dg, k = 60, 25, 48, 5
    db = ref.normalizeFeat(rng.randn(n_db, Dg).astype(np.float32))
    q = ref.normalizeFeat((db[rng.randint(0, n_db, n_q)] + 0.6 * rng.randn(n_q, Dg)).astype(np.float32))
    gt = [list(range(max(0, i * 2 - 2), min(n_db, i * 2 + 3))) for i in range(n_q)]
    gt[3] = []
    recall, per_query, matches = ref.get_recall(db, q, gt, analysis=True, k=k)
    nbrs = np.stack([m["img_id_r"] for m in matches])
    qr = ref.convert_to_queries_results_for_map([list(r) for r in nbrs], gt)
    np.savez_compressed(os.path.join(OUT, "anyloc_recall.npz"), db=db, q=q, k=k,
                        gt=np.array([np.array(x + [-1] * (8 - len(x))) for x in gt]), recall=np.asarray(recall),
                        per_query=np.asarray(per_query), nbrs=nbrs, map=ref.calculate_map(qr),
                        ap=np.array([ref.calculate_ap(r) for r in qr], dtype=np.float64))
    print("anyloc_recall", recall)


def netvlad_case(name="netvlad_antiburst"):
    """a9: the reference's NetVLAD.forward with anti-burst weighting (VLAD-BuFF/models/aggregators/aggregation.py:266-361,
    getWeights :148-162), per-cluster loop branch, CPU fp32; includes a burst (repeated tokens) and non-default ab params."""
    mod = ref_shim.load_netvlad_module()
    g = torch.Generator().manual_seed(31)
    B, D, H, K = 2, 96, 9, 32
    x = torch.randn(B, D, H, H, generator=g)
    x[:, :, 0, :3] = x[:, :, 0, :1]
    cent = torch.rand(K, D, generator=g)                          # aggregation.py:216 init
    W = 12.0 * cent / cent.norm(dim=1, keepdim=True)               # init_params-style alpha * c_hat (:245-256)
    outs = {}
    for tag, ab in (("default", (8.0, 7.0, 1.0)), ("alt", (5.0, 3.0, 0.5))):
        outs["out_" + tag] = ref_shim.netvlad_reference_forward(mod, x, cent, W, ab).numpy()
        outs["ab_" + tag] = np.asarray(ab, dtype=np.float32)
    np.savez_compressed(os.path.join(OUT, name + ".npz"), x=x.numpy(), centroids=cent.numpy(), conv_weight=W.numpy(), **outs)


def main():
    os.makedirs(OUT, exist_ok=True)
    netvlad_case()
    ref = ref_shim.load()
    anyloc_case(ref)
    seg_vlad_img, _ = _cpu_patched(ref)
    torch.manual_seed(17)
    np.random.seed(17)
    # small-D cases (full output kept)
    agg_case(ref, seg_vlad_img, "agg_small_o2", D=48, H=84, W=112, S=7, order=2, seed=11)
    agg_case(ref, seg_vlad_img, "agg_small_o0", D=48, H=84, W=112, S=5, order=0, seed=12)
    agg_case(ref, seg_vlad_img, "agg_small_S3", D=32, H=70, W=98, S=3, order=3, seed=13)
    agg_case(ref, seg_vlad_img, "agg_unnorm_o1", D=64, H=98, W=126, S=9, order=1, seed=14,
             normalized=False)
    # real vocabulary (17places 'indoor' domain, place_rec_global_config.py:35), D_t = 1536
    voc = os.path.join(ref_shim.REF_ROOT, "cache/vocabulary/dinov2_vitg14/l31_value_c32/indoor/c_centers.pt")
    centers = torch.load(voc, map_location="cpu").float().contiguous()
    agg_case(ref, seg_vlad_img, "agg_realvocab_o3", D=1536, H=84, W=112, S=6, order=3, seed=15,
             centers=centers, col_stride=61)
    vote_case(ref, "vote_a", n_qimg=6, segs=10, n_rimg=40, rsegs=10, seed=21)
    vote_case(ref, "vote_ties", n_qimg=5, segs=7, n_rimg=12, rsegs=6, seed=22, dup_sims=True)
    adjacency_case(ref, "adjacency")
    pca_case(ref, "pca_apply")
    # normalizeFeat
    x = np.random.RandomState(5).randn(9, 33)
    np.savez_compressed(os.path.join(OUT, "normalize_feat.npz"), x=x, y=ref.normalizeFeat(x.copy()))


if __name__ == "__main__":
    main()
