"""Drop-in mirror of the reference's `place_rec_main.py` hot path: `recall_segloc` (place_rec_main.py:44-96)
plus the per-image descriptor loop (place_rec_main.py:244-355) as a batched, device-resident function.

`recall_segloc` keeps the reference signature and return value (Recall@1..5 list).  The faiss
IndexFlatL2 search is replaced by the tcgen05 kNN, `get_matches` by the vote kernel.
"""
from __future__ import annotations

import os
import pickle
from typing import List, Optional, Sequence

import numpy as np
import torch

from . import engine, func_vpr
from ._lib import TOKENS_DN

K_SEARCH = 200   # place_rec_main.py:56,60
K_VOTE = 50      # place_rec_main.py:78-79
N_PRED = 5       # place_rec_main.py:84


def _dev():
    if not torch.cuda.is_available():
        raise RuntimeError("segvlad: no CUDA device (the SegVLAD hot path has no CPU fallback)")
    return torch.device("cuda", torch.cuda.current_device())


def search_and_vote(seg_ft_ref: torch.Tensor, seg_ft_qry: torch.Tensor, seg_range_q, im_inds_ref, n_qimg: int,
                    pca: bool, k_search: int = K_SEARCH, k_vote: int = K_VOTE, n_pred: int = N_PRED):
    """Device-resident core of recall_segloc: returns (d2 [Nq,k], idx [Nq,k], VoteResult)."""
    dev = _dev()
    host_f32 = (not seg_ft_ref.is_cuda and not seg_ft_qry.is_cuda and not pca
                and seg_ft_ref.dtype == torch.float32 and seg_ft_qry.dtype == torch.float32)
    if host_f32:
        # CPU fp32 descriptors (pin them for full PCIe rate): stream the bank while scanning it
        d2, idx, _, _ = engine.knn_from_host(seg_ft_qry, seg_ft_ref, k_search)
    else:
        f64 = seg_ft_ref.dtype == torch.float64
        prep = (lambda x: engine.Bank.prepare_f64(x.to(dev), normalize_rows=pca)) if f64 or pca else \
               (lambda x: engine.Bank.prepare(x.to(dev)))
        rbank = prep(seg_ft_ref)
        qbank = prep(seg_ft_qry)
        d2, idx = engine.knn(qbank, rbank, k_search)
    perm, off = func_vpr._ranges_to_offsets(seg_range_q, n_qimg)
    qrow = None if perm is None else torch.from_numpy(perm.astype(np.int32)).to(dev)
    im = np.asarray(im_inds_ref).astype(np.int64)
    n_rimg = int(im.max()) + 1
    res = engine.vote(idx, d2, torch.from_numpy(off.astype(np.int32)).to(dev), torch.from_numpy(im).to(dev), n_rimg,
                      n_pred=n_pred, k_vote=k_vote, sims_is_d2=True, qrow_index=qrow)
    return d2, idx, res


def recall_segloc(workdir, dataset_name, experiment_config, experiment_name, segFtVLAD1, segFtVLAD2, gt, segRange2,
                  imInds1, map_calculate, domain, save_results=True):
    """place_rec_main.py:44-96.  segFtVLAD1/2: [Nr,D]/[Nq,D] tensors (CPU or CUDA; fp64 like the reference or
    fp32); with experiment_config['pca'] rows are L2-normalised first (normalizeFeat).  The descriptor
    dimension is taken from the data (the reference hard-codes 1024 / 49152 for faiss)."""
    pca = bool(experiment_config["pca"])
    d2, idx, res = search_and_vote(segFtVLAD1, segFtVLAD2, segRange2, imInds1, len(gt), pca)
    if save_results:
        out_folder = f"{workdir}/results/global/"
        os.makedirs(f"{out_folder}/{experiment_name}", exist_ok=True)
        pkl = f"{out_folder}/{experiment_name}/{dataset_name}_matches_sims_domain_{domain}__{experiment_config['results_pkl_suffix']}"
        with open(pkl, "wb") as fh:   # same keys as the reference: 'sims' holds the squared distances
            pickle.dump({"sims": d2.cpu().numpy(), "matches": idx.cpu().numpy()}, fh)
        print(f"Results saved to {pkl}")
    p = res.preds.cpu().numpy()
    max_seg_preds = [p[i][p[i] >= 0].astype(np.int64) for i in range(len(gt))]
    max_seg_recalls = func_vpr.calc_recall(max_seg_preds, gt, N_PRED)
    print("VLAD + PCA Results \n ")
    if map_calculate:
        queries_results = func_vpr.convert_to_queries_results_for_map(max_seg_preds, gt)
        print(f"Mean Average Precision (mAP): {func_vpr.calculate_map(queries_results)}")
    print("Max Seg Logs: ", max_seg_recalls)
    return max_seg_recalls


def build_segment_descriptors(tokens: Sequence[torch.Tensor], masks: Sequence[Sequence[np.ndarray]],
                              c_centers: torch.Tensor, cfg: dict, order: int, desc_dim: int = 1536,
                              batch_images: int = 32, out_dtype=torch.float64, adjacency: Optional[Sequence] = None,
                              pca_model_path: Optional[str] = None):
    """Batched equivalent of the per-image loop place_rec_main.py:244-281 (reference side) / :309-352 (query
    side): for every image, SuperSegment adjacency on the host (scipy, as in the reference), then ONE batched
    aggregation launch per `batch_images` images; descriptors stay on the GPU.
    tokens[i]: [1,D,dh,dw] fp32 (CPU or CUDA); masks[i]: list of [Hm,Wm] bool.
    With `pca_model_path` (experiment_config['pca'], place_rec_main.py:261-272) every batch is projected on the device
    with the whitening PCA (func_vpr.apply_pca_transform_from_pkl semantics) before it is kept, so the [S, K*D] fp64
    block never leaves the GPU and only [S, n_components] rows accumulate.
    Returns (segFtVLAD [S_total, K*D or n_components] CUDA, imInds [S_total] int64 numpy)."""
    dev = _dev()
    H, W = cfg["desired_height"], cfg["desired_width"]
    N = (H // 14) * (W // 14)
    centers = c_centers.to(dev)
    outs, im_inds = [], []
    for b0 in range(0, len(tokens), batch_images):
        b1 = min(len(tokens), b0 + batch_images)
        tok = torch.stack([tokens[i].reshape(desc_dim, N) for i in range(b0, b1)]).to(dev, non_blocking=True)
        # ONE upload and one membership / centroid launch for the masks of the whole batch (same mask resolution)
        counts = [len(masks[i]) for i in range(b0, b1)]
        flat = [m for i in range(b0, b1) for m in masks[i]]
        if not flat:
            continue
        shapes = {np.asarray(m).shape for m in flat}
        cents = None
        if len(shapes) == 1:
            ms = torch.from_numpy(np.ascontiguousarray(np.stack(flat))).to(dev, non_blocking=True)
            bits = [engine.mask_to_membership(ms, H, W, 14)]
            if order and adjacency is None:
                cents = engine.mask_centroids(ms).cpu().numpy()          # [S_batch, 2]: input of the host Delaunay
        else:
            bits = [engine.mask_to_membership(torch.from_numpy(np.ascontiguousarray(np.asarray(masks[i]))).to(dev), H, W, 14)
                    for i in range(b0, b1) if len(masks[i])]
        adjs, s0 = [], 0
        for j, i in enumerate(range(b0, b1)):
            if adjacency is not None:
                adjs.append(None if adjacency[i] is None else torch.as_tensor(adjacency[i]))
            elif order:
                c = None if cents is None else cents[s0:s0 + counts[j]]
                adjs.append(func_vpr.nbrMasksAGGFastSingle(masks[i], order, centroids=c))
            else:
                adjs.append(None)
            s0 += counts[j]
            im_inds.append(np.full(len(masks[i]), i, dtype=np.int64))
        if pca_model_path is not None:
            comp, mean, ev = func_vpr._load_pca(pca_model_path)
            if engine.pca_fusable(desc_dim, centers.shape[0], comp.shape[0]):
                # fused: the aggregation writes (descriptor - mean) as the projection's bf16 operand planes; the fp64
                # [S, K*D] block of place_rec_main.py:259-272 never exists
                gd = engine.aggregate_project_pca(tok, N, desc_dim, TOKENS_DN, centers, torch.cat(bits), counts,
                                                  adjs if order else None, comp, mean, ev)
            else:
                gd = engine.aggregate_batch(tok, N, desc_dim, TOKENS_DN, centers, torch.cat(bits), counts,
                                            adjs if order else None, out_dtype=out_dtype)
                gd = func_vpr.apply_pca_transform_from_pkl(gd, pca_model_path, device_out=True)
        else:
            gd = engine.aggregate_batch(tok, N, desc_dim, TOKENS_DN, centers, torch.cat(bits), counts,
                                        adjs if order else None, out_dtype=out_dtype)
        outs.append(gd)
    return torch.cat(outs), np.concatenate(im_inds)


def recall_anyloc(dino_r_path, dino_q_path, cfg, vlad, gt, topk_value=5):
    """AnyLoc-VLAD-DINOv2 baseline branch, place_rec_main.py:379-391: whole-image VLAD of every reference / query image
    (`func_vpr.aggFt(..., 'vlad')`), normalizeFeat, Recall@1..topk through `func_vpr.get_recall`.
    `dino_*_path`: HDF5 path (needs h5py) or any mapping with the same indexing (store.DirStore).
    Returns (recall in percent, match_info, imFts1_vlad, imFts2_vlad)."""
    im1 = func_vpr.aggFt(dino_r_path, None, None, cfg, "vlad", vlad, upsample=True)
    im2 = func_vpr.aggFt(dino_q_path, None, None, cfg, "vlad", vlad, upsample=True)
    recall, match_info = func_vpr.get_recall(func_vpr.normalizeFeat(im1), func_vpr.normalizeFeat(im2), gt, k=topk_value)
    return recall, match_info, im1, im2
