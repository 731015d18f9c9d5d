"""GPU parity of the PCA-whitening projection (row f1) vs the golden vector produced by the reference's
apply_pca_transform_from_pkl and vs the oracle at the real shape (49152 -> 1024)."""
import os
import pickle

import numpy as np
import pytest
import torch

from oracle import segvlad_oracle as O
from revisit_anything_b200 import engine, func_vpr

pytestmark = pytest.mark.gpu


def _cmp_tc(got, want):
    """Tensor-core path: 1e-5 relative (north_star tolerance for descriptors) with an absolute floor of 1e-6 x the row's
    largest element for outputs that are ~0 by cancellation of the 49152-term sum."""
    got, want = np.asarray(got, dtype=np.float64), np.asarray(want, dtype=np.float64)
    floor = 1e-6 * np.abs(want).max(axis=-1, keepdims=True)
    err = np.abs(got - want)
    bad = err > 1e-5 * np.abs(want) + floor
    assert not bad.any(), f"{int(bad.sum())} elements out of tolerance, max abs err {err.max():.3e}"
    rel = np.linalg.norm(got - want, axis=-1) / np.linalg.norm(want, axis=-1)
    assert rel.max() < 5e-6, f"row-wise relative error {rel.max():.3e}"


@pytest.mark.parametrize("mode", ["tc", "fp64"])
def test_golden_reference_vector(golden_dir, monkeypatch, mode):
    monkeypatch.setenv("SEGVLAD_PCA_TC", "1" if mode == "tc" else "0")
    g = np.load(os.path.join(golden_dir, "pca_apply.npz"))
    y = engine.pca_project(torch.from_numpy(g["X"]).cuda(), torch.from_numpy(g["components"]).cuda(),
                           torch.from_numpy(g["mean"]).cuda(), torch.from_numpy(g["explained_variance"]).cuda())
    want = O.pca_apply(g["X"], g["mean"], g["components"], g["explained_variance"])
    if mode == "tc":
        _cmp_tc(y.cpu().numpy(), g["Y"])
        _cmp_tc(y.cpu().numpy(), want)
        return
    # golden = reference function with sklearn 1.9 (fp32 bias term, ~1e-8 abs); kernel = sklearn-1.3.2 formula in fp64
    np.testing.assert_allclose(y.cpu().numpy(), g["Y"], rtol=2e-5, atol=5e-8)
    np.testing.assert_allclose(y.cpu().numpy(), want, rtol=1e-10, atol=1e-12)         # fp64 on both sides


@pytest.mark.parametrize("mode", ["tc", "fp64"])
def test_real_shape_vs_oracle_and_normalize(monkeypatch, mode):
    monkeypatch.setenv("SEGVLAD_PCA_TC", "1" if mode == "tc" else "0")
    rng = np.random.RandomState(3)
    S, Din, Dout = 37, 49152, 1024
    X = rng.randn(S, Din) / np.sqrt(Din)
    W = (rng.randn(Dout, Din) / np.sqrt(Din)).astype(np.float32)
    mu = rng.randn(Din).astype(np.float32).astype(np.float64) * 1e-3
    ev = (rng.rand(Dout).astype(np.float32) + 0.1) * 1e-4
    want = O.pca_apply(X, mu, W, ev)
    args = [torch.from_numpy(a).cuda() for a in (X, W, mu, ev)]
    got = engine.pca_project(*args).cpu().numpy()
    gotn = engine.pca_project(*args, normalize_rows=True).cpu().numpy()
    if mode == "tc":
        _cmp_tc(got, want)
        _cmp_tc(gotn, O.normalize_feat(want))
        return
    np.testing.assert_allclose(got, want, rtol=1e-9, atol=1e-11)
    np.testing.assert_allclose(gotn, O.normalize_feat(want), rtol=1e-9, atol=1e-11)


@pytest.mark.parametrize("S,Din,Dout", [(300, 4096, 200), (129, 1000, 64), (5, 64, 16), (1, 49152, 1024), (40, 512, 33)])
def test_tensor_core_projection_shapes(S, Din, Dout):
    # ragged rows (S % 128), components not a multiple of 128 (zero-padded planes), D_in with a partial 64-channel stage
    # (1000) and a single stage (64), K split with a short last chunk, un-centred large mean
    rng = np.random.RandomState(S + Din)
    X = rng.randn(S, Din) * 0.05 + 0.3
    W = (rng.randn(Dout, Din) / np.sqrt(Din)).astype(np.float32)
    mu = X.mean(axis=0) if S > 1 else rng.randn(Din) * 0.05 + 0.3
    ev = (rng.rand(Dout).astype(np.float32) + 0.1)
    want = O.pca_apply(X, mu, W, ev)
    args = [torch.from_numpy(np.ascontiguousarray(a)).cuda() for a in (X, W, mu, ev)]
    _cmp_tc(engine.pca_project(*args).cpu().numpy(), want)


def test_apply_pca_transform_from_pkl_dropin(tmp_path):
    from sklearn.decomposition import PCA
    rng = np.random.RandomState(5)
    train = (rng.randn(200, 64) @ rng.randn(64, 64) * 0.1).astype(np.float32)
    pca = PCA(n_components=16, whiten=True, svd_solver="arpack").fit(train)
    path = tmp_path / "pca.pkl"
    pickle.dump(pca, open(path, "wb"))
    X = torch.from_numpy(rng.randn(9, 64) * 0.1)
    y = func_vpr.apply_pca_transform_from_pkl(X, str(path))
    assert not y.is_cuda and y.dtype == torch.float64
    np.testing.assert_allclose(y.numpy(), pca.transform(X.numpy()), rtol=2e-5, atol=5e-8)   # sklearn 1.9: fp32 bias


def _synthetic_pca(Din, Dout, seed, dev):
    g = torch.Generator(device=dev).manual_seed(seed)
    qm, _ = torch.linalg.qr(torch.randn(Din, Dout, generator=g, device=dev))
    comp = qm.T.contiguous().float()
    mean = (torch.randn(Din, generator=g, device=dev, dtype=torch.float64) * 2e-3)
    ev = (torch.rand(Dout, generator=g, device=dev) * 1e-4 + 1e-5).float()
    return comp, mean, ev


@pytest.mark.parametrize("D,K,H,W,counts,Dout,order", [(1536, 32, 480, 640, [150, 37], 1024, 3),     # published shape, 2 tiles
                                                       (128, 64, 140, 182, [9, 1, 130], 200, 2),     # empty clusters, Dout pad
                                                       (64, 32, 196, 266, [5], 32, 0)])
def test_fused_aggregation_projection_matches_unfused_and_oracle(monkeypatch, D, K, H, W, counts, Dout, order):
    """Row f1: the aggregation epilogue writes (descriptor - mean) as bf16 operand planes and the projection reads them by
    TMA -- no [S, K*D] fp64 matrix.  Against the unfused kernels (fp64 descriptors -> tensor-core projection) and the
    oracle (fp64 throughout)."""
    from revisit_anything_b200 import synth
    from revisit_anything_b200._lib import TOKENS_DN
    dev = torch.device("cuda")
    N = (H // 14) * (W // 14)
    cfg = {"desired_height": H, "desired_width": W}
    centers = synth.make_centers(K, D, 23)
    comp, mean, ev = _synthetic_pca(K * D, Dout, 5, dev)
    toks, bits, adjs, wants = [], [], [], []
    for i, S in enumerate(counts):
        t = synth.make_tokens(D, H // 14, W // 14, 640 + i, centers)
        m = synth.make_masks(S, H // 2, W // 2, 650 + i)
        adj = torch.from_numpy(O.neighbour_adjacency(m, order)) if order else None
        v = O.seg_vlad_single_img(t, m, centers, cfg, adj)[0].numpy()
        wants.append(O.pca_apply(v, mean.cpu().numpy(), comp.cpu().numpy(), ev.cpu().numpy()))
        toks.append(t.reshape(D, N)); adjs.append(adj)
        bits.append(engine.mask_to_membership(torch.from_numpy(np.asarray(m)).to(dev), H, W))
    tok = torch.stack(toks).to(dev)
    assert engine.pca_fusable(D, K, Dout)
    fused = engine.aggregate_project_pca(tok, N, D, TOKENS_DN, centers.to(dev), torch.cat(bits), counts,
                                         adjs if order else None, comp, mean, ev)
    desc = engine.aggregate_batch(tok, N, D, TOKENS_DN, centers.to(dev), torch.cat(bits), counts, adjs if order else None)
    unfused = engine.pca_project(desc, comp, mean, ev)
    _cmp_tc(fused.cpu().numpy(), np.concatenate(wants))
    _cmp_tc(fused.cpu().numpy(), unfused.cpu().numpy())
    # with the row normalisation of normalizeFeat
    fn = engine.aggregate_project_pca(tok, N, D, TOKENS_DN, centers.to(dev), torch.cat(bits), counts,
                                      adjs if order else None, comp, mean, ev, normalize_rows=True)
    _cmp_tc(fn.cpu().numpy(), O.normalize_feat(np.concatenate(wants)))


def test_fused_path_zero_residual_blocks():
    # tokens equal to their centres: populated clusters with zero block norm -> the row-norm correction must also work on
    # the planes output (rownorm_fixup_planes_kernel)
    from revisit_anything_b200._lib import TOKENS_ND, TOKENS_PRENORMALIZED
    dev = torch.device("cuda")
    D, K, N, S = 64, 32, 96, 4
    g = torch.Generator().manual_seed(0)
    centers = torch.nn.functional.normalize(torch.randn(K, D, generator=g), dim=1)
    x = torch.nn.functional.normalize(torch.randn(N, D, generator=g), dim=1)
    x[:20] = centers[3]
    x[20:30] = centers[7]
    member = torch.rand(S, N, generator=g) < 0.4
    member[0, :30] = True
    member[0, 30:] = False
    lab0, _ = O.assign_labels(x, centers)
    member[1] = (lab0 != 3) & (torch.rand(N, generator=g) < 0.5)
    member[1, :20] = True
    want, _, _ = O.vlad_single(x, centers, member, None)
    comp, mean, ev = _synthetic_pca(K * D, 48, 7, dev)
    bits = engine.pack_membership(member.to(dev))
    got = engine.aggregate_project_pca(x.t().contiguous().to(dev), N, D, 0 | TOKENS_PRENORMALIZED, centers.to(dev), bits, [S],
                                       None, comp, mean, ev)
    ref = O.pca_apply(want.numpy(), mean.cpu().numpy(), comp.cpu().numpy(), ev.cpu().numpy())
    _cmp_tc(got.cpu().numpy(), ref)
