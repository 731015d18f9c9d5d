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
