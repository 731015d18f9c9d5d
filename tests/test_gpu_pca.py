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


def test_golden_reference_vector(golden_dir):
    g = np.load(os.path.join(golden_dir, "pca_apply.npz"))
    y = engine.pca_project(torch.from_numpy(g["X"]).cuda(), torch.from_numpy(g["components"]).cuda(),
                           torch.from_numpy(g["mean"]).cuda(), torch.from_numpy(g["explained_variance"]).cuda())
    # golden = reference function with sklearn 1.9 (fp32 bias term, ~1e-8 abs); kernel = sklearn-1.3.2 formula in fp64
    np.testing.assert_allclose(y.cpu().numpy(), g["Y"], rtol=2e-5, atol=5e-8)
    want = O.pca_apply(g["X"], g["mean"], g["components"], g["explained_variance"])
    np.testing.assert_allclose(y.cpu().numpy(), want, rtol=1e-10, atol=1e-12)         # fp64 on both sides


def test_real_shape_vs_oracle_and_normalize():
    rng = np.random.RandomState(3)
    S, Din, Dout = 37, 49152, 1024
    X = rng.randn(S, Din) / np.sqrt(Din)
    W = (rng.randn(Dout, Din) / np.sqrt(Din)).astype(np.float32)
    mu = rng.randn(Din).astype(np.float32).astype(np.float64) * 1e-3
    ev = (rng.rand(Dout).astype(np.float32) + 0.1) * 1e-4
    want = O.pca_apply(X, mu, W, ev)
    args = [torch.from_numpy(a).cuda() for a in (X, W, mu, ev)]
    got = engine.pca_project(*args).cpu().numpy()
    np.testing.assert_allclose(got, want, rtol=1e-9, atol=1e-11)
    gotn = engine.pca_project(*args, normalize_rows=True).cpu().numpy()
    np.testing.assert_allclose(gotn, O.normalize_feat(want), rtol=1e-9, atol=1e-11)


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
