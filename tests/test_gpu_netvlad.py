"""GPU parity of the NetVLAD anti-burst aggregation (config 5) vs the oracle restatement of
VLAD-BuFF/models/aggregators/aggregation.py:266-361 (fp32)."""
import numpy as np
import pytest
import torch

from gpu_util import assert_desc_close
from oracle import segvlad_oracle as O
from revisit_anything_b200 import engine

pytestmark = pytest.mark.gpu


def _params(K, D, seed):
    g = torch.Generator().manual_seed(seed)
    cent = torch.rand(K, D, generator=g)                        # aggregation.py:216 init
    cn = cent / cent.norm(dim=1, keepdim=True)
    alpha = 12.0                                                # init_params-style scaling (aggregation.py:245)
    return cent, alpha * cn


@pytest.mark.parametrize("B,D,H,K,ab", [(3, 768, 23, 64, (8.0, 7.0, 1.0)), (2, 768, 23, 128, (8.0, 7.0, 1.0)),
                                        (2, 96, 9, 32, (5.0, 3.0, 0.5)), (1, 64, 5, 16, (8.0, 7.0, 1.0))])
def test_netvlad_antiburst_vs_oracle(B, D, H, K, ab):
    g = torch.Generator().manual_seed(B * 1000 + D + K)
    x = torch.randn(B, D, H, H, generator=g)
    x[:, :, 0, :3] = x[:, :, 0, :1]                             # a burst: repeated tokens get down-weighted
    cent, W = _params(K, D, K)
    want = O.netvlad_antiburst(x.reshape(B, D, -1), cent, W, ab)
    got = engine.netvlad_antiburst(x.cuda(), cent.cuda(), W.cuda(), ab)
    assert got.shape == (B, K * D)
    np.testing.assert_allclose(got.norm(dim=1).cpu().numpy(), 1.0, atol=1e-5)
    assert_desc_close(got.cpu().numpy(), want.numpy(), rtol=2e-5)   # fp32 on both sides (reference: fp32 / fp16 autocast)


def test_tensor_core_path_matches_ffma_path_and_is_deterministic(monkeypatch):
    # tcgen05 kernels (fp16 hi/lo split operands, csrc/netvlad_tc.cu) against the fp32 FFMA kernels (SEGVLAD_NETVLAD_TC=0), on a
    # bursty image (a quarter of the tokens are copies of one token) and a K that is not 128; two runs are bit-identical
    g = torch.Generator().manual_seed(77)
    B, D, H, K = 3, 768, 23, 64
    x = torch.randn(B, D, H, H, generator=g)
    x[1, :, :6, :] = x[1, :, :1, :1]
    cent, W = _params(K, D, 5)
    xc, cc, wc = x.cuda(), cent.cuda(), W.cuda()
    a = engine.netvlad_antiburst(xc, cc, wc)
    b = engine.netvlad_antiburst(xc, cc, wc)
    assert torch.equal(a, b)
    monkeypatch.setenv("SEGVLAD_NETVLAD_TC", "0")
    c = engine.netvlad_antiburst(xc, cc, wc)
    assert_desc_close(a.cpu().numpy(), c.cpu().numpy(), rtol=2e-5)
    want = O.netvlad_antiburst(x.reshape(B, D, -1), cent, W)
    assert_desc_close(a.cpu().numpy(), want.numpy(), rtol=2e-5)


def test_unsupported_cluster_count_uses_the_ffma_kernels():
    g = torch.Generator().manual_seed(3)
    x = torch.randn(2, 64, 6, 6, generator=g)
    cent, W = _params(24, 64, 9)                                # K = 24: not a multiple of 16
    got = engine.netvlad_antiburst(x.cuda(), cent.cuda(), W.cuda())
    assert_desc_close(got.cpu().numpy(), O.netvlad_antiburst(x.reshape(2, 64, -1), cent, W).numpy(), rtol=2e-5)
