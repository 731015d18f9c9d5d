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
