"""CPU checks of the arithmetic assumptions the tensor-core kernels rest on (no GPU needed).

* csrc/aggregate_tc.cu, assign_tc.cu, project_tc.cu split every fp32 operand into three bf16 pieces (hi, mid, lo) with
  round-to-nearest and rely on hi + mid + lo == x EXACTLY (all 24 mantissa bits), so that products with a 0/1 mask are exact
  and six-term products are fp32-equivalent.
* csrc/knn.cu stores a row as fp16 of x * 2^(14 - e) (row maximum in [2^14, 2^15)): no overflow for any fp32 input and a
  relative rounding error of at most 2^-11 per element above the flushed sub-normal range.
"""
import numpy as np
import torch
from hypothesis import given, settings, strategies as st

from revisit_anything_b200 import distributed as D


def _split3(x: torch.Tensor):
    hi = x.to(torch.bfloat16)
    r1 = x - hi.float()
    mid = r1.to(torch.bfloat16)
    r2 = r1 - mid.float()
    lo = r2.to(torch.bfloat16)
    return hi, mid, lo, r2 - lo.float()


def test_bf16_three_way_split_is_exact():
    g = torch.Generator().manual_seed(0)
    parts = [torch.randn(200_000, generator=g),                                   # unit scale (normalised tokens, residuals)
             torch.randn(200_000, generator=g) * 1e-3,
             torch.randn(200_000, generator=g) * 3e4,                             # un-normalised backbone features
             torch.rand(100_000, generator=g) * 2 - 1,
             (torch.randint(-(2 ** 24), 2 ** 24, (100_000,), generator=g).float() * 2.0 ** -24),   # every mantissa pattern
             torch.tensor([0.0, -0.0, 1.0, -1.0, 2.0 ** -100, 1.9999999, 3.3e38, -3.3e38, 1.17549435e-38])]   # (|x| above bf16's
             # largest finite value 3.39e38 would round the first piece to inf: outside any descriptor's range)
    x = torch.cat(parts)
    hi, mid, lo, rest = _split3(x)
    assert torch.equal(rest, torch.zeros_like(rest)), "residual after three bf16 pieces must be exactly zero"
    assert torch.equal(hi.float() + mid.float() + lo.float(), x)                  # (fp32 adds of the pieces are exact too)
    # magnitudes: the second and third piece are at most 2^-8 / 2^-16 of the first (what the accumulator split relies on)
    nz = hi.float() != 0
    assert (mid.float().abs()[nz] <= hi.float().abs()[nz] * 2.0 ** -8).all()
    assert (lo.float().abs()[nz] <= hi.float().abs()[nz] * 2.0 ** -16).all()


def test_mask_times_split_planes_reproduces_fp32_sums():
    # V = M (0/1) @ R evaluated as M @ hi + M @ mid + M @ lo in fp64 equals M @ R in fp64: the split loses nothing
    g = torch.Generator().manual_seed(1)
    R = torch.randn(96, 64, generator=g) * 0.1
    M = (torch.rand(37, 96, generator=g) < 0.4).double()
    hi, mid, lo, _ = _split3(R)
    want = M @ R.double()
    got = M @ lo.double() + M @ mid.double() + M @ hi.double()
    assert torch.equal(got, want)


def test_fp16_row_scaling_never_overflows_and_bounds_the_rounding():
    g = torch.Generator().manual_seed(2)
    rows = [torch.randn(64, 1536, generator=g),
            torch.randn(64, 1536, generator=g) * 1e30, torch.randn(64, 1536, generator=g) * 1e-30,
            torch.randn(64, 1536, generator=g) * torch.logspace(-6, 6, 1536)]
    for x in rows:
        mx = x.abs().amax(dim=1, keepdim=True)
        e = torch.floor(torch.log2(mx))
        sc = torch.exp2(14 - e)
        y = x * sc                                               # row maximum in [2^14, 2^15)
        assert float(y.abs().max()) < 65504.0
        h = y.to(torch.float16)
        assert torch.isfinite(h.float()).all()
        normal = y.abs() >= 6.103515625e-05                      # above the flushed fp16 sub-normal range
        rel = ((h.float() - y).abs() / y.abs().clamp_min(1e-30))[normal]
        assert float(rel.max()) <= 2.0 ** -11 * 1.0001


@settings(max_examples=60, deadline=None)
@given(st.integers(0, 5000), st.integers(1, 16))
def test_shard_bounds_partition_the_rows(n_rows, world):
    b = D.shard_bounds(n_rows, world)
    assert len(b) == world and b[0][0] == 0 and b[-1][1] == n_rows
    sizes = [hi - lo for lo, hi in b]
    assert all(lo2 == hi1 for (_, hi1), (lo2, _) in zip(b[:-1], b[1:]))           # contiguous, in order
    assert max(sizes) - min(sizes) <= 1 and sizes == sorted(sizes, reverse=True)    # nearly equal, larger shards first


@settings(max_examples=40, deadline=None)
@given(st.lists(st.tuples(st.floats(0, 4, width=32), st.integers(-1, 2 ** 31 - 1)), min_size=1, max_size=64))
def test_pack_topk_roundtrip_is_bit_exact(pairs):
    d2 = torch.tensor([[p[0] for p in pairs]], dtype=torch.float32)
    idx = torch.tensor([[p[1] for p in pairs]], dtype=torch.int64)
    a, b = D.unpack_topk(D.pack_topk(d2, idx))
    assert torch.equal(a.view(torch.int32), d2.view(torch.int32)) and torch.equal(b, idx)
