"""GPU parity of the matching path: tcgen05 kNN and the SIMT cross-check vs the oracle's restated
faiss IndexFlatL2 search.  Tolerance 1e-5 relative on distances / cosine scores (north_star); indices
bit-exact outside fp32 near-ties."""
import numpy as np
import pytest
import torch

from gpu_util import assert_knn_close
from oracle import segvlad_oracle as O
from revisit_anything_b200 import engine, synth

pytestmark = pytest.mark.gpu
DEV = "cuda"


def _ref(q, r, k):
    return O.flat_l2_search_fp64(q.cpu().numpy(), r.cpu().numpy(), k)


def _run_tc(q, r, k, off=0):
    d2, idx = engine.knn(engine.Bank.prepare(q), engine.Bank.prepare(r), k, row_offset=off)
    torch.cuda.synchronize()
    return d2.cpu().numpy(), idx.cpu().numpy()


def _run_simt(q, r, k, off=0):
    d2, idx = engine.knn_simt(q, r, k, row_offset=off)
    torch.cuda.synchronize()
    return d2.cpu().numpy(), idx.cpu().numpy()


@pytest.mark.parametrize("runner", [_run_simt, _run_tc], ids=["simt", "tcgen05"])
@pytest.mark.parametrize("Nq,Nr,D,k", [(300, 5000, 96, 200), (257, 4097, 64, 50), (129, 9000, 100, 200)])
def test_small_vs_oracle(runner, Nq, Nr, D, k):
    q, r = synth.make_descriptor_bank(Nq, Nr, D, seed=Nq + D, planted=40, device=DEV)
    d2, idx = runner(q, r, k)
    d64, i64 = _ref(q, r, k + 8)
    assert (np.diff(d2, axis=1) >= 0).all()
    frac = assert_knn_close(d2, idx, d64, i64, k_check=k)
    assert frac > 0.9
    # also against the oracle's fp32 restatement of faiss (what the reference would print)
    D2o, Io = O.flat_l2_search(q.cpu().numpy(), r.cpu().numpy(), k)
    np.testing.assert_allclose(d2, D2o, rtol=1e-5, atol=2e-6)
    assert (idx == Io).mean() > 0.995


def test_tcgen05_multi_round_1536d():
    # D = 1536 (config-2 descriptor size), several filter rounds (4096 -> ~31k -> rest), ragged edges
    q, r = synth.make_descriptor_bank(1000, 60001, 1536, seed=2, planted=300, device=DEV)
    d2, idx = _run_tc(q, r, 200, off=1_000_000)
    d2s, idxs = _run_simt(q, r, 200, off=1_000_000)
    d64, i64 = _ref(q, r, 208)
    assert_knn_close(d2, idx - 1_000_000, d64, i64, k_check=200)
    assert_knn_close(d2s, idxs - 1_000_000, d64, i64, k_check=200)
    np.testing.assert_allclose(d2, d2s, rtol=1e-5, atol=2e-6)
    # cosine scores (sims = 2 - d2) within 1e-5 relative of the fp64 value
    np.testing.assert_allclose(2 - d2, 2 - d64[:, :200], rtol=1e-5, atol=1e-6)


@pytest.mark.parametrize("D", [512, 1024])
def test_tcgen05_other_dims(D):
    q, r = synth.make_descriptor_bank(640, 12345, D, seed=D, planted=100, device=DEV)
    d2, idx = _run_tc(q, r, 200)
    d64, i64 = _ref(q, r, 208)
    assert_knn_close(d2, idx, d64, i64, k_check=200)


@pytest.mark.parametrize("runner", [_run_simt, _run_tc], ids=["simt", "tcgen05"])
def test_fewer_refs_than_k_and_exact_duplicates(runner):
    q, r = synth.make_descriptor_bank(70, 120, 128, seed=9, planted=10, device=DEV)
    r[5] = r[17]
    r[99] = r[17]                                 # exactly equal rows -> exactly equal distances
    q[3] = r[17]                                  # distance ~0 -> clamp at 0
    d2, idx = runner(q, r, 200)
    assert (idx[:, 120:] == -1).all() and np.isinf(d2[:, 120:]).all()
    assert (d2[:, :120] >= 0).all()
    d64, i64 = _ref(q, r, 120)
    np.testing.assert_allclose(d2[:, :120], d64, rtol=1e-5, atol=2e-6)
    for row in range(70):                         # ties come out in ascending index order
        pos = [int(np.where(idx[row] == j)[0][0]) for j in (5, 17, 99)]
        assert pos == sorted(pos)
    assert set(idx[3, :3].tolist()) == {5, 17, 99}


def test_adversarial_order_triggers_safe_schedule():
    # every later reference is closer than all earlier ones for every query: the optimistic chunk schedule
    # overflows the candidate buffers and the library must fall back to the conservative schedule
    g = torch.Generator(device=DEV).manual_seed(0)
    D, Nr, Nq = 64, 30000, 130
    base = torch.nn.functional.normalize(torch.randn(1, D, generator=g, device=DEV), dim=1)
    noise = torch.nn.functional.normalize(torch.randn(Nr, D, generator=g, device=DEV), dim=1)
    a = torch.linspace(0.05, 0.95, Nr, device=DEV)[:, None]
    r = torch.nn.functional.normalize(a * base + (1 - a) * 0.3 * noise, dim=1)
    q = torch.nn.functional.normalize(base + 0.01 * torch.randn(Nq, D, generator=g, device=DEV), dim=1)
    for runner in (_run_simt, _run_tc):
        d2, idx = runner(q, r, 200)
        d64, i64 = _ref(q, r, 208)
        assert_knn_close(d2, idx, d64, i64, k_check=200)
        assert np.median(idx) > Nr - 2000            # the nearest refs are the last ones scanned


def test_merge_topk_matches_single_shard():
    q, r = synth.make_descriptor_bank(300, 9000, 256, seed=4, planted=50, device=DEV)
    k = 200
    d2_full, idx_full = _run_tc(q, r, k)
    qb = engine.Bank.prepare(q)
    parts_d, parts_i = [], []
    bounds = [0, 2100, 4500, 9000]
    for g in range(3):
        d, i = engine.knn(qb, engine.Bank.prepare(r[bounds[g]:bounds[g + 1]]), k, row_offset=bounds[g])
        parts_d.append(d)
        parts_i.append(i)
    md, mi = engine.merge_topk(torch.stack(parts_d), torch.stack(parts_i))
    np.testing.assert_array_equal(md.cpu().numpy(), d2_full)
    np.testing.assert_array_equal(mi.cpu().numpy(), idx_full)


def test_bank_prepare_f64_normalizes_like_normalizeFeat():
    g = torch.Generator().manual_seed(3)
    x = torch.randn(500, 200, generator=g, dtype=torch.float64) * 3.0
    y = torch.randn(40, 200, generator=g, dtype=torch.float64) * 0.2
    rb = engine.Bank.prepare_f64(x.to(DEV), normalize_rows=True)
    qb = engine.Bank.prepare_f64(y.to(DEV), normalize_rows=True)
    d2, idx = engine.knn(qb, rb, 50)
    xn, yn = O.normalize_feat(x.numpy()), O.normalize_feat(y.numpy())
    d64, i64 = O.flat_l2_search_fp64(yn, xn, 58)
    assert_knn_close(d2.cpu().numpy(), idx.cpu().numpy(), d64, i64, k_check=50)


def test_knn_from_host_streamed_matches_resident():
    # pinned host inputs, reference bank streamed in sub-chunks while being scanned; result == resident search
    q, r = synth.make_descriptor_bank(700, 50000, 256, seed=12, planted=100, device=DEV)
    d2_res, idx_res = _run_tc(q, r, 200, off=77)
    qh, rh = q.cpu().pin_memory(), r.cpu().pin_memory()
    d2, idx, qb, rb = engine.knn_from_host(qh, rh, 200, row_offset=77)
    torch.cuda.synchronize()
    np.testing.assert_array_equal(d2.cpu().numpy(), d2_res)
    np.testing.assert_array_equal(idx.cpu().numpy(), idx_res)
    d2b, idxb = engine.knn(qb, rb, 200, row_offset=77)          # the banks are resident and re-usable afterwards
    np.testing.assert_array_equal(d2b.cpu().numpy(), d2_res)
    # pageable (non-pinned) host memory also works (no overlap, same answer)
    d2p, idxp, _, _ = engine.knn_from_host(q.cpu(), r.cpu(), 200, row_offset=77)
    np.testing.assert_array_equal(idxp.cpu().numpy(), idx_res)
