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


def test_merge_topk_with_shards_shorter_than_k():
    # a shard with fewer rows than k returns (+inf, -1) padded lists (like faiss); the merge must skip the padding, and a
    # bank with fewer than k rows in total stays padded after the merge
    q, r = synth.make_descriptor_bank(130, 700, 128, seed=9, planted=20, device=DEV)
    k = 200
    qb = engine.Bank.prepare(q)
    for bounds in ([0, 60, 150, 700], [0, 50, 120, 180]):
        n = bounds[-1]
        d2_full, idx_full = _run_tc(q, r[:n], k)
        parts_d, parts_i = [], []
        for g in range(3):
            d, i = engine.knn(qb, engine.Bank.prepare(r[bounds[g]:bounds[g + 1]]), k, row_offset=bounds[g])
            parts_d.append(d)
            parts_i.append(i)
        md, mi = engine.merge_topk(torch.stack(parts_d), torch.stack(parts_i))
        np.testing.assert_array_equal(md.cpu().numpy(), d2_full)
        np.testing.assert_array_equal(mi.cpu().numpy(), idx_full)
        if n < k:
            assert (mi.cpu().numpy()[:, n:] == -1).all() and np.isinf(md.cpu().numpy()[:, n:]).all()


def test_independent_third_party_brute_force_agrees():
    # a6 is unpinned against faiss (absent offline); the oracle's flat_l2_search and the kernels come from the same hand.
    # Two third-party exact searches break that common mode: scikit-learn's NearestNeighbors(algorithm="brute") and
    # scipy's cdist, both on the fp32 inputs in fp64.
    from scipy.spatial.distance import cdist
    from sklearn.neighbors import NearestNeighbors
    q, r = synth.make_descriptor_bank(200, 6000, 128, seed=21, planted=40, device=DEV)
    k = 50
    d2, idx = _run_tc(q, r, k)
    qn, rn = q.cpu().numpy().astype(np.float64), r.cpu().numpy().astype(np.float64)
    nn = NearestNeighbors(n_neighbors=k + 8, algorithm="brute", metric="euclidean").fit(rn)
    dist, ind = nn.kneighbors(qn)
    assert_knn_close(d2, idx, dist ** 2, ind.astype(np.int64), k_check=k)
    dm = cdist(qn, rn, metric="sqeuclidean")
    order = np.argsort(dm, axis=1, kind="stable")[:, :k + 8]
    assert_knn_close(d2, idx, np.take_along_axis(dm, order, 1), order.astype(np.int64), k_check=k)
    # and the oracle itself against the same third parties (CPU twin of this check: tests/test_oracle_golden.py)
    D2o, Io = O.flat_l2_search(q.cpu().numpy(), r.cpu().numpy(), k)
    assert_knn_close(D2o, Io, np.take_along_axis(dm, order, 1), order.astype(np.int64), k_check=k)


def test_view_bank_equals_copy_bank():
    # Bank.prepare references the caller's fp32 rows by default (segvlad_bank_prepare_view); copy=True owns them: same lists,
    # also through the row-offset / sliced-shard path (a row slice of a contiguous matrix is itself contiguous and aligned)
    q, r = synth.make_descriptor_bank(500, 30000, 512, seed=31, planted=60, device=DEV)
    qv, rv = engine.Bank.prepare(q), engine.Bank.prepare(r)
    qc, rc = engine.Bank.prepare(q, copy=True), engine.Bank.prepare(r, copy=True)
    assert rv.src is not None and rc.src is None
    a = engine.knn(qv, rv, 200)
    b = engine.knn(qc, rc, 200)
    c = engine.knn(qv, rc, 200)
    torch.cuda.synchronize()
    assert torch.equal(a[0], b[0]) and torch.equal(a[1], b[1]) and torch.equal(a[0], c[0]) and torch.equal(a[1], c[1])
    s = engine.knn(qv, engine.Bank.prepare(r[10000:]), 200, row_offset=10000)
    d64, i64 = _ref(q, r[10000:], 208)
    assert_knn_close(s[0].cpu().numpy(), s[1].cpu().numpy() - 10000, d64, i64, k_check=200)


def test_async_search_packed_output_and_overflow_flag():
    # segvlad_knn_async: no host synchronisation, packed (d2 bits << 32 | int32 global row) lists for the all-gather, the
    # overflow flag published on the device; identical lists to the synchronous call
    from revisit_anything_b200 import distributed as Dm
    q, r = synth.make_descriptor_bank(300, 21000, 192, seed=12, planted=50, device=DEV)
    qb, rb = engine.Bank.prepare(q), engine.Bank.prepare(r)
    k, off = 200, 1_000_000
    d2s, idxs = engine.knn(qb, rb, k, row_offset=off)
    payload = torch.full((300 * k + Dm.TRAILER,), -7, dtype=torch.int64, device=DEV)
    d2a, idxa, flag = engine.knn_async(qb, rb, k, off, schedule=0, packed_out=payload)
    torch.cuda.synchronize()
    assert int(flag.item()) == 0
    assert torch.equal(d2a, d2s) and torch.equal(idxa, idxs)
    assert torch.equal(payload[:300 * k].view(300, k).cpu(), Dm.pack_topk(d2s.cpu(), idxs.cpu()))
    assert bool((payload[300 * k:] == -7).all())                 # nothing written behind the lists
    for sched in (0, 1):                                         # packed-only mode, both schedules
        p2 = torch.empty(300 * k, dtype=torch.int64, device=DEV)
        a, b, f2 = engine.knn_async(qb, rb, k, off, schedule=sched, packed_out=p2, unpacked=False)
        torch.cuda.synchronize()
        assert a is None and b is None and int(f2.item()) == 0 and torch.equal(p2, payload[:300 * k])
    with pytest.raises(ValueError):                              # int32 bound of the packed rows is a host-side check
        engine.knn_async(qb, rb, k, 2 ** 31 - 100, packed_out=payload)
    # adversarial order (test above): the fast schedule must raise the flag, the conservative one must not and be exact
    g = torch.Generator(device=DEV).manual_seed(0)
    D, Nr, Nq = 64, 30000, 130
    base = torch.nn.functional.normalize(torch.randn(1, D, generator=g, device=DEV), dim=1)
    noise = torch.nn.functional.normalize(torch.randn(Nr, D, generator=g, device=DEV), dim=1)
    a = torch.linspace(0.05, 0.95, Nr, device=DEV)[:, None]
    r = torch.nn.functional.normalize(a * base + (1 - a) * 0.3 * noise, dim=1)
    q = torch.nn.functional.normalize(base + 0.01 * torch.randn(Nq, D, generator=g, device=DEV), dim=1)
    qb, rb = engine.Bank.prepare(q), engine.Bank.prepare(r)
    _, _, f0 = engine.knn_async(qb, rb, 200, 0, schedule=0)
    d2c, idxc, f1 = engine.knn_async(qb, rb, 200, 0, schedule=1)
    torch.cuda.synchronize()
    assert int(f0.item()) == 1 and int(f1.item()) == 0
    d2s, idxs = engine.knn(qb, rb, 200)
    assert torch.equal(d2c, d2s) and torch.equal(idxc, idxs)
    # the host-side pipeline repeats the step by itself when the flag is set
    ops = Dm.EngineOps()
    qoff = torch.tensor([0, 60, 130], dtype=torch.int32, device=DEV)
    rimg = (torch.arange(Nr, device=DEV) // 100).to(torch.int32)
    d2p, idxp, preds = Dm.sharded_search_and_vote(ops, qb, rb, 0, qoff, rimg, 300, 200, 50, 5)
    assert torch.equal(d2p, d2s) and torch.equal(idxp, idxs)
    assert torch.equal(preds, ops.vote(idxs, d2s, qoff, rimg, 300, 5, 50))


def test_merge_packed_is_the_same_k_way_merge():
    # packed gather buffer (with trailer words) -> rank-by-binary-search merge; against the unpacked sort-based merge and
    # the single-shard search, incl. shards shorter than k (padding) and an exact cross-shard tie
    from revisit_anything_b200 import distributed as Dm
    q, r = synth.make_descriptor_bank(130, 700, 128, seed=9, planted=20, device=DEV)
    r[10] = r[400]
    k = 200
    qb = engine.Bank.prepare(q)
    for bounds in ([0, 60, 150, 700], [0, 50, 120, 180], [0, 350, 700], [0, 90, 170, 260, 350, 430, 520, 610, 700],
                   [0, 10, 20, 30, 40, 50, 60, 70, 700]):        # 8 shards: balanced, and one shard holding almost everything
        n, G = bounds[-1], len(bounds) - 1
        d2_full, idx_full = engine.knn(qb, engine.Bank.prepare(r[:n]), k)
        buf = torch.zeros((G, 130 * k + Dm.TRAILER), dtype=torch.int64, device=DEV)
        pd, pi = [], []
        for g in range(G):
            d, i, _ = engine.knn_async(qb, engine.Bank.prepare(r[bounds[g]:bounds[g + 1]]), k, bounds[g], packed_out=buf[g])
            pd.append(d)
            pi.append(i)
        md, mi, mp = engine.merge_topk_packed(buf, 130, k, want_packed=True)
        md2, mi2 = engine.merge_topk(torch.stack(pd), torch.stack(pi))
        torch.cuda.synchronize()
        assert torch.equal(md, d2_full) and torch.equal(mi, idx_full)
        assert torch.equal(md, md2) and torch.equal(mi, mi2)
        assert torch.equal(mp.cpu(), Dm.pack_topk(md.cpu(), mi.cpu()))


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


def _exact_d2_fp64(q, r):
    q64, r64 = q.double(), r.double()
    d = (q64 * q64).sum(1)[:, None] + (r64 * r64).sum(1)[None, :] - 2.0 * (q64 @ r64.T)
    return d.clamp_min(0.0)


@pytest.mark.parametrize("D", [64, 520, 1536, 4096])
@pytest.mark.parametrize("kind", ["unit", "huge", "tiny", "mixed_rows", "heavy_tail", "near_duplicates"])
def test_error_model_bounds_the_fp16_pass(D, kind):
    """The filter is exact only if |approximate d2 - fp32 d2| <= E_row for EVERY pair (knn.cu header): check the bound
    the library assumes against fp64 on all pairs of a 512 x 4096 problem, for inputs that stress the fp16 plane."""
    g = torch.Generator(device=DEV).manual_seed(D + len(kind))
    Nq, Nr = 512, 4096
    q = torch.randn(Nq, D, generator=g, device=DEV)
    r = torch.randn(Nr, D, generator=g, device=DEV)
    if kind == "unit":
        q, r = torch.nn.functional.normalize(q, dim=1), torch.nn.functional.normalize(r, dim=1)
    elif kind == "huge":
        q, r = q * 3.0e5, r * 7.0e5                      # far outside the fp16 range without the row scaling
    elif kind == "tiny":
        q, r = q * 1.0e-9, r * 3.0e-10
    elif kind == "mixed_rows":
        q = q * torch.logspace(-4, 4, Nq, device=DEV)[:, None]
        r = r * torch.logspace(-3, 3, Nr, device=DEV)[torch.randperm(Nr, generator=g, device=DEV)][:, None]
    elif kind == "heavy_tail":                            # a few huge channels: most of the row falls into the low fp16 bits
        q = q * torch.exp(3.0 * torch.randn(Nq, D, generator=g, device=DEV))
        r = r * torch.exp(3.0 * torch.randn(Nr, D, generator=g, device=DEV))
    else:                                                 # d2 ~ 0 by cancellation
        base = torch.nn.functional.normalize(torch.randn(8, D, generator=g, device=DEV), dim=1)
        q = base[torch.arange(Nq, device=DEV) % 8] + 1e-3 * q / D ** 0.5
        r = base[torch.arange(Nr, device=DEV) % 8] + 1e-3 * r / D ** 0.5
    q, r = q.contiguous(), r.contiguous()
    approx, bound = engine.knn_debug_approx(engine.Bank.prepare(q), engine.Bank.prepare(r))
    torch.cuda.synchronize()
    exact = _exact_d2_fp64(q, r)
    err = (approx.double() - exact).abs()
    ratio = (err / bound.double()[:, None]).max().item()
    assert torch.isfinite(bound).all() and (bound > 0).all()
    assert ratio <= 1.0, f"approximate score off by {ratio:.3f} x the assumed bound"
    # the bound must also be useful: for unit-norm rows a small fraction of the d2 spread (~ 2 / sqrt(D))
    if kind == "unit":
        assert bound.max().item() < 0.05 * 2.0 / D ** 0.5 + 1e-3


@pytest.mark.parametrize("runner", [_run_simt, _run_tc], ids=["simt", "tcgen05"])
def test_massive_duplicates_select_by_index(runner):
    # 6000 identical rows: every candidate buffer overflows under the optimistic schedule and the approximate scores
    # cannot separate them; the conservative schedule must return the k smallest INDICES of the tie group
    q, r = synth.make_descriptor_bank(200, 20000, 128, seed=21, planted=0, device=DEV)
    dup = torch.randperm(20000, generator=torch.Generator().manual_seed(1))[:6000].to(DEV)
    r[dup] = q[7]
    d2, idx = runner(q, r, 200)
    want = np.sort(dup.cpu().numpy())[:200]
    np.testing.assert_array_equal(idx[7], want)
    assert (d2[7] <= 2e-6).all()
    d64, i64 = _ref(q, r, 208)
    rows = [i for i in range(200) if i != 7]
    np.testing.assert_allclose(d2[rows], d64[rows, :200], rtol=1e-5, atol=2e-6)


def test_unnormalised_rows_of_mixed_magnitude():
    # faiss.IndexFlatL2 takes arbitrary fp32 rows: per-row power-of-two scaling keeps the fp16 plane in range
    g = torch.Generator(device=DEV).manual_seed(33)
    q = torch.randn(300, 256, generator=g, device=DEV) * torch.logspace(-2, 2, 300, device=DEV)[:, None]
    r = torch.randn(20000, 256, generator=g, device=DEV) * torch.logspace(-2, 2, 20000, device=DEV)[
        torch.randperm(20000, generator=g, device=DEV)][:, None]
    d2, idx = _run_tc(q.contiguous(), r.contiguous(), 100)
    d2s, idxs = _run_simt(q.contiguous(), r.contiguous(), 100)
    d64, i64 = _ref(q, r, 100)
    np.testing.assert_allclose(d2, d64, rtol=2e-5)
    assert (idx == i64).mean() > 0.999 and (idx == idxs).mean() > 0.999
