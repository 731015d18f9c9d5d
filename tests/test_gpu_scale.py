"""Full-size properties of the matching path (sizes the CPU oracle cannot finish in seconds): multi-block query sets,
large banks, size-independent invariants + brute-force fp64 check on a sample of rows (torch on the GPU)."""
import numpy as np
import pytest
import torch

from gpu_util import assert_knn_close
from revisit_anything_b200 import engine, synth

pytestmark = pytest.mark.gpu
DEV = "cuda"


def _brute_fp64(q, r, k):
    d = (q.double() ** 2).sum(1)[:, None] + (r.double() ** 2).sum(1)[None, :] - 2.0 * (q.double() @ r.double().T)
    d.clamp_(min=0)
    v, i = torch.sort(d, dim=1, stable=True)
    return v[:, :k].cpu().numpy(), i[:, :k].cpu().numpy()


def test_config4_slice_two_query_blocks_and_shards():
    # 20k queries (2 query blocks of the workspace) x 300k refs x 512-D (config-4 descriptor size), k = 200
    Nq, Nr, D, k = 20000, 300000, 512, 200
    q, r = synth.make_descriptor_bank(Nq, Nr, D, seed=44, planted=2000, device=DEV)
    qb, rb = engine.Bank.prepare(q), engine.Bank.prepare(r)
    d2, idx = engine.knn(qb, rb, k)
    torch.cuda.synchronize()
    assert bool((d2[:, 1:] >= d2[:, :-1]).all()), "distances must be ascending"
    assert int(idx.min()) >= 0 and int(idx.max()) < Nr
    srt = torch.sort(idx, dim=1).values
    assert bool((srt[:, 1:] != srt[:, :-1]).all()), "a reference row may appear once per query"
    # returned distances are the exact fp32 distances of the returned rows (sample of rows, fp64 recomputation)
    rows = torch.randperm(Nq, device=DEV)[:256]
    diff = q[rows].double()[:, None, :] - r[idx[rows]].double()
    np.testing.assert_allclose(d2[rows].cpu().numpy(), (diff * diff).sum(-1).cpu().numpy(), rtol=1e-5, atol=2e-6)
    # completeness on a sample: brute force over the whole bank in fp64
    sample = torch.cat([rows[:48], torch.tensor([0, 16383, 16384, Nq - 1], device=DEV)])   # incl. block boundaries
    d64, i64 = _brute_fp64(q[sample], r, k + 8)
    assert_knn_close(d2[sample].cpu().numpy(), idx[sample].cpu().numpy(), d64, i64, k_check=k)
    # row-sharded search + merge == single search (bit-identical)
    bounds = [0, 90000, 200000, Nr]
    pd, pi = [], []
    for g in range(3):
        a, b = engine.knn(qb, engine.Bank.prepare(r[bounds[g]:bounds[g + 1]]), k, row_offset=bounds[g])
        pd.append(a)
        pi.append(b)
    md, mi = engine.merge_topk(torch.stack(pd), torch.stack(pi))
    assert torch.equal(md, d2) and torch.equal(mi, idx)


def test_single_cta_variant_agrees(monkeypatch):
    # cta_group::1 and cta_group::2 kernels feed the same exact re-score: identical outputs
    q, r = synth.make_descriptor_bank(3000, 50000, 1536, seed=45, planted=300, device=DEV)
    qb, rb = engine.Bank.prepare(q), engine.Bank.prepare(r)
    d2a, ia = engine.knn(qb, rb, 200)
    monkeypatch.setenv("SEGVLAD_KNN_CTAS", "1")
    d2b, ib = engine.knn(qb, rb, 200)
    torch.cuda.synchronize()
    assert torch.equal(d2a, d2b) and torch.equal(ia, ib)


def test_rescore_by_reference_row_is_bit_identical_to_the_gather(monkeypatch):
    # final re-score through the inverted candidate lists (one warp per reference row, bank streamed once) against the
    # per-query gather kernel: same per-lane summation order -> identical (d2, idx), incl. planted hubs (a reference row
    # that is a candidate of MANY queries) and a D that is not a multiple of 4 (scalar path)
    for Nq, Nr, D in [(3000, 50000, 1536), (700, 20000, 510)]:
        q, r = synth.make_descriptor_bank(Nq, Nr, D, seed=46, planted=300, device=DEV)
        r[123] = torch.nn.functional.normalize(q[:50].mean(0), dim=0)      # hub: near the centroid of 50 queries
        qb, rb = engine.Bank.prepare(q), engine.Bank.prepare(r)
        d2a, ia = engine.knn(qb, rb, 200)
        monkeypatch.setenv("SEGVLAD_KNN_RESCORE_REF", "0")
        d2b, ib = engine.knn(qb, rb, 200)
        monkeypatch.delenv("SEGVLAD_KNN_RESCORE_REF")
        torch.cuda.synchronize()
        assert torch.equal(d2a, d2b) and torch.equal(ia, ib)
        rows = torch.arange(0, Nq, 97, device=DEV)
        d64, i64 = _brute_fp64(q[rows], r, 208)
        assert_knn_close(d2a[rows].cpu().numpy(), ia[rows].cpu().numpy(), d64, i64, k_check=200)


def test_inverted_rescore_falls_back_to_the_gather_when_the_pair_lists_do_not_fit():
    # 3000 exact copies of one row sit at the k-th-neighbour boundary of every query: ~3000 candidates per query survive the
    # select (> the 512 pairs per query row the inverted lists hold) -> the device-side flag leaves the re-score to the
    # gather kernel; the result must still be the exact (d2, idx)-ordered top-k: duplicates in index order
    Nq, Nr, D, k = 400, 20000, 256, 200
    q, r = synth.make_descriptor_bank(Nq, Nr, D, seed=47, planted=0, device=DEV)
    centre = torch.nn.functional.normalize(q.mean(0), dim=0)
    q = torch.nn.functional.normalize(q * 0.2 + centre, dim=1)          # all queries close to one direction
    dup_rows = torch.arange(5000, 8000, device=DEV)
    r[dup_rows] = centre                                                 # the 3000 nearest references of every query, all equal
    d2, idx = engine.knn(engine.Bank.prepare(q), engine.Bank.prepare(r), k)
    torch.cuda.synchronize()
    assert torch.equal(idx, dup_rows[:k].expand(Nq, k))                  # ties broken by index
    want = ((q - centre) ** 2).sum(1, keepdim=True).expand(Nq, k)
    np.testing.assert_allclose(d2.cpu().numpy(), want.cpu().numpy(), rtol=1e-5, atol=2e-6)
    d2s, idxs = engine.knn_simt(q, r, k)
    assert torch.equal(idx, idxs)


@pytest.mark.parametrize("D,force", [(512, "8"), (1536, "16"), (96, "8")])
def test_epilogue_width_variants_agree(monkeypatch, D, force):
    # 8 or 16 epilogue warps in the scan kernel (default: 16 for D <= 1024, else 8; csrc/knn.cu kEpi): the candidate sets may
    # be stored in a different order, the exact re-score + (d2, idx) selection makes the output identical
    q, r = synth.make_descriptor_bank(2000, 40000, D, seed=48, planted=200, device=DEV)
    qb, rb = engine.Bank.prepare(q), engine.Bank.prepare(r)
    d2a, ia = engine.knn(qb, rb, 200)
    monkeypatch.setenv("SEGVLAD_KNN_EPI", force)
    d2b, ib = engine.knn(qb, rb, 200)
    monkeypatch.setenv("SEGVLAD_KNN_CTAS", "1")
    d2c, ic = engine.knn(qb, rb, 200)
    torch.cuda.synchronize()
    assert torch.equal(d2a, d2b) and torch.equal(ia, ib) and torch.equal(d2a, d2c) and torch.equal(ia, ic)
