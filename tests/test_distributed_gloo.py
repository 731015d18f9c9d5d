"""World-size-2 gloo test (CPU) of the row-sharded search plumbing: shard bounds, global row offsets,
int32 packing, the single all-gather, k-way merge order and the vote on the merged list.  The per-shard
search / merge / vote arithmetic is supplied by the oracle here (the CUDA ops are covered by -m gpu tests)."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from oracle import segvlad_oracle as O
from revisit_anything_b200 import distributed as D
from revisit_anything_b200 import synth


class OracleOps(D.OpsBase):
    def prepare(self, x):
        return x

    def search(self, q, r, k, row_offset):
        d2, idx = O.flat_l2_search(q.numpy(), r.numpy(), k)
        idx = np.where(idx >= 0, idx + row_offset, -1)
        return torch.from_numpy(d2), torch.from_numpy(idx)

    def merge(self, d2_parts, idx_parts):
        G, Nq, k = d2_parts.shape
        d = d2_parts.permute(1, 0, 2).reshape(Nq, G * k).numpy()
        i = idx_parts.permute(1, 0, 2).reshape(Nq, G * k).numpy()
        key_i = np.where(i < 0, np.iinfo(np.int64).max, i)
        order = np.lexsort((key_i, d), axis=1)[:, :k]
        return torch.from_numpy(np.take_along_axis(d, order, 1)), torch.from_numpy(np.take_along_axis(i, order, 1))

    def vote(self, idx, d2, qimg_offsets, rseg_to_rimg, n_rimg, n_pred, k_vote):
        sims, matches = O.sims_from_d2(d2.numpy(), idx.numpy(), k_vote)
        off = qimg_offsets.numpy()
        rng = [np.arange(off[i], off[i + 1]) for i in range(len(off) - 1)]
        preds = O.get_matches_wt_borda(matches, len(rng), sims, rng, rseg_to_rimg.numpy(), n=n_pred)
        return preds


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, q, r, imq_off, imr, k, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        torch.set_num_threads(2)
        lo, hi = D.shard_bounds(r.shape[0], world)[rank]
        ops = OracleOps()
        d2, idx, preds = D.sharded_search_and_vote(ops, q, r[lo:hi], lo, imq_off, imr, int(imr.max()) + 1,
                                                   k_search=k, k_vote=50, n_pred=5)
        out[rank] = (d2.numpy(), idx.numpy(), [list(map(int, p)) for p in preds])
    finally:
        dist.destroy_process_group()


def test_shard_bounds():
    assert D.shard_bounds(10, 4) == [(0, 3), (3, 6), (6, 8), (8, 10)]
    assert D.shard_bounds(3, 4) == [(0, 1), (1, 2), (2, 3), (3, 3)]
    b = D.shard_bounds(8_000_000, 8)
    assert b[0] == (0, 1_000_000) and b[-1][1] == 8_000_000


def test_pack_roundtrip():
    d2 = torch.tensor([[0.0, 1.5, float("inf")]], dtype=torch.float32)
    idx = torch.tensor([[7, 2 ** 31 - 1, -1]], dtype=torch.int64)
    keys = D.pack_topk(d2, idx)
    assert bool((keys[0, 1:] > keys[0, :-1]).all())             # sorted lists stay sorted as integer keys; padding is last
    a, b = D.unpack_topk(keys)
    assert torch.equal(a, d2) and torch.equal(b, idx)
    with pytest.raises(ValueError):
        D.pack_topk(d2, torch.tensor([[2 ** 31, 0, 0]], dtype=torch.int64))


def test_pack_order_property():
    # property behind the k-way merge of packed lists: integer order of the keys == (d2 ascending, then row as uint32
    # ascending: the -1 padding last), and the round trip is exact -- for any non-negative fp32 d2 (zero, sub-normal, inf)
    hyp = pytest.importorskip("hypothesis")
    st = hyp.strategies

    @hyp.settings(max_examples=200, deadline=None)
    @hyp.given(st.lists(st.tuples(st.floats(min_value=0.0, allow_nan=False, width=32), st.integers(-1, 2 ** 31 - 1)),
                        min_size=1, max_size=48))
    def check(pairs):
        d2 = torch.tensor([[p[0] for p in pairs]], dtype=torch.float32)
        idx = torch.tensor([[p[1] for p in pairs]], dtype=torch.int64)
        keys = D.pack_topk(d2, idx)
        a, b = D.unpack_topk(keys)
        assert torch.equal(a.view(torch.int32), d2.view(torch.int32)) and torch.equal(b, idx)
        want = sorted(range(len(pairs)), key=lambda i: (float(d2[0, i]), int(idx[0, i]) & 0xFFFFFFFF, i))
        got = sorted(range(len(pairs)), key=lambda i: (int(keys[0, i]), i))
        assert want == got

    check()


@pytest.mark.parametrize("world,k", [(2, 64), (4, 80)])
def test_world2_gloo_matches_single_process(world, k):
    # world 4: uneven shards (75, 75, 75, 74 rows), each SHORTER than k = 80 -> padded (+inf, -1) lists enter the merge
    q, r, imq, imr = synth.make_structured_bank(n_ref_img=23, n_qry_img=6, segs_per_img=13, D=48, seed=3, noise=1.0)
    r[40] = r[200]                                    # an exact cross-shard tie: merge must keep idx order
    imq_off = torch.from_numpy(np.arange(0, 6 * 13 + 1, 13).astype(np.int32))
    imr_t = torch.from_numpy(imr)
    mgr = mp.Manager()
    out = mgr.dict()
    port = _free_port()
    mp.spawn(_worker, args=(world, port, q, r, imq_off, imr_t, k, out), nprocs=world, join=True)
    D2, I = O.flat_l2_search(q.numpy(), r.numpy(), k)
    sims, matches = O.sims_from_d2(D2, I, 50)
    rng = [np.arange(i * 13, (i + 1) * 13) for i in range(6)]
    preds = [list(map(int, p)) for p in O.get_matches_wt_borda(matches, 6, sims, rng, imr, n=5)]
    for rank in range(world):
        d2, idx, p = out[rank]
        np.testing.assert_array_equal(d2, D2)
        np.testing.assert_array_equal(idx, I)
        assert p == preds
