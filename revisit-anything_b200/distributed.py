"""Row-sharded reference bank across ranks (one process per GPU) -- SURVEY.md 8e.

The reference has no distributed code; the only axis that scales is the number of reference segments, so the
bank is split into contiguous row shards (one per rank), the queries and the seg->image map are replicated,
every rank searches its shard for ALL queries, and ONE all-gather of the per-shard top-k lists (packed as
int32 pairs: fp32 distance bits + global row) makes the full candidate set available everywhere; the k-way
merge and the vote then run locally.  With the deterministic (d2, idx) tie order the result is identical to
the single-GPU search by construction.

The numerical work is delegated to an `ops` object: `EngineOps` (CUDA kernels through the C ABI) in
production; the CPU test-suite injects a checker-backed ops object to exercise the sharding / packing /
gather / merge plumbing over gloo with world_size 2.
"""
from __future__ import annotations

from typing import List, Optional, Sequence, Tuple

import numpy as np
import torch
import torch.distributed as dist


def shard_bounds(n_rows: int, world: int) -> List[Tuple[int, int]]:
    """Contiguous, nearly equal row shards; shard g = [lo, hi)."""
    base, rem = divmod(n_rows, world)
    out, lo = [], 0
    for g in range(world):
        hi = lo + base + (1 if g < rem else 0)
        out.append((lo, hi))
        lo = hi
    return out


def pack_topk(d2: torch.Tensor, idx: torch.Tensor) -> torch.Tensor:
    """(fp32 d2, int64 global idx < 2^31) -> [Nq, k, 2] int32 payload for a single all-gather."""
    assert d2.dtype == torch.float32 and idx.dtype == torch.int64
    if idx.numel() and int(idx.max()) >= 2 ** 31:
        raise ValueError("global reference row does not fit int32")
    return torch.stack([d2.contiguous().view(torch.int32), idx.to(torch.int32)], dim=-1).contiguous()


def unpack_topk(payload: torch.Tensor) -> Tuple[torch.Tensor, torch.Tensor]:
    d2 = payload[..., 0].contiguous().view(torch.float32)
    idx = payload[..., 1].to(torch.int64)
    return d2, idx


class EngineOps:
    """CUDA implementation (libsegvlad.so)."""

    def __init__(self):
        from . import engine
        self.engine = engine

    def prepare(self, x):
        return self.engine.Bank.prepare(x)

    def search(self, qbank, rbank, k, row_offset):
        return self.engine.knn(qbank, rbank, k, row_offset=row_offset)

    def merge(self, d2_parts, idx_parts):
        return self.engine.merge_topk(d2_parts, idx_parts)

    def vote(self, idx, d2, qimg_offsets, rseg_to_rimg, n_rimg, n_pred, k_vote):
        res = self.engine.vote(idx, d2, qimg_offsets, rseg_to_rimg, n_rimg, n_pred=n_pred, k_vote=k_vote,
                               sims_is_d2=True)
        return res.preds


def gather_merge(ops, d2, idx, group=None):
    """Per-shard top-k lists -> ONE all-gather -> merged global top-k (identical on every rank)."""
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    if world == 1:
        return d2, idx
    payload = pack_topk(d2, idx)
    gathered = torch.empty((world * payload.shape[0],) + tuple(payload.shape[1:]), dtype=payload.dtype,
                           device=payload.device)
    dist.all_gather_into_tensor(gathered, payload, group=group)      # the single collective of the path
    d2_parts, idx_parts = unpack_topk(gathered.view((world,) + tuple(payload.shape)))
    return ops.merge(d2_parts, idx_parts)


def sharded_search(ops, qbank, local_rbank, row_offset: int, k: int, group=None):
    """Local top-k on this rank's shard -> one all-gather -> merged global top-k (on every rank)."""
    d2, idx = ops.search(qbank, local_rbank, k, row_offset)
    return gather_merge(ops, d2, idx, group)


def sharded_search_and_vote(ops, qbank, local_rbank, row_offset: int, qimg_offsets, rseg_to_rimg, n_rimg: int,
                            k_search: int = 200, k_vote: int = 50, n_pred: int = 5, group=None):
    d2, idx = sharded_search(ops, qbank, local_rbank, row_offset, k_search, group)
    preds = ops.vote(idx, d2, qimg_offsets, rseg_to_rimg, n_rimg, n_pred, k_vote)
    return d2, idx, preds
