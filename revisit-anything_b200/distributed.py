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


TRAILER = 8   # int64 words behind a shard's [Nq * k] packed lists in the gather buffer; word 0 = the shard's overflow flag


def pack_topk(d2: torch.Tensor, idx: torch.Tensor) -> torch.Tensor:
    """(fp32 d2, int64 global idx < 2^31) -> int64 keys (fp32 bits << 32 | uint32(idx)): what the final selection kernel
    writes into the all-gather send buffer (host-side twin for checker ops and tests).  d2 >= 0, so the keys of a sorted
    list ascend as integers; padding (+inf, -1) becomes the largest key."""
    assert d2.dtype == torch.float32 and idx.dtype == torch.int64
    if idx.numel() and int(idx.max()) >= 2 ** 31:
        raise ValueError("global reference row does not fit int32")
    bits = d2.contiguous().view(torch.int32).to(torch.int64) & 0xFFFFFFFF
    return (bits << 32) | (idx & 0xFFFFFFFF)


def unpack_topk(keys: torch.Tensor) -> Tuple[torch.Tensor, torch.Tensor]:
    d2 = (keys >> 32).to(torch.int32).contiguous().view(torch.float32)
    low = keys & 0xFFFFFFFF
    idx = torch.where(low >= 2 ** 31, low - 2 ** 32, low)
    return d2, idx


class OpsBase:
    """Plumbing shared by the ops objects: an implementation provides search / merge / vote on unpacked lists (the CPU
    test-suite's checker-backed ops do); the packed gather-buffer protocol is derived from them here.  EngineOps overrides
    every method with the CUDA entry points that work on the packed layout directly."""

    def n_rows(self, bank) -> int:
        return bank.shape[0]

    def new_gather_buffer(self, world: int, stride: int, like) -> torch.Tensor:
        return torch.zeros((world, stride), dtype=torch.int64)

    def search_async(self, qbank, rbank, k, row_offset, schedule, payload):
        d2, idx = self.search(qbank, rbank, k, row_offset)
        if payload is None:
            return d2, idx, torch.zeros(1, dtype=torch.int64)
        n = d2.numel()
        payload[:n] = pack_topk(d2, idx).reshape(-1)
        payload[n:] = 0
        return None, None, payload[n:n + 1]

    def merge_packed(self, buf, Nq, k):
        d2p, idxp = unpack_topk(buf[:, :Nq * k].reshape(buf.shape[0], Nq, k))
        return self.merge(d2p, idxp)

    def overflowed(self, flags) -> bool:
        return bool(flags.any())


class EngineOps(OpsBase):
    """CUDA implementation (libsegvlad.so)."""

    def __init__(self):
        from . import engine
        self.engine = engine

    def prepare(self, x):
        return self.engine.Bank.prepare(x)

    def n_rows(self, bank) -> int:
        return bank.n

    def new_gather_buffer(self, world, stride, like):
        return torch.empty((world, stride), dtype=torch.int64, device=like.buf.device)

    def search(self, qbank, rbank, k, row_offset):
        return self.engine.knn(qbank, rbank, k, row_offset=row_offset)

    def search_async(self, qbank, rbank, k, row_offset, schedule, payload):
        """No host synchronisation.  payload (this rank's slot of the gather buffer) receives the packed lists straight from
        the final selection kernel, and the overflow flag in the first trailer word."""
        if payload is None:
            return self.engine.knn_async(qbank, rbank, k, row_offset, schedule)
        n = qbank.n * k
        payload[n:].zero_()
        flag = payload[n:n + 1].view(torch.int32)[:1]          # low half of the trailer word (little endian)
        self.engine.knn_async(qbank, rbank, k, row_offset, schedule, packed_out=payload, overflow=flag, unpacked=False)
        return None, None, payload[n:n + 1]

    def merge(self, d2_parts, idx_parts):
        return self.engine.merge_topk(d2_parts, idx_parts)

    def merge_packed(self, buf, Nq, k):
        return self.engine.merge_topk_packed(buf, Nq, k)

    def overflowed(self, flags) -> bool:
        return bool(flags.ne(0).any().item())                   # the step's one host synchronisation, AFTER the vote

    def vote(self, idx, d2, qimg_offsets, rseg_to_rimg, n_rimg, n_pred, k_vote):
        res = self.engine.vote(idx, d2, qimg_offsets, rseg_to_rimg, n_rimg, n_pred=n_pred, k_vote=k_vote,
                               sims_is_d2=True)
        return res.preds


def sharded_search(ops, qbank, local_rbank, row_offset: int, k: int, group=None, schedule: int = 0):
    """Local top-k on this rank's shard, written by the final selection kernel as packed (fp32 bits, int32 global row) keys
    straight into this rank's slot of the gather buffer -> ONE in-place all-gather -> k-way merge of the packed lists
    (identical on every rank).  Nothing between the search and the merge touches the host.  Returns (d2, idx, flags):
    `flags` holds every shard's overflow flag (after the gather all ranks see the same values, so they take the same
    decision without another collective); check it with ops.overflowed() after the vote."""
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    if world == 1:
        return ops.search_async(qbank, local_rbank, k, row_offset, schedule, None)
    rank = dist.get_rank(group)
    Nq = ops.n_rows(qbank)
    buf = ops.new_gather_buffer(world, Nq * k + TRAILER, qbank)
    ops.search_async(qbank, local_rbank, k, row_offset, schedule, buf[rank])
    mine = buf[rank] if buf.is_cuda else buf[rank].clone()       # NCCL gathers in place; gloo wants a separate input
    dist.all_gather_into_tensor(buf.view(-1), mine, group=group)   # the single collective of the path
    d2, idx = ops.merge_packed(buf, Nq, k)
    return d2, idx, buf[:, Nq * k]


def gather_merge(ops, d2, idx, group=None, max_row: Optional[int] = None):
    """Already-computed per-shard lists (e.g. of the host-streamed search, which returns unpacked results) -> packed ->
    ONE all-gather -> merged global top-k.  `max_row` (largest global row any shard can return) is checked on the host
    against the int32 range; without it the rows are checked on the device tensor (CPU tensors only -- never a device
    synchronisation)."""
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    if world == 1:
        return d2, idx
    if max_row is not None:
        if max_row >= 2 ** 31:
            raise ValueError("global reference row does not fit int32")
        bits = d2.contiguous().view(torch.int32).to(torch.int64) & 0xFFFFFFFF
        keys = (bits << 32) | (idx & 0xFFFFFFFF)
    else:
        keys = pack_topk(d2.cpu(), idx.cpu()).to(d2.device) if d2.is_cuda else pack_topk(d2, idx)
    Nq, k = d2.shape
    buf = torch.empty((world, Nq * k), dtype=torch.int64, device=d2.device)
    dist.all_gather_into_tensor(buf.view(-1), keys.reshape(-1).contiguous(), group=group)
    return ops.merge_packed(buf, Nq, k)


def sharded_search_and_vote(ops, qbank, local_rbank, row_offset: int, qimg_offsets, rseg_to_rimg, n_rimg: int,
                            k_search: int = 200, k_vote: int = 50, n_pred: int = 5, group=None):
    """search -> [all-gather -> merge] -> vote, enqueued without a host synchronisation in between; the overflow flags
    are read once, after the vote.  A set flag (adversarially ordered / massively duplicated bank under the fast chunk
    schedule) repeats the step with the conservative schedule on every rank."""
    for schedule in (0, 1):
        d2, idx, flags = sharded_search(ops, qbank, local_rbank, row_offset, k_search, group, schedule)
        preds = ops.vote(idx, d2, qimg_offsets, rseg_to_rimg, n_rimg, n_pred, k_vote)
        if not ops.overflowed(flags):
            break
    return d2, idx, preds
