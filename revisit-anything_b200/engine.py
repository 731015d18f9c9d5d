"""Device-resident API of the SegVLAD engine: thin torch-tensor wrappers over the C ABI.

Every function takes/returns CUDA tensors, launches on torch's current stream and allocates its
workspace from torch's caching allocator (the library itself never allocates).  PyTorch is plumbing
here (device memory, streams); all arithmetic of the hot path happens inside libsegvlad.so.
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass
from typing import List, Optional, Sequence, Tuple

import numpy as np
import torch

from . import _lib
from ._lib import OUT_F32, OUT_F64, TOKENS_DN, TOKENS_ND, TOKENS_PRENORMALIZED, check, lib


def _ptr(t: Optional[torch.Tensor]):
    return None if t is None else C.c_void_p(t.data_ptr())


def _stream():
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def _need_cuda(*ts):
    for t in ts:
        if t is not None and not t.is_cuda:
            raise ValueError("segvlad engine expects CUDA tensors (no CPU fallback)")


def _ws(nbytes: int, device) -> torch.Tensor:
    return torch.empty(max(int(nbytes), 256), dtype=torch.uint8, device=device)


# --------------------------------------------------------------------------------------------
# membership
# --------------------------------------------------------------------------------------------
def mask_to_membership(masks_u8: torch.Tensor, H: int, W: int, patch: int = 14) -> torch.Tensor:
    """[S,Hm,Wm] uint8/bool CUDA pixel masks -> [S, ceil(N/32)] int32 membership bitmask
    (func_vpr.py:1088-1092 semantics; kernel segvlad_mask_to_membership)."""
    _need_cuda(masks_u8)
    m = masks_u8.to(torch.uint8).contiguous()
    S, Hm, Wm = m.shape
    N = (H // patch) * (W // patch)
    bits = torch.zeros((S, (N + 31) // 32), dtype=torch.int32, device=m.device)
    check(lib().segvlad_mask_to_membership(_ptr(m), S, Hm, Wm, H, W, patch, _ptr(bits), _stream()),
          "segvlad_mask_to_membership")
    return bits


def mask_centroids(masks_u8: torch.Tensor) -> torch.Tensor:
    """[S,Hm,Wm] uint8/bool CUDA pixel masks -> [S,2] fp64 (x, y) centroids, the `np.nonzero(m).mean` of func_vpr.py:1314."""
    _need_cuda(masks_u8)
    m = masks_u8.to(torch.uint8).contiguous()
    S, Hm, Wm = m.shape
    out = torch.empty((S, 2), dtype=torch.float64, device=m.device)
    check(lib().segvlad_mask_centroids(_ptr(m), S, Hm, Wm, _ptr(out), _stream()), "segvlad_mask_centroids")
    return out


def pack_membership(member_bool: torch.Tensor) -> torch.Tensor:
    """[S,N] bool CUDA -> [S, ceil(N/32)] int32 bit rows (format conversion for callers that already hold
    the reference's `mask_idx` tensor, e.g. the vlad_single drop-in)."""
    S, N = member_bool.shape
    Wd = (N + 31) // 32
    pad = torch.zeros((S, Wd * 32), dtype=torch.int64, device=member_bool.device)
    pad[:, :N] = member_bool.to(torch.int64)
    w = (pad.view(S, Wd, 32) << torch.arange(32, device=member_bool.device, dtype=torch.int64)).sum(-1)
    w = w - ((w >= 2 ** 31).to(torch.int64) << 32)      # reinterpret the uint32 word as int32
    return w.to(torch.int32).contiguous()


# --------------------------------------------------------------------------------------------
# aggregation
# --------------------------------------------------------------------------------------------
def aggregate_batch(tokens: torch.Tensor, N: int, D: int, token_layout: int, centers: torch.Tensor,
                    member_bits: torch.Tensor, seg_counts: Sequence[int],
                    adj: Optional[Sequence[Optional[torch.Tensor]]] = None, out_dtype=torch.float64,
                    return_labels: bool = False):
    """Batched per-(Super)Segment VLAD.  tokens: n_images*N*D fp32 (layout per `token_layout`);
    member_bits [S_total, ceil(N/32)] int32; seg_counts: segments per image; adj: per-image [S_i,S_i]
    bool/uint8 CUDA tensors or None (order 0).  Returns [S_total, K*D] (+ labels [n_images, N])."""
    _need_cuda(tokens, centers, member_bits)
    B = len(seg_counts)
    K = centers.shape[0]
    dev = tokens.device
    tokens = tokens.contiguous().float()
    centers = centers.contiguous().float()
    member_bits = member_bits.contiguous()
    seg_off = np.zeros(B + 1, dtype=np.int32)
    seg_off[1:] = np.cumsum(np.asarray(seg_counts, dtype=np.int64))
    S_total = int(seg_off[-1])
    assert tokens.numel() == B * N * D, "tokens size mismatch"
    assert member_bits.shape == (S_total, (N + 31) // 32), "member_bits shape mismatch"
    adj_flat = None
    if adj is not None and any(a is not None for a in adj):
        parts = []
        for b in range(B):
            Si = int(seg_counts[b])
            a = adj[b]
            if a is None:
                a = torch.eye(Si, dtype=torch.uint8, device=dev)
            assert a.shape == (Si, Si), "adjacency shape mismatch"
            parts.append((a.to(dev) != 0).to(torch.uint8).reshape(-1))   # any nonzero entry counts (reference: .bool())
        adj_flat = torch.cat(parts).contiguous() if parts else None
    odt = OUT_F64 if out_dtype == torch.float64 else OUT_F32
    out = torch.empty((S_total, K * D), dtype=torch.float64 if odt == OUT_F64 else torch.float32, device=dev)
    labels = torch.empty((B, N), dtype=torch.int32, device=dev) if return_labels else None
    nbytes = lib().segvlad_aggregate_workspace_bytes(B, N, D, K, S_total)
    ws = _ws(nbytes, dev)
    check(lib().segvlad_aggregate_batch(_ptr(tokens), B, N, D, token_layout, _ptr(centers), K, _ptr(member_bits),
                                        seg_off.ctypes.data_as(C.c_void_p), _ptr(adj_flat), _ptr(out), odt,
                                        _ptr(labels), _ptr(ws), ws.numel(), _stream()),
          "segvlad_aggregate_batch")
    return (out, labels) if return_labels else out


def aggregate_project_pca(tokens: torch.Tensor, N: int, D: int, token_layout: int, centers: torch.Tensor,
                          member_bits: torch.Tensor, seg_counts: Sequence[int],
                          adj: Optional[Sequence[Optional[torch.Tensor]]], components: torch.Tensor, mean: torch.Tensor,
                          explained_variance: torch.Tensor, normalize_rows: bool = False) -> torch.Tensor:
    """Aggregation fused with the PCA-whitening projection: the aggregation epilogue writes
    (descriptor - mean) as the bf16 planes the tensor-core projection consumes (segvlad_aggregate_batch_pca), the
    projection takes both operands by TMA (segvlad_pca_project_planes).  The [S, K*D] fp64 descriptor matrix is never
    written.  Returns Y [S_total, D_out] fp64.  Use pca_fusable() first."""
    _need_cuda(tokens, centers, member_bits, components, mean, explained_variance)
    B = len(seg_counts)
    K = centers.shape[0]
    dev = tokens.device
    tokens = tokens.contiguous().float()
    centers = centers.contiguous().float()
    member_bits = member_bits.contiguous()
    seg_off = np.zeros(B + 1, dtype=np.int32)
    seg_off[1:] = np.cumsum(np.asarray(seg_counts, dtype=np.int64))
    S_total = int(seg_off[-1])
    W = components.contiguous().float()
    Dout, Din = W.shape
    assert Din == K * D and tokens.numel() == B * N * D and member_bits.shape == (S_total, (N + 31) // 32)
    adj_flat = None
    if adj is not None and any(a is not None for a in adj):
        parts = []
        for b in range(B):
            Si = int(seg_counts[b])
            a = adj[b] if adj[b] is not None else torch.eye(Si, dtype=torch.uint8, device=dev)
            parts.append((a.to(dev) != 0).to(torch.uint8).reshape(-1))
        adj_flat = torch.cat(parts).contiguous()
    mean32 = _mean_f32(mean)
    xp = torch.empty((3, S_total, Din), dtype=torch.bfloat16, device=dev)
    ws = _ws(lib().segvlad_aggregate_workspace_bytes(B, N, D, K, S_total), dev)
    check(lib().segvlad_aggregate_batch_pca(_ptr(tokens), B, N, D, token_layout, _ptr(centers), K, _ptr(member_bits),
                                            seg_off.ctypes.data_as(C.c_void_p), _ptr(adj_flat), _ptr(mean32), _ptr(xp), None,
                                            _ptr(ws), ws.numel(), _stream()), "segvlad_aggregate_batch_pca")
    Y = torch.empty((S_total, Dout), dtype=torch.float64, device=dev)
    ws2 = _ws(lib().segvlad_pca_tc_workspace_bytes(S_total, Din, Dout), dev)
    check(lib().segvlad_pca_project_planes(_ptr(xp), S_total, Din, _ptr(_pca_planes(W)),
                                           _ptr(explained_variance.contiguous().float()), Dout, int(normalize_rows), _ptr(Y),
                                           _ptr(ws2), ws2.numel(), _stream()), "segvlad_pca_project_planes")
    return Y


def pca_fusable(D: int, K: int, Dout: int) -> bool:
    """The fused aggregation -> projection path needs the tensor-core kernels of both stages (D_t % 64 == 0)."""
    import os
    if os.environ.get("SEGVLAD_PCA_FUSED", "1") == "0" or os.environ.get("SEGVLAD_AGG_TC", "1") == "0":
        return False
    return D % 64 == 0 and bool(lib().segvlad_pca_tc_supported(K * D, Dout))


_MEAN32 = {}


def _mean_f32(mean: torch.Tensor) -> torch.Tensor:
    key = (mean.data_ptr(), mean.numel(), mean._version, mean.device.index)
    ent = _MEAN32.get(key)
    if ent is None:
        _MEAN32.clear()
        ent = _MEAN32[key] = (mean.to(torch.float32).contiguous(), mean)
    return ent[0]


def aggregate_residuals(residuals: torch.Tensor, labels: torch.Tensor, N: int, D: int, K: int,
                        member_bits: torch.Tensor, seg_counts: Sequence[int],
                        adj: Optional[Sequence[Optional[torch.Tensor]]] = None, out_dtype=torch.float64):
    """Inner aggregation stage on caller-provided residual rows [n_images*N, D] and labels
    (vlad_matmuls_per_cluster semantics, func_vpr.py:1181-1210)."""
    _need_cuda(residuals, labels, member_bits)
    B = len(seg_counts)
    dev = residuals.device
    residuals = residuals.contiguous().float()
    labels = labels.contiguous().to(torch.int32)
    seg_off = np.zeros(B + 1, dtype=np.int32)
    seg_off[1:] = np.cumsum(np.asarray(seg_counts, dtype=np.int64))
    S_total = int(seg_off[-1])
    adj_flat = None
    if adj is not None and any(a is not None for a in adj):
        parts = []
        for b in range(B):
            Si = int(seg_counts[b])
            a = adj[b] if adj[b] is not None else torch.eye(Si, dtype=torch.uint8, device=dev)
            parts.append((a.to(dev) != 0).to(torch.uint8).reshape(-1))
        adj_flat = torch.cat(parts).contiguous()
    odt = OUT_F64 if out_dtype == torch.float64 else OUT_F32
    out = torch.empty((S_total, K * D), dtype=torch.float64 if odt == OUT_F64 else torch.float32, device=dev)
    ws = _ws(lib().segvlad_aggregate_workspace_bytes(B, N, D, K, S_total), dev)
    check(lib().segvlad_aggregate_residuals(_ptr(residuals), _ptr(labels), B, N, D, K, _ptr(member_bits.contiguous()),
                                            seg_off.ctypes.data_as(C.c_void_p), _ptr(adj_flat), _ptr(out), odt,
                                            _ptr(ws), ws.numel(), _stream()), "segvlad_aggregate_residuals")
    return out


# --------------------------------------------------------------------------------------------
# matching
# --------------------------------------------------------------------------------------------
@dataclass
class Bank:
    """Kernel-ready resident form of an [n, D] descriptor matrix (scaled fp16 plane, norms / scan error bounds,
    fp32 rows for the exact re-score)."""
    buf: torch.Tensor
    n: int
    D: int
    src: Optional[torch.Tensor] = None     # view banks: the fp32 matrix the bank references (kept alive here)

    @staticmethod
    def prepare(x: torch.Tensor, copy: bool = False) -> "Bank":
        """fp32 [n, D] CUDA matrix -> bank.  By default the bank REFERENCES x for the exact re-score (no 4 n D byte copy;
        do not modify x while the bank is in use); copy=True stores its own copy of the rows."""
        _need_cuda(x)
        x = x.contiguous().float()
        n, D = x.shape
        buf = _ws(lib().segvlad_bank_bytes(n, D), x.device)
        if copy or x.data_ptr() % 16 != 0:
            check(lib().segvlad_bank_prepare(_ptr(x), n, D, _ptr(buf), _stream()), "segvlad_bank_prepare")
            return Bank(buf, n, D)
        check(lib().segvlad_bank_prepare_view(_ptr(x), n, D, _ptr(buf), _stream()), "segvlad_bank_prepare_view")
        return Bank(buf, n, D, x)

    @staticmethod
    def prepare_f64(x: torch.Tensor, normalize_rows: bool = False) -> "Bank":
        """fp64 descriptors (the reference's segFtVLAD dtype): optional normalizeFeat in fp64, fp32 cast, split."""
        _need_cuda(x)
        x = x.contiguous().double()
        n, D = x.shape
        buf = _ws(lib().segvlad_bank_bytes(n, D), x.device)
        check(lib().segvlad_bank_prepare_f64(_ptr(x), n, D, int(normalize_rows), _ptr(buf), _stream()),
              "segvlad_bank_prepare_f64")
        return Bank(buf, n, D)


def knn(q: Bank, r: Bank, k: int, row_offset: int = 0) -> Tuple[torch.Tensor, torch.Tensor]:
    """tcgen05 all-pairs squared-L2 (one fp16 pass, error-bounded filter, exact fp32 re-score) + top-k.  Returns (d2 [Nq,k] fp32 ascending,
    idx [Nq,k] int64 global rows; (+inf,-1) padded)."""
    assert q.D == r.D
    dev = q.buf.device
    d2 = torch.empty((q.n, k), dtype=torch.float32, device=dev)
    idx = torch.empty((q.n, k), dtype=torch.int64, device=dev)
    ws = _ws(lib().segvlad_knn_workspace_bytes(q.n, r.n, q.D, k), dev)
    check(lib().segvlad_knn(_ptr(q.buf), q.n, _ptr(r.buf), r.n, int(row_offset), q.D, k, _ptr(d2), _ptr(idx),
                            _ptr(ws), ws.numel(), _stream()), "segvlad_knn")
    return d2, idx


def knn_async(q: Bank, r: Bank, k: int, row_offset: int = 0, schedule: int = 0, packed_out: Optional[torch.Tensor] = None,
              overflow: Optional[torch.Tensor] = None, unpacked: bool = True):
    """segvlad_knn_async: the same search with NO host synchronisation.  Returns (d2, idx, overflow): `overflow` is a
    device int32[1] the caller must read at its next synchronisation point (after the vote); if it is non-zero the
    candidate buffers overflowed under the fast schedule and the call has to be repeated with schedule=1.
    packed_out: optional int64 CUDA tensor with >= Nq*k elements that receives the packed lists
    ((fp32 bits of d2) << 32 | uint32(int32 global row)) -- e.g. this rank's slot of the all-gather buffer."""
    assert q.D == r.D
    dev = q.buf.device
    d2 = torch.empty((q.n, k), dtype=torch.float32, device=dev) if unpacked else None
    idx = torch.empty((q.n, k), dtype=torch.int64, device=dev) if unpacked else None
    if overflow is None:
        overflow = torch.empty(1, dtype=torch.int32, device=dev)
    if packed_out is not None:
        assert packed_out.is_cuda and packed_out.dtype == torch.int64 and packed_out.is_contiguous()
        assert packed_out.numel() >= q.n * k
    ws = _ws(lib().segvlad_knn_workspace_bytes(q.n, r.n, q.D, k), dev)
    check(lib().segvlad_knn_async(_ptr(q.buf), q.n, _ptr(r.buf), r.n, int(row_offset), q.D, k, int(schedule), _ptr(d2),
                                  _ptr(idx), _ptr(packed_out), _ptr(overflow), _ptr(ws), ws.numel(), _stream()),
          "segvlad_knn_async")
    return d2, idx, overflow


def merge_topk_packed(parts: torch.Tensor, Nq: int, k: int, want_packed: bool = False):
    """parts: int64 CUDA [G, stride] (stride >= Nq*k), shard g's sorted packed lists at parts[g, :Nq*k] (the gathered
    buffer of the row-sharded search) -> merged (d2 [Nq,k] fp32, idx [Nq,k] int64) (+ packed [Nq,k] int64)."""
    _need_cuda(parts)
    assert parts.dtype == torch.int64 and parts.is_contiguous() and parts.dim() == 2 and parts.shape[1] >= Nq * k
    G, stride = parts.shape
    d2 = torch.empty((Nq, k), dtype=torch.float32, device=parts.device)
    idx = torch.empty((Nq, k), dtype=torch.int64, device=parts.device)
    packed = torch.empty((Nq, k), dtype=torch.int64, device=parts.device) if want_packed else None
    check(lib().segvlad_merge_topk_packed(_ptr(parts), G, stride, Nq, k, _ptr(d2), _ptr(idx), _ptr(packed), _stream()),
          "segvlad_merge_topk_packed")
    return (d2, idx, packed) if want_packed else (d2, idx)


def knn_from_host(q_host: torch.Tensor, r_host: torch.Tensor, k: int, row_offset: int = 0, device=None):
    """Search with both fp32 descriptor matrices in host memory (pin them for full PCIe rate): the H2D transfer
    of the reference bank is pipelined with the tensor-core scan.  Returns (d2, idx, qbank, rbank) -- the banks are
    resident afterwards."""
    assert not q_host.is_cuda and not r_host.is_cuda and q_host.dtype == torch.float32 and r_host.dtype == torch.float32
    q_host, r_host = q_host.contiguous(), r_host.contiguous()
    dev = torch.device("cuda", torch.cuda.current_device()) if device is None else device
    Nq, D = q_host.shape
    Nr = r_host.shape[0]
    qb = Bank(_ws(lib().segvlad_bank_bytes(Nq, D), dev), Nq, D)
    rb = Bank(_ws(lib().segvlad_bank_bytes(Nr, D), dev), Nr, D)
    d2 = torch.empty((Nq, k), dtype=torch.float32, device=dev)
    idx = torch.empty((Nq, k), dtype=torch.int64, device=dev)
    ws = _ws(lib().segvlad_knn_workspace_bytes(Nq, Nr, D, k), dev)
    check(lib().segvlad_knn_from_host(C.c_void_p(q_host.data_ptr()), Nq, C.c_void_p(r_host.data_ptr()), Nr,
                                      int(row_offset), D, k, _ptr(qb.buf), _ptr(rb.buf), _ptr(d2), _ptr(idx), _ptr(ws),
                                      ws.numel(), _stream(), None), "segvlad_knn_from_host")
    return d2, idx, qb, rb


def knn_debug_approx(q: Bank, r: Bank) -> Tuple[torch.Tensor, torch.Tensor]:
    """Test hook (r.n <= 4096): the tensor-core pass's approximate d2 for every pair [Nq, Nr] and the per-query
    error bound [Nq] the filter assumes on |approx - fp32 value|."""
    assert q.D == r.D
    dev = q.buf.device
    approx = torch.empty((q.n, r.n), dtype=torch.float32, device=dev)
    bound = torch.empty((q.n,), dtype=torch.float32, device=dev)
    ws = _ws(lib().segvlad_knn_workspace_bytes(q.n, r.n, q.D, 1), dev)
    check(lib().segvlad_knn_debug_approx(_ptr(q.buf), q.n, _ptr(r.buf), r.n, q.D, _ptr(approx), _ptr(bound), _ptr(ws),
                                         ws.numel(), _stream()), "segvlad_knn_debug_approx")
    return approx, bound


def knn_simt(q: torch.Tensor, r: torch.Tensor, k: int, row_offset: int = 0):
    """fp32 FFMA cross-check path (same selection machinery, no tensor cores)."""
    _need_cuda(q, r)
    q = q.contiguous().float()
    r = r.contiguous().float()
    dev = q.device
    d2 = torch.empty((q.shape[0], k), dtype=torch.float32, device=dev)
    idx = torch.empty((q.shape[0], k), dtype=torch.int64, device=dev)
    ws = _ws(lib().segvlad_knn_workspace_bytes(q.shape[0], r.shape[0], q.shape[1], k), dev)
    check(lib().segvlad_knn_simt(_ptr(q), q.shape[0], _ptr(r), r.shape[0], int(row_offset), q.shape[1], k,
                                 _ptr(d2), _ptr(idx), _ptr(ws), ws.numel(), _stream()), "segvlad_knn_simt")
    return d2, idx


def merge_topk(d2_parts: torch.Tensor, idx_parts: torch.Tensor):
    """[G,Nq,k] per-shard lists -> merged [Nq,k] ((d2, idx) ascending)."""
    _need_cuda(d2_parts, idx_parts)
    G, Nq, k = d2_parts.shape
    d2_parts = d2_parts.contiguous().float()
    idx_parts = idx_parts.contiguous().to(torch.int64)
    d2 = torch.empty((Nq, k), dtype=torch.float32, device=d2_parts.device)
    idx = torch.empty((Nq, k), dtype=torch.int64, device=d2_parts.device)
    check(lib().segvlad_merge_topk(_ptr(d2_parts), _ptr(idx_parts), G, Nq, k, _ptr(d2), _ptr(idx), _stream()),
          "segvlad_merge_topk")
    return d2, idx


# --------------------------------------------------------------------------------------------
# vote
# --------------------------------------------------------------------------------------------
@dataclass
class VoteResult:
    preds: torch.Tensor            # [n_qimg, n_pred] int32, -1 padded
    pred_scores: torch.Tensor      # [n_qimg, n_pred] fp64
    scores: Optional[torch.Tensor]  # [n_qimg, n_rimg] fp64 (dense) or None
    counts: Optional[torch.Tensor]  # [n_qimg, n_rimg] int32 (dense) or None
    minmax: torch.Tensor           # [2] fp32


def vote(matches: torch.Tensor, sims: torch.Tensor, qimg_offsets: torch.Tensor, rseg_to_rimg: torch.Tensor,
         n_rimg: int, n_pred: int = 5, k_vote: int = 50, sims_is_d2: bool = False, dense: bool = False,
         max_segs: Optional[int] = None, qrow_index: Optional[torch.Tensor] = None) -> VoteResult:
    """Similarity-weighted segment->image vote (get_matches 'max_seg_topk_wt_borda_Im' semantics) + hit
    counts.  matches/sims: [Nq, ld] int64 / fp32 (first k_vote columns used); qimg_offsets [n_qimg+1] int32;
    qrow_index (optional int32): rows of matches/sims in concatenated segRangeQuery order (None: identity).  The
    min/max normalisation always spans all Nq rows (func_vpr.py:211-212)."""
    _need_cuda(matches, sims, qimg_offsets, rseg_to_rimg)
    assert matches.dtype == torch.int64 and sims.dtype == torch.float32
    assert matches.stride(1) == 1 and sims.stride(1) == 1 and matches.stride(0) == sims.stride(0)
    Nq, ld = matches.shape[0], matches.stride(0) if matches.shape[0] > 1 else matches.shape[1]
    k_vote = min(k_vote, matches.shape[1])
    dev = matches.device
    offsets_in = qimg_offsets
    if max_segs is None and not qimg_offsets.is_cuda and qimg_offsets.numel() > 1:
        max_segs = int((qimg_offsets[1:] - qimg_offsets[:-1]).max())      # host offsets: no device read at all
    qimg_offsets = qimg_offsets.to(device=dev, dtype=torch.int32).contiguous()
    rseg_to_rimg = rseg_to_rimg.to(device=dev, dtype=torch.int32).contiguous()
    n_qimg = qimg_offsets.numel() - 1
    if qrow_index is not None:
        qrow_index = qrow_index.to(device=dev, dtype=torch.int32).contiguous()
    if max_segs is None:
        # a device -> host read; remembered ON the offsets tensor object (with its version counter), so that a pipeline voting
        # repeatedly with the same query-image layout (every bench step, every shard merge) synchronises for it once, not
        # before every vote -- and a different tensor can never pick up a stale value
        memo = getattr(offsets_in, "_segvlad_max_segs", None)
        if memo is not None and memo[0] == offsets_in._version:
            max_segs = memo[1]
        else:
            max_segs = int((qimg_offsets[1:] - qimg_offsets[:-1]).max().item()) if n_qimg > 0 else 0
            try:
                offsets_in._segvlad_max_segs = (offsets_in._version, max_segs)
            except AttributeError:
                pass
    preds = torch.empty((n_qimg, n_pred), dtype=torch.int32, device=dev)
    pscores = torch.empty((n_qimg, n_pred), dtype=torch.float64, device=dev)
    scores = torch.empty((n_qimg, n_rimg), dtype=torch.float64, device=dev) if dense else None
    counts = torch.empty((n_qimg, n_rimg), dtype=torch.int32, device=dev) if dense else None
    mm = torch.empty(2, dtype=torch.float32, device=dev)
    ws = _ws(lib().segvlad_vote_workspace_bytes(Nq, k_vote, n_qimg, max_segs), dev)
    check(lib().segvlad_vote(_ptr(matches), _ptr(sims), ld, int(sims_is_d2), k_vote, Nq, _ptr(qimg_offsets),
                             _ptr(qrow_index), n_qimg, max_segs, _ptr(rseg_to_rimg), rseg_to_rimg.numel(), n_rimg, n_pred, _ptr(preds),
                             _ptr(pscores), _ptr(scores), _ptr(counts), _ptr(mm), _ptr(ws), ws.numel(), _stream()),
          "segvlad_vote")
    return VoteResult(preds, pscores, scores, counts, mm)


# --------------------------------------------------------------------------------------------
# PCA-whitening projection (row f1)
# --------------------------------------------------------------------------------------------
_PCA_PLANES = {}   # components tensor -> bf16 planes of the tensor-core projection (one model at a time)


def _pca_planes(W: torch.Tensor) -> torch.Tensor:
    key = (W.data_ptr(), tuple(W.shape), W._version, W.device.index)
    ent = _PCA_PLANES.get(key)
    if ent is None:
        Dout, Din = W.shape
        planes = _ws(lib().segvlad_pca_planes_bytes(Din, Dout), W.device)
        check(lib().segvlad_pca_prepare_planes(_ptr(W), Dout, Din, _ptr(planes), _stream()), "segvlad_pca_prepare_planes")
        _PCA_PLANES.clear()
        ent = _PCA_PLANES[key] = (planes, W)     # W is kept alive: its address is the key
    return ent[0]


def pca_project(X: torch.Tensor, components: torch.Tensor, mean: torch.Tensor, explained_variance: torch.Tensor,
                normalize_rows: bool = False) -> torch.Tensor:
    """Y = ((X - mean) @ components^T) / sqrt(explained_variance) (sklearn PCA.transform, whiten=True); optional
    normalizeFeat.  X [S, D_in] CUDA (any float dtype, read as fp64), returns [S, D_out] fp64 CUDA.  Default: tcgen05 kernel
    (fp32-equivalent split operands, fp64 across 512-channel chunks, ~1e-6 relative); SEGVLAD_PCA_TC=0 or an unsupported
    D_in selects the fp64 CUDA-core kernel."""
    _need_cuda(X, components, mean, explained_variance)
    X = X.contiguous().double()
    S, Din = X.shape
    W = components.contiguous().float()
    Dout = W.shape[0]
    mu = mean.contiguous().double()
    ev = explained_variance.contiguous().float()
    Y = torch.empty((S, Dout), dtype=torch.float64, device=X.device)
    if lib().segvlad_pca_tc_supported(Din, Dout):
        planes = _pca_planes(W)
        ws = _ws(lib().segvlad_pca_tc_workspace_bytes(S, Din, Dout), X.device)
        check(lib().segvlad_pca_project_tc(_ptr(X), S, Din, _ptr(planes), _ptr(mu), _ptr(ev), Dout, int(normalize_rows),
                                           _ptr(Y), _ptr(ws), ws.numel(), _stream()), "segvlad_pca_project_tc")
        return Y
    ws = _ws(lib().segvlad_pca_workspace_bytes(S, Din, Dout), X.device)
    check(lib().segvlad_pca_project(_ptr(X), S, Din, _ptr(W), _ptr(mu), _ptr(ev), Dout, int(normalize_rows), _ptr(Y),
                                    _ptr(ws), ws.numel(), _stream()), "segvlad_pca_project")
    return Y


# --------------------------------------------------------------------------------------------
# NetVLAD + anti-burst (config 5)
# --------------------------------------------------------------------------------------------
def netvlad_antiburst(x: torch.Tensor, centroids: torch.Tensor, conv_weight: torch.Tensor,
                      ab_params=(8.0, 7.0, 1.0)) -> torch.Tensor:
    """x [B,D,N] (or [B,D,H,W]) fp32 CUDA -> [B, K*D] fp32 (aggregation.py:266-361 semantics, antiburst on)."""
    _need_cuda(x, centroids, conv_weight)
    B, D = x.shape[0], x.shape[1]
    x = x.reshape(B, D, -1).contiguous().float()
    N = x.shape[2]
    K = centroids.shape[0]
    cw = conv_weight.reshape(K, D).contiguous().float()
    out = torch.empty((B, K * D), dtype=torch.float32, device=x.device)
    ws = _ws(lib().segvlad_netvlad_workspace_bytes(B, N, D, K), x.device)
    check(lib().segvlad_netvlad_antiburst(_ptr(x), B, N, D, _ptr(centroids.contiguous().float()), _ptr(cw), K,
                                          float(ab_params[0]), float(ab_params[1]), float(ab_params[2]), _ptr(out),
                                          _ptr(ws), ws.numel(), _stream()), "segvlad_netvlad_antiburst")
    return out
