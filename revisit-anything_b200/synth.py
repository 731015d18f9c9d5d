"""Seeded synthetic inputs for the SegVLAD hot path.

No real 17places tokens / SAM masks are available offline, so every test and bench config is a
shape-faithful synthetic: DINOv2-like tokens [1,D_t,dh,dw] fp32 (unit-norm over channels as stored by
the reference, func_vpr.py:561), SAM-like masks at half resolution (unions of rectangles), and
unit-norm segment-descriptor banks with planted near-duplicates for the matching stage.
"""
from __future__ import annotations

import numpy as np
import torch


def make_centers(K: int, D: int, seed: int = 0) -> torch.Tensor:
    """Un-normalised k-means-like centres with ||c|| ~ 0.5 (as the cached c_centers.pt)."""
    g = torch.Generator().manual_seed(1000 + seed)
    c = torch.randn(K, D, generator=g)
    c = 0.5 * c / c.norm(dim=1, keepdim=True)
    c = c * (0.8 + 0.4 * torch.rand(K, 1, generator=g))
    return c.contiguous()


def make_tokens(D: int, dh: int, dw: int, seed: int, centers: torch.Tensor | None = None,
                rank: int = 6, normalized: bool = True) -> torch.Tensor:
    """[1,D,dh,dw] fp32: randn + per-image low-rank component (uneven cluster population)
    (+ a pull towards random centres so assignment margins are well above fp32 noise)."""
    g = torch.Generator().manual_seed(2000 + seed)
    N = dh * dw
    x = torch.randn(D, N, generator=g)
    u = torch.randn(D, rank, generator=g)
    v = torch.randn(rank, N, generator=g)
    x = x + 1.5 * (u @ v) / rank ** 0.5
    if centers is not None:
        K = centers.shape[0]
        pick = torch.randint(0, K, (N,), generator=g)
        # skew the cluster histogram: half the tokens go to 4 clusters
        heavy = torch.randint(0, K, (4,), generator=g)
        sel = torch.rand(N, generator=g) < 0.5
        pick[sel] = heavy[torch.randint(0, 4, (int(sel.sum()),), generator=g)]
        cn = centers / centers.norm(dim=1, keepdim=True)
        x = x / x.norm(dim=0, keepdim=True) + 0.35 * cn[pick].T
    if normalized:
        x = x / x.norm(dim=0, keepdim=True).clamp_min(1e-12)
    return x.reshape(1, D, dh, dw).contiguous()


def make_tokens_skewed(D: int, dh: int, dw: int, seed: int, centers: torch.Tensor, frac: float,
                       heavy: int = 0) -> torch.Tensor:
    """[1,D,dh,dw] fp32 unit-norm tokens of a sky / road dominated image: a fraction `frac` of the tokens is pulled
    towards ONE centre (`heavy`), the rest towards random centres -> one cluster holds ~frac of the image."""
    g = torch.Generator().manual_seed(2500 + seed)
    N = dh * dw
    K = centers.shape[0]
    x = torch.randn(D, N, generator=g)
    x = x / x.norm(dim=0, keepdim=True)
    pick = torch.randint(0, K, (N,), generator=g)
    pick[torch.rand(N, generator=g) < frac] = heavy
    cn = centers / centers.norm(dim=1, keepdim=True)
    x = x + 0.6 * cn[pick].T
    x = x / x.norm(dim=0, keepdim=True).clamp_min(1e-12)
    return x.reshape(1, D, dh, dw).contiguous()


def make_masks(S: int, Hm: int, Wm: int, seed: int) -> list:
    """S non-empty bool masks [Hm,Wm]: union of 1-3 axis-aligned rectangles, area fraction
    log-uniform in [2e-3, 0.2]."""
    rng = np.random.RandomState(3000 + seed)
    out = []
    for _ in range(S):
        m = np.zeros((Hm, Wm), dtype=bool)
        for _ in range(rng.randint(1, 4)):
            frac = float(np.exp(rng.uniform(np.log(2e-3), np.log(0.2))))
            area = frac * Hm * Wm
            ar = float(np.exp(rng.uniform(-0.7, 0.7)))
            h = int(np.clip(round((area * ar) ** 0.5), 1, Hm))
            w = int(np.clip(round(area / max(h, 1)), 1, Wm))
            y0 = rng.randint(0, Hm - h + 1)
            x0 = rng.randint(0, Wm - w + 1)
            m[y0:y0 + h, x0:x0 + w] = True
        out.append(m)
    return out


def jitter_masks(masks: list, seed: int, px: int = 4) -> list:
    rng = np.random.RandomState(4000 + seed)
    out = []
    for m in masks:
        dy, dx = rng.randint(-px, px + 1, size=2)
        j = np.roll(np.roll(m, dy, axis=0), dx, axis=1)
        if not j.any():
            j = m.copy()
        out.append(j)
    return out


def make_descriptor_bank(Nq: int, Nr: int, D: int, seed: int, planted: int = 1000,
                         cos: float = 0.9, dtype=torch.float32, device="cpu"):
    """Unit-norm query/ref segment descriptors; `planted` queries get a near-duplicate (cos~0.9) in
    the bank so that top-k is not pure noise."""
    g = torch.Generator(device=device).manual_seed(5000 + seed)
    r = torch.randn(Nr, D, generator=g, device=device, dtype=dtype)
    r = r / r.norm(dim=1, keepdim=True)
    q = torch.randn(Nq, D, generator=g, device=device, dtype=dtype)
    q = q / q.norm(dim=1, keepdim=True)
    p = min(planted, Nq, Nr)
    if p > 0:
        qi = torch.randperm(Nq, generator=g, device=device)[:p]
        ri = torch.randperm(Nr, generator=g, device=device)[:p]
        mix = cos * r[ri] + (1 - cos * cos) ** 0.5 * q[qi]
        q[qi] = mix / mix.norm(dim=1, keepdim=True)
    return q.contiguous(), r.contiguous()


def make_structured_bank(n_ref_img: int, n_qry_img: int, segs_per_img: int, D: int, seed: int,
                         noise: float = 0.6, device="cpu"):
    """Place-recognition-shaped bank: query image i re-observes ref image i (each segment descriptor =
    ref segment + noise), so the vote has a right answer and Recall@1 is meaningful."""
    g = torch.Generator(device=device).manual_seed(6000 + seed)
    Nr = n_ref_img * segs_per_img
    r = torch.randn(Nr, D, generator=g, device=device)
    r = r / r.norm(dim=1, keepdim=True)
    Nq = n_qry_img * segs_per_img
    src = (torch.arange(Nq, device=device) % Nr)
    q = r[src] + noise * torch.randn(Nq, D, generator=g, device=device) / D ** 0.5
    q = q / q.norm(dim=1, keepdim=True)
    im_inds_ref = (torch.arange(Nr) // segs_per_img).numpy().astype(np.int64)
    im_inds_qry = (torch.arange(Nq) // segs_per_img).numpy().astype(np.int64)
    return q.contiguous(), r.contiguous(), im_inds_qry, im_inds_ref
