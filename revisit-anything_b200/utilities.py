"""Mirror of the one class of the reference's `utilities.py` that the AnyLoc baseline of place_rec_main.py uses:
`VLAD` (utilities.py:624-1000), hard-assignment mode.  `generate` / `generate_multi` run the same sm_100a aggregation
kernels as SegVLAD with ONE all-ones segment per image; there is no CPU fallback.

    vlad = VLAD(32, desc_dim=None, dist_mode="cosine", vlad_mode="hard", cache_dir=".../c32")
    vlad.fit(None)                       # loads <cache_dir>/c_centers.pt like the reference (utilities.py:766-775)
    gd = vlad.generate(tokens_nd)        # [K*D] fp32 CPU tensor, utilities.py:827-905

The vocabulary is never trained here (`fit` with descriptors needs fast_pytorch_kmeans, which the reference uses
offline in vlad_c_centers_pt_gen.py; out of the hot path): `fit(train_descs)` without a cached vocabulary raises.
"""
from __future__ import annotations

import os
from typing import List, Optional, Sequence, Union

import numpy as np
import torch

from . import engine
from ._lib import TOKENS_ND, TOKENS_PRENORMALIZED


class VLAD:
    def __init__(self, num_clusters: int, desc_dim: Optional[int] = None, intra_norm: bool = True,
                 norm_descs: bool = True, dist_mode: str = "cosine", vlad_mode: str = "hard", soft_temp: float = 1.0,
                 cache_dir: Optional[str] = None) -> None:
        self.num_clusters = num_clusters
        self.desc_dim = desc_dim
        self.intra_norm = intra_norm
        self.norm_descs = norm_descs
        self.mode = dist_mode
        self.vlad_mode = str(vlad_mode).lower()
        assert self.vlad_mode in ["soft", "hard"]
        self.soft_temp = soft_temp
        self.c_centers: Optional[torch.Tensor] = None
        self.kmeans = None
        self.cache_dir = None if cache_dir is None else os.path.abspath(os.path.expanduser(cache_dir))
        self._centers_dev = None

    # -- vocabulary -------------------------------------------------------------------------------------------------
    def can_use_cache_vlad(self) -> bool:
        return self.cache_dir is not None and os.path.exists(f"{self.cache_dir}/c_centers.pt")

    def fit(self, train_descs: Union[np.ndarray, torch.Tensor, None]):
        if self.can_use_cache_vlad():
            self.set_centers(torch.load(f"{self.cache_dir}/c_centers.pt", map_location="cpu"))
            return
        if train_descs is None:
            raise ValueError("No training descriptors given")
        raise NotImplementedError("segvlad: k-means vocabulary training is offline tooling in the reference "
                                  "(vlad_c_centers_pt_gen.py); provide <cache_dir>/c_centers.pt or call set_centers()")

    def set_centers(self, c_centers: torch.Tensor):
        assert c_centers.ndim == 2 and c_centers.shape[0] == self.num_clusters, "c_centers must be [num_clusters, D]"
        self.c_centers = c_centers.detach().to("cpu", torch.float32)
        self._centers_dev = None
        if self.desc_dim is None:
            self.desc_dim = int(c_centers.shape[1])

    # -- aggregation ------------------------------------------------------------------------------------------------
    def _check(self):
        if self.c_centers is None:
            raise ValueError("VLAD: call fit()/set_centers() first")
        if self.vlad_mode != "hard" or self.mode != "cosine" or not self.intra_norm:
            raise NotImplementedError("segvlad: only the configuration the drivers use is built "
                                      "(dist_mode='cosine', vlad_mode='hard', intra_norm=True; place_rec_main.py:156)")
        if not torch.cuda.is_available():
            raise RuntimeError("segvlad: no CUDA device (no CPU fallback)")

    def _centers(self, dev):
        if self._centers_dev is None or self._centers_dev.device != dev:
            self._centers_dev = self.c_centers.to(dev)
        return self._centers_dev

    def generate_multi(self, multi_query: Union[np.ndarray, torch.Tensor, list], cache_ids=None,
                       device_out: bool = False):
        """utilities.py:907-926.  A [B, N, D] tensor/array goes through ONE batched launch; a list of [N_i, D] items
        is grouped by N_i.  Returns a [B, K*D] fp32 tensor (or a list for ragged input, like the reference)."""
        self._check()
        dev = torch.device("cuda", torch.cuda.current_device())
        ragged = isinstance(multi_query, list)
        items = [torch.as_tensor(q, dtype=torch.float32) for q in multi_query] if ragged else None
        if not ragged:
            t = torch.as_tensor(multi_query, dtype=torch.float32)
            out = self._run(t.to(dev), dev)
            return out if device_out else out.cpu()
        res: List[Optional[torch.Tensor]] = [None] * len(items)
        by_n = {}
        for i, it in enumerate(items):
            by_n.setdefault(int(it.shape[0]), []).append(i)
        for n, ids in by_n.items():
            out = self._run(torch.stack([items[i] for i in ids]).to(dev), dev)
            for j, i in enumerate(ids):
                res[i] = out[j] if device_out else out[j].cpu()
        return res

    def _run(self, tok_bnd: torch.Tensor, dev) -> torch.Tensor:
        B, N, D = tok_bnd.shape
        if D != self.c_centers.shape[1]:
            raise ValueError("descriptor dim does not match the vocabulary")
        words = (N + 31) // 32
        row = np.full(words, -1, dtype=np.int32)            # all tokens are members of the single segment
        if N % 32:
            row[-1] = np.int32((1 << (N % 32)) - 1)
        bits = torch.from_numpy(np.tile(row, (B, 1))).to(dev)
        layout = TOKENS_ND | (0 if self.norm_descs else TOKENS_PRENORMALIZED)
        return engine.aggregate_batch(tok_bnd, N, D, layout, self._centers(dev), bits, [1] * B, None,
                                      out_dtype=torch.float32)

    def generate(self, query_descs: Union[np.ndarray, torch.Tensor], cache_id: Optional[str] = None) -> torch.Tensor:
        """utilities.py:827-905: [n_q, D] descriptors of one image -> normalised VLAD [K*D] (fp32, CPU)."""
        q = torch.as_tensor(query_descs, dtype=torch.float32)
        return self.generate_multi(q[None])[0]
