"""Import the UNMODIFIED reference `func_vpr` in the build container (test infrastructure only).

/root/reference does not exist on the GPU box; callers must use `available()` and skip otherwise.
The reference imports a dozen packages that are absent offline (faiss, h5py, natsort, ...); none of
them is touched by the functions on the hot path, so they are stubbed with MagicMock
(SURVEY.md Appendix C).  Importing has side effects (chdir, seed_everything(42) at
utilities.py:1011) which are undone here.
"""
from __future__ import annotations

import os
import sys
from unittest.mock import MagicMock

REF_ROOT = "/root/reference"
_STUBS = [
    "faiss", "faiss.contrib", "faiss.contrib.torch_utils", "h5py", "natsort", "matplotlib",
    "matplotlib.pyplot", "utm", "tkinter", "fast_pytorch_kmeans", "pytorch_lightning",
    "pytorch_metric_learning", "pytorch_metric_learning.losses", "pytorch_metric_learning.miners",
    "pytorch_metric_learning.distances", "wandb", "prettytable", "networkx",
]
_cached = None


def available() -> bool:
    return os.path.isfile(os.path.join(REF_ROOT, "func_vpr.py"))


def load():
    """Returns the reference `func_vpr` module (cached)."""
    global _cached
    if _cached is not None:
        return _cached
    if not available():
        raise RuntimeError("reference not mounted at /root/reference")
    import torch

    for name in _STUBS:
        try:
            __import__(name)
        except Exception:
            sys.modules[name] = MagicMock()
    pl = sys.modules["pytorch_lightning"]
    if isinstance(pl, MagicMock):
        pl.LightningModule = torch.nn.Module
        pl.LightningDataModule = object
    cwd = os.getcwd()
    rng_t = torch.get_rng_state()
    import random

    import numpy as np
    rng_n, rng_p = np.random.get_state(), random.getstate()
    added = [REF_ROOT, os.path.join(REF_ROOT, "sam")]
    sys.path[:0] = added
    # the repo has its own drop-in module called func_vpr inside the package, never top-level,
    # so the top-level name resolves to the reference here.
    saved_stdout = sys.stdout
    try:
        os.chdir(REF_ROOT)
        sys.stdout = open(os.devnull, "w")
        import func_vpr as ref_func_vpr  # noqa
    finally:
        sys.stdout.close() if sys.stdout is not saved_stdout else None
        sys.stdout = saved_stdout
        os.chdir(cwd)
        torch.set_rng_state(rng_t)
        np.random.set_state(rng_n)
        random.setstate(rng_p)
        torch.backends.cudnn.deterministic = False
    _cached = ref_func_vpr
    return ref_func_vpr


def vlad_single_cpu(ref, query_descs, c_centers, masks, adj_mat=None):
    """The reference's vlad_single body (func_vpr.py:1145-1175) cannot run without a GPU because of
    its hard-coded .to('cuda'); this drives the reference's OWN vlad_matmuls_per_cluster with
    device='cpu', fed by the two label/residual lines executed with torch ops on CPU
    (BASELINE.md section 2 describes exactly this arrangement)."""
    import torch
    import torch.nn.functional as F

    cn = F.normalize(c_centers, dim=1)
    labels = torch.argmax(query_descs @ cn.T, dim=1)
    res = query_descs - c_centers[labels]
    adj = None if adj_mat is None else adj_mat.double()
    out, _ = ref.vlad_matmuls_per_cluster(c_centers.shape[0], masks.double(), res.double(), labels,
                                          adjMat=adj, device="cpu")
    return out, labels
