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


def load_netvlad_module():
    """The UNMODIFIED VLAD-BuFF/models/aggregators/aggregation.py (NetVLAD + anti-burst, row a9), imported on the CPU with
    its absent third-party imports (faiss, tqdm if missing) stubbed; used to pin oracle.netvlad_antiburst."""
    import importlib.util
    if not os.path.isfile(os.path.join(REF_ROOT, "VLAD-BuFF", "models", "aggregators", "aggregation.py")):
        raise RuntimeError("reference VLAD-BuFF not mounted")
    for name in ("faiss", "tqdm"):
        try:
            __import__(name)
        except Exception:
            sys.modules[name] = MagicMock()
    spec = importlib.util.spec_from_file_location(
        "ref_vladbuff_aggregation", os.path.join(REF_ROOT, "VLAD-BuFF", "models", "aggregators", "aggregation.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def netvlad_reference_forward(mod, x_bdhw, centroids, conv_weight, ab_params=(8.0, 7.0, 1.0), for_loop_alt=False):
    """Runs the reference's NetVLAD.forward (aggregation.py:266-361) with antiburst=True and the evaluation defaults
    (eval.py:384-386: ab_w, ab_b, ab_p = 8, 7, 1; ab_relu / ab_inv / ab_soft off; no nv_pca), centroids and the 1x1-conv weight
    set the way init_params does (:245-256).  `for_loop_alt` selects the broadcast branch (:346-349) instead of the per-cluster
    loop (:351-358); both are the same arithmetic."""
    import types

    import torch
    K, D = centroids.shape
    args = types.SimpleNamespace(
        expName="pin", nv_pca=None, nv_pca_alt=False, nv_pca_alt_mlp=False, antiburst=True, ab_w=float(ab_params[0]),
        ab_b=float(ab_params[1]), ab_p=float(ab_params[2]), ab_fixed=True, ab_gen=0, ab_t=None, ab_kp=None, ab_testOnly=False,
        ab_wOnly=False, ab_relu=False, ab_inv=False, ab_soft=False, forLoopAlt=bool(for_loop_alt), storeSAB=False)
    net = mod.NetVLAD(clusters_num=K, dim=D, normalize_input=True, work_with_tokens=False, args=args)
    with torch.no_grad():
        net.centroids = torch.nn.Parameter(centroids.clone().float())
        net.conv.weight = torch.nn.Parameter(conv_weight.clone().float().reshape(K, D, 1, 1))
        net.conv.bias = None
        net.eval()
        return net(x_bdhw.float())
