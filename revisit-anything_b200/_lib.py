"""ctypes binding of libsegvlad.so (include/segvlad.h).  There is NO fallback: if the shared library
is missing or a call fails, this raises -- the product path never routes through CPU code."""
from __future__ import annotations

import ctypes as C
import os

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "libsegvlad.so")

OUT_F64, OUT_F32 = 0, 1
TOKENS_DN, TOKENS_ND, TOKENS_PRENORMALIZED = 0, 1, 2

_p = C.c_void_p
_SIGS = {
    "segvlad_version": (C.c_int, []),
    "segvlad_last_error": (C.c_char_p, []),
    "segvlad_launch_count": (C.c_uint64, []),
    "segvlad_profile_enable": (None, [C.c_int]),
    "segvlad_profile_read": (C.c_int, [C.c_int, _p, _p]),
    "segvlad_profile_reset": (None, []),
    "segvlad_aggregate_workspace_bytes": (C.c_size_t, [C.c_int] * 5),
    "segvlad_aggregate_batch": (C.c_int, [_p, C.c_int, C.c_int, C.c_int, C.c_int, _p, C.c_int, _p, _p, _p, _p,
                                          C.c_int, _p, _p, C.c_size_t, _p]),
    "segvlad_aggregate_batch_pca": (C.c_int, [_p, C.c_int, C.c_int, C.c_int, C.c_int, _p, C.c_int, _p, _p, _p, _p, _p, _p, _p,
                                              C.c_size_t, _p]),
    "segvlad_aggregate_residuals": (C.c_int, [_p, _p, C.c_int, C.c_int, C.c_int, C.c_int, _p, _p, _p, _p, C.c_int, _p,
                                              C.c_size_t, _p]),
    "segvlad_mask_to_membership": (C.c_int, [_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, _p, _p]),
    "segvlad_mask_centroids": (C.c_int, [_p, C.c_int, C.c_int, C.c_int, _p, _p]),
    "segvlad_bank_bytes": (C.c_size_t, [C.c_int, C.c_int]),
    "segvlad_bank_prepare": (C.c_int, [_p, C.c_int, C.c_int, _p, _p]),
    "segvlad_bank_prepare_view": (C.c_int, [_p, C.c_int, C.c_int, _p, _p]),
    "segvlad_bank_prepare_f64": (C.c_int, [_p, C.c_int, C.c_int, C.c_int, _p, _p]),
    "segvlad_knn_workspace_bytes": (C.c_size_t, [C.c_int] * 4),
    "segvlad_knn": (C.c_int, [_p, C.c_int, _p, C.c_int, C.c_int64, C.c_int, C.c_int, _p, _p, _p, C.c_size_t, _p]),
    "segvlad_knn_async": (C.c_int, [_p, C.c_int, _p, C.c_int, C.c_int64, C.c_int, C.c_int, C.c_int, _p, _p, _p, _p, _p,
                                    C.c_size_t, _p]),
    "segvlad_knn_from_host": (C.c_int, [_p, C.c_int, _p, C.c_int, C.c_int64, C.c_int, C.c_int, _p, _p, _p, _p, _p,
                                        C.c_size_t, _p, _p]),
    "segvlad_knn_simt": (C.c_int, [_p, C.c_int, _p, C.c_int, C.c_int64, C.c_int, C.c_int, _p, _p, _p, C.c_size_t, _p]),
    "segvlad_knn_debug_approx": (C.c_int, [_p, C.c_int, _p, C.c_int, C.c_int, _p, _p, _p, C.c_size_t, _p]),
    "segvlad_merge_topk": (C.c_int, [_p, _p, C.c_int, C.c_int, C.c_int, _p, _p, _p]),
    "segvlad_merge_topk_packed": (C.c_int, [_p, C.c_int, C.c_size_t, C.c_int, C.c_int, _p, _p, _p, _p]),
    "segvlad_vote_workspace_bytes": (C.c_size_t, [C.c_int] * 4),
    "segvlad_vote": (C.c_int, [_p, _p, C.c_int, C.c_int, C.c_int, C.c_int, _p, _p, C.c_int, C.c_int, _p, C.c_int, C.c_int,
                               C.c_int, _p, _p, _p, _p, _p, _p, C.c_size_t, _p]),
    "segvlad_pca_workspace_bytes": (C.c_size_t, [C.c_int] * 3),
    "segvlad_pca_project": (C.c_int, [_p, C.c_int, C.c_int, _p, _p, _p, C.c_int, C.c_int, _p, _p, C.c_size_t, _p]),
    "segvlad_pca_tc_supported": (C.c_int, [C.c_int, C.c_int]),
    "segvlad_pca_planes_bytes": (C.c_size_t, [C.c_int, C.c_int]),
    "segvlad_pca_prepare_planes": (C.c_int, [_p, C.c_int, C.c_int, _p, _p]),
    "segvlad_pca_tc_workspace_bytes": (C.c_size_t, [C.c_int] * 3),
    "segvlad_pca_project_tc": (C.c_int, [_p, C.c_int, C.c_int, _p, _p, _p, C.c_int, C.c_int, _p, _p, C.c_size_t, _p]),
    "segvlad_pca_project_planes": (C.c_int, [_p, C.c_int, C.c_int, _p, _p, C.c_int, C.c_int, _p, _p, C.c_size_t, _p]),
    "segvlad_netvlad_workspace_bytes": (C.c_size_t, [C.c_int] * 4),
    "segvlad_netvlad_antiburst": (C.c_int, [_p, C.c_int, C.c_int, C.c_int, _p, _p, C.c_int, C.c_float, C.c_float,
                                            C.c_float, _p, _p, C.c_size_t, _p]),
}
EXPORTED = tuple(_SIGS)

_lib = None


class SegVladError(RuntimeError):
    pass


def lib():
    """Load (once) and return the ctypes handle; raises if the library was not built."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise SegVladError(
                f"{LIB_PATH} not found: build it with `python -m revisit_anything_b200.build` "
                "(there is no CPU / PyTorch fallback for the SegVLAD hot path)")
        h = C.CDLL(LIB_PATH)
        for name, (res, args) in _SIGS.items():
            if not hasattr(h, name):
                continue  # optional entry points of later milestones are checked by tests/test_abi.py
            fn = getattr(h, name)
            fn.restype, fn.argtypes = res, args
        _lib = h
    return _lib


def check(rc: int, what: str):
    if rc != 0:
        msg = lib().segvlad_last_error().decode("utf-8", "replace")
        exc = ValueError if rc == -1 else SegVladError
        raise exc(f"{what} failed (code {rc}): {msg}")
