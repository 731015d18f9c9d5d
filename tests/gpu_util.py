import numpy as np


def assert_knn_close(d2, idx, d2_ref64, idx_ref64, tol=2e-6, k_check=None):
    """Compare a kNN result with an fp64 reference.  Distances must agree to `tol` (absolute, on d2 of
    O(1) magnitude <=> 1e-5 relative with margin); indices must agree wherever the fp64 distance gap to
    the neighbouring ranks exceeds 2*tol (inside a near-tie group any order is a legitimate fp32 result,
    but the returned index must still carry (almost) the same distance)."""
    d2 = np.asarray(d2, dtype=np.float64)
    k = d2.shape[1] if k_check is None else k_check
    ref = d2_ref64[:, :k]
    np.testing.assert_allclose(d2[:, :k], ref, rtol=1e-5, atol=tol)
    full = d2_ref64                      # gaps are taken on the longer reference list so that the k-th
    gap_prev = np.concatenate([np.full((full.shape[0], 1), np.inf), np.diff(full, axis=1)], axis=1)[:, :k]
    gap_next = np.concatenate([np.diff(full, axis=1), np.full((full.shape[0], 1), np.inf)], axis=1)[:, :k]
    if full.shape[1] == k:               # entry also knows its next neighbour (callers pass k + 8 columns)
        gap_next[:, -1] = 0.0
    clear = (gap_prev > 2 * tol) & (gap_next > 2 * tol)
    bad = clear & (np.asarray(idx)[:, :k] != idx_ref64[:, :k])
    assert not bad.any(), f"{int(bad.sum())} index mismatches outside near-ties, first at {np.argwhere(bad)[:5]}"
    return float(clear.mean())


def assert_desc_close(got, want, rtol=1e-5):
    """VLAD descriptor parity: elementwise 1e-5 relative, with an absolute floor of 1e-6 x the row's largest
    element for entries that are ~0 by cancellation."""
    got, want = np.asarray(got, dtype=np.float64), np.asarray(want, dtype=np.float64)
    floor = 1e-6 * np.abs(want).max(axis=-1, keepdims=True)
    err = np.abs(got - want)
    bad = err > rtol * np.abs(want) + floor
    assert not bad.any(), f"{int(bad.sum())} elements out of tolerance, max abs err {err.max():.3e}"
