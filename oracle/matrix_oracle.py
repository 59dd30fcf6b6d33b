"""TEST INFRASTRUCTURE ONLY -- CPU restatement of the SURVEY 8(f) rank-4 rows: PWM similarity (smeagol/matrices.py:
224-243 reverse_complement, :263-287 matrix_correlation, :290-353 order_pms / align_pms, :356-395 ncorr, :396-433
pairwise_ncorrs) and window counting (smeagol/utils.py:135-180 tiling windows, smeagol/scan.py:381-423
_count_in_window / count_in_sliding_windows).  Pinned against recorded outputs of the reference by
tests/test_oracle_golden.py (tests/golden/reference_golden_f4.json).  Never imported by the product package.
"""
import numpy as np
import pandas as pd


def reverse_complement(mat):
    """matrices.py:224-243 AS WRITTEN: the row-reversed copy has every column overwritten from the UN-reversed input
    (rc[:, 0] = mat[:, 3], ...), so the result is the input with its columns swapped (A<->T, C<->G) and the rows in
    their original order.  pairwise_ncorrs is pinned to this behaviour (recorded outputs of the unmodified module); the
    reference's own expectation in tests/test_matrices.py:199-213 (0.46869686 for PWMs x, y) is not what its code
    returns here (0.26298287)."""
    return np.asarray(mat)[:, ::-1].copy()


def pearson(x, y):
    """np.corrcoef(x, y)[0, 1] (matrices.py:284): covariance over the product of the standard deviations, clipped."""
    x = np.asarray(x, dtype=np.float64).ravel()
    y = np.asarray(y, dtype=np.float64).ravel()
    dx, dy = x - x.mean(), y - y.mean()
    n1 = len(x) - 1
    cxy, cxx, cyy = np.dot(dx, dy) / n1, np.dot(dx, dx) / n1, np.dot(dy, dy) / n1
    with np.errstate(invalid="ignore", divide="ignore"):
        c = cxy / np.sqrt(cxx) / np.sqrt(cyy)
    return float(np.clip(c, -1.0, 1.0)) if c == c else float("nan")


def _pymax(vals):
    """Python's max(): keeps the first item unless a later one compares greater (NaN never does)."""
    best = vals[0]
    for v in vals[1:]:
        if v > best:
            best = v
    return best


def ncorr(X, Y):
    """matrices.py:356-395.  The reference passes min_overlap=None on to align_pms whatever its own argument was."""
    X, Y = np.asarray(X, dtype=np.float64), np.asarray(Y, dtype=np.float64)
    if len(Y) < len(X):
        X, Y = Y, X
    Lx, Ly = len(X), len(Y)
    mo = min(3, Lx)
    vals = []
    for i in range(mo - Lx, Ly - mo + 1):
        xs, xe = max(-i, 0), min(Lx, Ly - i)
        ys, ye = max(i, 0), min(Ly, i + Lx)
        w = ye - ys
        vals.append(pearson(X[xs:xe], Y[ys:ye]) * w / (Lx + Ly - w))
    return _pymax(vals)


def pairwise_ncorrs(mats):
    n = len(mats)
    sims = np.zeros((n, n))
    for i in range(n):
        for j in range(i + 1, n):
            a, b = np.asarray(mats[i]), np.asarray(mats[j])
            sims[i, j] = sims[j, i] = _pymax([ncorr(a, b), ncorr(reverse_complement(a), b), ncorr(a, reverse_complement(b)),
                                             ncorr(reverse_complement(a), reverse_complement(b))])
        sims[i, i] = 1
    return sims


def tiling_windows(ids, lens, width, shift=None):
    shift = width if shift is None else shift
    frames = []
    for rid, L in zip(ids, lens):
        starts = list(range(0, L, shift))
        frames.append(pd.DataFrame({"id": [rid] * len(starts), "start": starts, "end": [min(s + width, L) for s in starts]}))
    return pd.concat(frames) if len(frames) > 1 else frames[0]


def count_in_sliding_windows(sites, ids, lens, matrix_id, width, shift=None):
    windows = tiling_windows(ids, lens, width, shift).reset_index(drop=True)
    counts = []
    for _, w in windows.iterrows():
        sel = (sites.id == w.id) & (sites.start >= w.start) & (sites.start < w.end)
        if matrix_id is not None:
            sel &= sites.Matrix_id == matrix_id
        counts.append(int(sel.sum()))
    windows["count"] = counts
    return windows
