"""PWM similarity -- API mirror of the similarity part of smeagol/matrices.py.

pairwise_ncorrs (matrices.py:396-433) is the library's other super-linear numeric loop (O(M^2) Python calls of ncorr,
each a loop over alignments with np.corrcoef): here ONE device kernel scores every pair (smg_pairwise_ncorrs, fp64).
The single-pair helpers keep the reference's names and semantics on the host, including its reverse_complement as
written (matrices.py:238-243: the columns are swapped, the rows keep their order).  PWM curation, clustering and
file readers of matrices.py are out of scope.
"""
import ctypes

import numpy as np
import torch

from . import _lib
from . import device as dev


def reverse_complement(mat):
    """matrices.py:224-243, as written there: `rc = mat[::-1, :].copy()` followed by `rc[:, b] = mat[:, 3 - b]` for every
    column, i.e. the input with its columns swapped (A<->T, C<->G) and the rows in their ORIGINAL order."""
    mat = np.asarray(mat)
    rc = mat[::-1, :].copy()
    rc[:, 0] = mat[:, 3]
    rc[:, 1] = mat[:, 2]
    rc[:, 2] = mat[:, 1]
    rc[:, 3] = mat[:, 0]
    return rc


def matrix_correlation(X, Y):
    """Pearson correlation of two equal-sized matrices unrolled into 1-D arrays (matrices.py:263-287)."""
    if X.shape != Y.shape:
        raise ValueError("Inputs do not have equal shapes.")
    return np.corrcoef(np.concatenate(X), np.concatenate(Y))[0, 1]


def order_pms(X, Y):
    """The narrower matrix first (matrices.py:290-303)."""
    if len(Y) < len(X):
        X, Y = Y, X
    return X, Y


def align_pms(X, Y, min_overlap=None):
    """All overlaps of two position matrices (matrices.py:306-353): (X_start, X_end, Y_start, Y_end) tuples."""
    X, Y = order_pms(X, Y)
    Lx, Ly = len(X), len(Y)
    if min_overlap is not None:
        assert type(min_overlap) == int
        assert (min_overlap > 0) & (min_overlap <= Lx)
    else:
        min_overlap = min(3, Lx)
    aln_starts = range(min_overlap - Lx, Ly - min_overlap + 1)
    return [(max(-i, 0), min(Lx, Ly - i), max(i, 0), min(Ly, i + Lx)) for i in aln_starts]


def ncorr(X, Y, min_overlap=None):
    """Normalised Pearson correlation of two position matrices (matrices.py:356-395; doi 10.1093/nar/gkx314).  As in the
    reference, min_overlap is accepted and NOT passed on to align_pms."""
    X, Y = order_pms(np.asarray(X), np.asarray(Y))
    vals = []
    for xs, xe, ys, ye in align_pms(X, Y, min_overlap=None):
        w = ye - ys
        vals.append(matrix_correlation(X[xs:xe, :], Y[ys:ye, :]) * w / (len(X) + len(Y) - w))
    return max(vals)


def pairwise_ncorrs(mats):
    """All pairwise normalised Pearson correlations between position matrices (matrices.py:396-433), on the device: one
    thread per pair, every alignment and the four orientation combinations in fp64.  Returns an (n, n) array with ones on
    the diagonal."""
    mats = [np.ascontiguousarray(np.asarray(m), dtype=np.float64) for m in mats]
    n = len(mats)
    if n == 0:
        return np.zeros((0, 0))
    for m in mats:
        if m.ndim != 2 or m.shape[1] != 4 or m.shape[0] < 1:
            raise ValueError("position matrices must have shape (W, 4)")
    d = dev._require_cuda()
    widths = np.array([m.shape[0] for m in mats], dtype=np.int32)
    offs = np.concatenate([[0], np.cumsum(widths)[:-1]]).astype(np.int64)
    flat = np.concatenate([m.reshape(-1) for m in mats])
    d_m, d_w, d_o = torch.from_numpy(flat).to(d), torch.from_numpy(widths).to(d), torch.from_numpy(offs).to(d)
    out = torch.zeros((n, n), dtype=torch.float64, device=d)
    with torch.cuda.device(d):
        _lib.check(_lib.load().smg_pairwise_ncorrs(dev._ptr(d_m), dev._ptr(d_w), dev._ptr(d_o), n, dev._ptr(out), dev._stream()),
                   "smg_pairwise_ncorrs")
    return out.cpu().numpy()
