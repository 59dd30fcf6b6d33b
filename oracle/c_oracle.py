"""TEST INFRASTRUCTURE ONLY -- ctypes wrapper of oracle/scan_oracle.c (built by `make -C oracle`, run by
__graft_entry__.build()).  Same semantics as oracle/scan_oracle.py::scan_rows(order="kernel", adjudicate=True),
fast enough (OpenMP) for benchmark-sized parity checks."""
import ctypes
import os
import subprocess

import numpy as np

from . import scan_oracle as so

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = os.path.join(_HERE, "_build", "liboracle.so")
_lib = None


def load():
    global _lib
    if _lib is None:
        if not os.path.exists(_LIB):
            subprocess.check_call(["make", "-C", _HERE])
        _lib = ctypes.CDLL(_LIB)
        _lib.orc_scan.restype = ctypes.c_int
    return _lib


def bits(n):
    return max(1, int(n - 1).bit_length()) if n > 1 else 1


def scan(code_rows, weights, frac_threshold, strands=3, out_rows=None, n_out_rows=None, want_hits=False, hit_cap=0):
    """code_rows: list of uint8 code arrays (forward rows).  Returns dict(counts (n_rows, M, 2), n_hits, n_near,
    key_sum, key_xor, n_flipped[, keys, scores sorted])."""
    lib = load()
    n_rows, M = len(code_rows), len(weights)
    lens = np.array([len(c) for c in code_rows], dtype=np.int64)
    row_off = np.concatenate([[0], np.cumsum(lens)[:-1]]).astype(np.int64)
    codes = np.ascontiguousarray(np.concatenate(code_rows).astype(np.uint8))
    w32 = [np.ascontiguousarray(np.asarray(w), dtype=np.float32) for w in weights]
    widths = np.array([w.shape[0] for w in w32], dtype=np.int32)
    tab_off = np.concatenate([[0], np.cumsum(widths)[:-1]]).astype(np.int64)
    tab = np.ascontiguousarray(np.concatenate([w.reshape(-1) for w in w32]))
    _, max_scores, _ = so.model_constants(weights)
    thr = np.ascontiguousarray(np.asarray(frac_threshold * max_scores, dtype=np.float64))
    band = np.ascontiguousarray(so.NEAR_REL * so.abs_scale(weights))
    if out_rows is None:
        out_rows = np.stack([np.arange(n_rows), n_rows + np.arange(n_rows)], axis=1)
        n_out_rows = 2 * n_rows
    out_rows = np.ascontiguousarray(out_rows, dtype=np.int32)
    pos_bits, motif_bits = bits(int(lens.max()) + 1), bits(M)
    counts = np.zeros((n_rows, M, 2), dtype=np.int64)
    totals = np.zeros(8, dtype=np.uint64)
    cap = int(hit_cap) if want_hits else 0
    keys = np.zeros(max(cap, 1), dtype=np.uint64)
    scores = np.zeros(max(cap, 1), dtype=np.float32)
    P = lambda a: a.ctypes.data_as(ctypes.c_void_p)  # noqa: E731
    rc = lib.orc_scan(P(codes), P(row_off), P(lens), n_rows, P(tab), P(widths), P(tab_off), M, P(thr), P(band),
                      int(strands), P(out_rows), pos_bits, motif_bits, P(counts), P(totals), P(keys), P(scores),
                      ctypes.c_int64(cap))
    if rc != 0:
        raise RuntimeError("oracle hit buffer overflow")
    res = {"counts": counts, "n_hits": int(totals[0]), "n_near": int(totals[1]), "key_sum": int(totals[2]),
           "key_xor": int(totals[3]), "n_flipped": int(totals[4]), "pos_bits": pos_bits, "motif_bits": motif_bits}
    if want_hits:
        n = int(totals[0])
        o = np.argsort(keys[:n], kind="stable")
        res["keys"], res["scores"] = keys[:n][o], scores[:n][o]
    return res


def variant_windows(code_row, weights, snv_pos, snv_alt_codes):
    """Config c5: max reference / alternate window score per (SNV, PWM) over the windows covering the SNV, both strands
    (fp32, kernel summation order).  code_row: uint8 integer codes of ONE sequence; snv_alt_codes: integer codes 0..16."""
    lib = load()
    w32 = [np.ascontiguousarray(np.asarray(w), dtype=np.float32) for w in weights]
    widths = np.array([w.shape[0] for w in w32], dtype=np.int32)
    tab_off = np.concatenate([[0], np.cumsum(widths)[:-1]]).astype(np.int64)
    tab = np.ascontiguousarray(np.concatenate([w.reshape(-1) for w in w32]))
    codes = np.ascontiguousarray(code_row, dtype=np.uint8)
    pos = np.ascontiguousarray(snv_pos, dtype=np.int64)
    alt = np.ascontiguousarray(snv_alt_codes, dtype=np.uint8)
    n, M = len(pos), len(w32)
    ref_max = np.empty((n, M), dtype=np.float32)
    alt_max = np.empty((n, M), dtype=np.float32)
    P = lambda a: a.ctypes.data_as(ctypes.c_void_p)  # noqa: E731
    lib.orc_variant_windows.restype = None
    lib.orc_variant_windows(P(codes), ctypes.c_int64(len(codes)), P(tab), P(widths), P(tab_off), M, P(pos), P(alt),
                            ctypes.c_int64(n), P(ref_max), P(alt_max))
    return ref_max, alt_max
