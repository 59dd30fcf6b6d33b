"""TEST INFRASTRUCTURE ONLY -- CPU oracle for the SMEAGOL PWM-scan hot path.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s cpu_baseline / ``--impl reference`` legs may
import this module, and only as the checker / reported baseline.  The product path (``smeagol_b200``) never
imports anything under ``oracle/`` and fails loudly if its CUDA library is missing.

Parity status: PINNED.  This restatement replays every golden vector / known-answer test the reference holds
for the path (tests/test_models.py:38-74, tests/test_scan.py:28-79, tests/test_encode.py:13-85,
tests/test_enrich.py:28-59 + data/real_sites.csv, real_counts.csv, shuf_stats.csv, tests/test_variant.py:47-63)
and additionally is checked against outputs of the reference's own modules run in the build container through
``oracle/ref_harness.py`` (fixtures in tests/golden/, generator tests/golden/make_golden.py).

The arithmetic of the reference lives in TensorFlow's CPU Conv1D (third-party, ``tensorflow>=2.5.0``,
requirements.txt:15, not vendored in /root/reference); its summation order is unspecified.  This file restates
the published algorithm of the reference graph (smeagol/models.py:50-68):

    out[n, p, m] = sum_{k < Wmax} sum_b E[code[n, p+k], b] * K[k, b, m]

with ``code`` zero-padded by Wmax-1 at the END (pad value 0 = 'Z' = zero vector), E the 17x4 embedding table
(smeagol/encode.py:11-29) and K the PWMs left-aligned / tail-zero-padded to Wmax (models.py:70-80).

Two summation orders are provided for the fp32 leg (both are "a fixed order", the reference's is unspecified):
  * order="ref"    : ascending k on the string that the reference scans (for '-' sense: the reverse-complement
                     string produced by encode.integer_encode(rc=True), encode.py:63-64);
  * order="kernel" : ascending k on the FORWARD string with the reverse-complemented table
                     Trc[k][b] = T[W-1-k][3-b] for the '-' sense (what the fused GPU kernel does).  For '+' sense
                     the two orders coincide.
Inner 4-term order for both: ((e0*T0 + e1*T1) + e2*T2) + e3*T3, each product and sum rounded (no FMA).
The fp64 leg (dtype=np.float64) uses the fp32-rounded operands (the reference casts weights and embedding to
fp32: models.py:88-90) accumulated in double; it is the adjudicator for near-threshold sites.
"""
import numpy as np

# ----------------------------------------------------------------------------- alphabet (encode.py:11-41)
CODES = "ZACGTUNWSMKRYBDHV"  # index == integer code
CODE_OF = {c: i for i, c in enumerate(CODES)}
_T = 1.0 / 3.0
EMB64 = np.array(
    [
        [0, 0, 0, 0],  # Z
        [1, 0, 0, 0],  # A
        [0, 1, 0, 0],  # C
        [0, 0, 1, 0],  # G
        [0, 0, 0, 1],  # T
        [0, 0, 0, 1],  # U
        [0.25, 0.25, 0.25, 0.25],  # N
        [0.5, 0, 0, 0.5],  # W
        [0, 0.5, 0.5, 0],  # S
        [0.5, 0.5, 0, 0],  # M
        [0, 0, 0.5, 0.5],  # K
        [0.5, 0, 0.5, 0],  # R
        [0, 0.5, 0, 0.5],  # Y
        [0, _T, _T, _T],  # B
        [_T, 0, _T, _T],  # D
        [_T, _T, 0, _T],  # H
        [_T, _T, _T, 0],  # V
    ],
    dtype=np.float64,
)
EMB32 = EMB64.astype(np.float32)  # models.py:89 loads the table into a float32 Embedding
SENSE_COMPLEMENT = {"+": "-", "-": "+"}  # encode.py:31-34

_DNA_COMP = str.maketrans("ACGTMRWSYKVHDBN", "TGCAKYWSRMBDHVN")
_RNA_COMP = str.maketrans("ACGUMRWSYKVHDBN", "UGCAKYWSRMBDHVN")
# complement on integer codes (used when the '-' strand is evaluated on the forward string)
COMP_CODE = np.array([CODE_OF[c] for c in "ZTGCAANWSKMYRVHDB"], dtype=np.uint8)

_LUT = np.full(256, 255, dtype=np.uint8)
for _c, _i in CODE_OF.items():
    _LUT[ord(_c)] = _i


def reverse_complement_str(seq):
    """Biopython 1.79 ``Seq.reverse_complement`` as used by encode.py:63-64 (see tests/test_encode.py:21-23)."""
    has_u, has_t = ("U" in seq), ("T" in seq)
    if has_u and has_t:
        raise ValueError("Mixed RNA/DNA found")
    return seq.translate(_RNA_COMP if has_u else _DNA_COMP)[::-1]


def integer_encode(seq, rc=False):
    """encode.py:47-68 -> integer codes (uint8 here; the reference stores them as float32).  KeyError on a
    character outside the 17-letter alphabet (encode.py:66)."""
    seq = str(seq)
    if rc:
        seq = reverse_complement_str(seq)
    raw = np.frombuffer(seq.encode("latin-1"), dtype=np.uint8)
    codes = _LUT[raw]
    bad = np.nonzero(codes == 255)[0]
    if bad.size:
        raise KeyError(seq[int(bad[0])])
    return codes


def reverse_complement_matrix(w):
    """True reverse complement of a (W,4) matrix: Trc[k][b] = T[W-1-k][3-b]."""
    return np.ascontiguousarray(np.asarray(w)[::-1, ::-1])


# ----------------------------------------------------------------------------- model constants (models.py:40-46)
def model_constants(weights):
    widths = np.array([w.shape[0] for w in weights])
    max_scores = np.array([np.max(w, axis=1).sum() for w in weights])  # dtype of the supplied weights
    return widths, max_scores, int(widths.max())


def abs_scale(weights):
    """sum_k max_b |fp32(T[k,b])| per PWM: scale of the near-threshold band (fp32 rounding error bound)."""
    return np.array([np.abs(w.astype(np.float32).astype(np.float64)).max(axis=1).sum() for w in weights])


NEAR_REL = 1e-5  # a site is "near-threshold" iff |score - thr| <= NEAR_REL * abs_scale (north_star: 1e-5)


def _term_table(w, dtype):
    """TT[k, code] = ((e0*T0 + e1*T1) + e2*T2) + e3*T3 evaluated in ``dtype`` from fp32-rounded operands."""
    t = np.asarray(w).astype(np.float32).astype(dtype)  # (W,4)
    e = EMB32.astype(dtype)  # (17,4)
    p = e[None, :, :] * t[:, None, :]  # (W,17,4) rounded products
    return ((p[..., 0] + p[..., 1]) + p[..., 2]) + p[..., 3]


def window_scores(codes, w, dtype=np.float32, n_out=None):
    """Scores of all windows of one row: s[p] = sum_{k<W} TT[k, code[p+k]], ascending k, in ``dtype``.
    codes beyond the end count as 'Z' (zero) -- the reference's end padding (models.py:54).  n_out defaults to
    len(codes) (the reference's conv output length, including the overhanging windows that are later trimmed)."""
    codes = np.asarray(codes, dtype=np.uint8)
    L = codes.shape[0]
    W = w.shape[0]
    if n_out is None:
        n_out = L
    tt = _term_table(w, dtype)
    padded = np.concatenate([codes, np.zeros(W, dtype=np.uint8)])
    acc = tt[0][padded[0:n_out]].copy()
    for k in range(1, W):
        acc = acc + tt[k][padded[k:k + n_out]]
    return acc.astype(dtype, copy=False)


def predict(code_rows, weights, dtype=np.float32):
    """PWMModel.predict (models.py:92-100): dense (N, L, C) scores."""
    code_rows = np.atleast_2d(np.asarray(code_rows)).astype(np.uint8)
    N, L = code_rows.shape
    out = np.empty((N, L, len(weights)), dtype=dtype)
    for n in range(N):
        for m, w in enumerate(weights):
            out[n, :, m] = window_scores(code_rows[n], w, dtype)
    return out


def predict_with_threshold(code_rows, weights, frac_threshold, score=False, seq_batch=0):
    """PWMModel.predict_with_threshold (models.py:119-178), fp32 scores compared against fp64 thresholds with a
    strict '>' (models.py:108), trimmed with pos + W <= L (models.py:110).  Row-major (row, pos, motif) order.
    seq_batch only changes batching in the reference (models.py:148-173), never results."""
    assert (frac_threshold >= 0) & (frac_threshold <= 1)
    widths, max_scores, _ = model_constants(weights)
    thresholds = frac_threshold * max_scores
    preds = predict(code_rows, weights, np.float32)
    hit = np.where(preds > thresholds)
    select = hit[1] + widths[hit[2]] <= preds.shape[1]
    hit = tuple(x[select] for x in hit)
    scores = preds[hit] if score else None
    return hit, scores


# ----------------------------------------------------------------------------- fused two-strand scan oracle
def scan_rows(seqs, weights, frac_threshold, rcomp="both", order="ref", adjudicate=True):
    """Scan forward strings the way scan_sequences does for ONE SeqEncoding group (encode.py:137-160 row order):
    rows = [fwd_0..fwd_{n-1}] (+ [rc_0..rc_{n-1}] for rcomp='both'; rc only for 'only').

    Returns dict with int64 arrays ``row, pos, motif`` in (row, pos, motif) order, float32 ``score``
    (fp32 leg in the requested order), float64 ``score64``, bool ``near`` (near-threshold band) and the number of
    band sites whose fp32 and fp64 decisions disagree (``n_flipped``).

    adjudicate=True : hit iff fp64 score > thr for sites in the band, fp32 decision elsewhere (north_star).
    adjudicate=False: pure reference semantics (fp32 score promoted to double > thr, models.py:108).
    All equal-length checks are the caller's business; rows may have different lengths here.
    """
    assert rcomp in ("none", "both", "only")
    widths, max_scores, _ = model_constants(weights)
    thr = frac_threshold * max_scores
    band = NEAR_REL * abs_scale(weights)
    n = len(seqs)
    fwd_codes = [integer_encode(s) for s in seqs]
    rows = []
    if rcomp in ("none", "both"):
        rows += [(i, "+") for i in range(n)]
    if rcomp in ("only", "both"):
        rows += [(i, "-") for i in range(n)]
    out = {k: [] for k in ("row", "pos", "motif", "score", "score64", "near")}
    n_flipped = 0
    for r, (i, strand) in enumerate(rows):
        L = len(seqs[i])
        if strand == "+":
            scan_codes = fwd_codes[i]
        elif order == "ref":
            scan_codes = integer_encode(seqs[i], rc=True)
        else:
            scan_codes = fwd_codes[i]
        per_motif = []
        for m, w in enumerate(weights):
            W = int(widths[m])
            nv = L - W + 1
            if nv <= 0:
                continue
            if strand == "-" and order == "kernel":
                wm = reverse_complement_matrix(w)
            else:
                wm = w
            s32 = window_scores(scan_codes, wm, np.float32, nv)
            s64 = window_scores(scan_codes, wm, np.float64, nv)
            if strand == "-" and order == "kernel":  # forward start p -> RC-string coordinate L - p - W
                s32, s64 = s32[::-1], s64[::-1]
            near = np.abs(s64 - thr[m]) <= band[m]
            hit32 = s32.astype(np.float64) > thr[m]
            hit64 = s64 > thr[m]
            n_flipped += int(np.count_nonzero(near & (hit32 != hit64)))
            hit = np.where(near, hit64, hit32) if adjudicate else hit32
            p = np.nonzero(hit)[0]
            per_motif.append((p, np.full(p.shape, m), s32[p], s64[p], near[p]))
        if per_motif:
            p = np.concatenate([x[0] for x in per_motif])
            mm = np.concatenate([x[1] for x in per_motif])
            o = np.lexsort((mm, p))
            out["row"].append(np.full(p.shape, r))
            out["pos"].append(p[o])
            out["motif"].append(mm[o])
            out["score"].append(np.concatenate([x[2] for x in per_motif])[o])
            out["score64"].append(np.concatenate([x[3] for x in per_motif])[o])
            out["near"].append(np.concatenate([x[4] for x in per_motif])[o])
    res = {}
    for k, dt in (("row", np.int64), ("pos", np.int64), ("motif", np.int64), ("score", np.float32),
                  ("score64", np.float64), ("near", bool)):
        res[k] = np.concatenate(out[k]).astype(dt) if out[k] else np.zeros(0, dtype=dt)
    res["n_flipped"] = n_flipped
    res["rows"] = rows
    return res


def count_table(res, n_rows, n_motifs):
    """Dense per-(row, motif) hit counts from a scan_rows() result."""
    t = np.zeros((n_rows, n_motifs), dtype=np.int64)
    np.add.at(t, (res["row"], res["motif"]), 1)
    return t


# ----------------------------------------------------------------------------- variant scoring (variant.py, scan.py:8-37)
def score_site(w, site_seq, position_wise=False):
    """scan.score_site (scan.py:8-24): fp64 dot of the soft one-hot site with the PWM."""
    codes = integer_encode(site_seq)
    assert codes.shape[0] == w.shape[0]
    pw = np.sum(EMB32[codes] * np.asarray(w), axis=1)  # float32 one-hot x weights (float64)
    return pw if position_wise else np.sum(pw)


def variant_effect_on_site(w, site_seq, rel_pos, alt):
    """variant._variant_effect_on_site (variant.py:29-40) for one (site, variant) pair -> (score, variant_score)
    as fractions of max_score."""
    max_score = np.sum(np.max(w, axis=1))
    ref = score_site(w, site_seq, position_wise=True)
    alt_base = np.sum(np.asarray(w)[rel_pos, :] * EMB64[CODE_OF[alt]])  # one_hot_dict holds python floats
    return np.sum(ref) / max_score, (np.sum(np.delete(ref, rel_pos)) + alt_base) / max_score


def variant_window_scan(seq, positions, alts, weights):
    """Batched restatement for BASELINE config 5: for every SNV (pos, alt) and PWM, over all windows that cover
    ``pos`` on both strands and lie inside the sequence, the maximum reference score and the maximum alternate
    score (fp64 from fp32-rounded weights).  Returns (ref_max, alt_max) of shape (n_snv, n_motifs); -inf where
    no window fits."""
    codes = integer_encode(seq)
    L = codes.shape[0]
    ns, M = len(positions), len(weights)
    ref_max = np.full((ns, M), -np.inf)
    alt_max = np.full((ns, M), -np.inf)
    for m, w in enumerate(weights):
        W = w.shape[0]
        for strand in (0, 1):
            tt = _term_table(reverse_complement_matrix(w) if strand else w, np.float64)
            for v in range(ns):
                pos = int(positions[v])
                a = CODE_OF[alts[v]]
                for s in range(max(0, pos - W + 1), min(pos, L - W) + 1):
                    win = codes[s:s + W]
                    ref = 0.0
                    alt = 0.0
                    for k in range(W):
                        ref = ref + tt[k][win[k]]
                        alt = alt + tt[k][a if s + k == pos else win[k]]
                    ref_max[v, m] = max(ref_max[v, m], ref)
                    alt_max[v, m] = max(alt_max[v, m], alt)
    return ref_max, alt_max
