"""CPU check of the bound the seed engine rests on (csrc/kmer.h; the CUDA path itself is compared with the
oracle hit for hit in tests/test_gpu_parity.py::test_kmer_index_engine_bit_identical).

A window of a PWM of width W scores  seed + rest  with  rest <= sum_{k outside the seed} max(0, max_b T[k][b]),  so every
window whose fp32 score exceeds the threshold has a seed whose fp32 score exceeds  threshold - rest_max - margin.
The test restates the host-side seed choice of api.cu::seed_set (z-score of the seed threshold under uniform bases) and
checks the superset property, with the kernel's fp32 summation order, on random PWMs and sequences for both strands."""
import numpy as np

S_VALUES = (11, 12)


def choose_seed(T, S):
    """Offset of the most selective S consecutive rows of the (W, 4) strand table T, and the bound of the rest."""
    W = T.shape[0]
    mx = np.maximum(T.max(axis=1), 0.0)
    mu, var = T.mean(axis=1), np.maximum(T.var(axis=1), 1e-12)
    thr_proxy = 0.7 * T.max(axis=1).sum()
    best = max(range(W - S + 1),
               key=lambda o: (thr_proxy - (mx.sum() - mx[o:o + S].sum()) - mu[o:o + S].sum()) / np.sqrt(var[o:o + S].sum()))
    rest = mx.sum() - mx[best:best + S].sum()
    margin = 1e-3 + np.abs(T).max(axis=1).sum() / 65536.0
    return best, float(rest + margin)


def fp32_scores(T32, idx, W):
    """Ascending-k fp32 sums (one rounding per row) of every window of the base-index array idx."""
    n = len(idx) - W + 1
    acc = T32[0][idx[0:n]].copy()
    for k in range(1, W):
        acc = (acc + T32[k][idx[k:k + n]]).astype(np.float32)
    return acc


def test_every_hit_has_a_passing_seed():
    rng = np.random.default_rng(7)
    n_checked = n_hits = 0
    for W in (13, 14, 16, 19, 24, 26):
        for _ in range(6):
            ppm = (rng.dirichlet([0.5] * 4, size=W) + 0.01) / 1.04
            T = np.log2(ppm / 0.25)
            idx = rng.integers(0, 4, size=60_000)
            for st in (0, 1):
                Ts = T if st == 0 else T[::-1, ::-1]          # Trc[k][b] = T[W-1-k][3-b] on the forward string
                T32 = Ts.astype(np.float32)
                for frac in (0.55, 0.7):
                    thr = frac * T.max(axis=1).sum()
                    full = fp32_scores(T32, idx, W)
                    hits = np.flatnonzero(full.astype(np.float64) > thr)
                    for S in S_VALUES:
                        o, rest = choose_seed(Ts, S)
                        seed = fp32_scores(T32[o:o + S], idx[o:len(idx) - (W - S - o)], S)
                        assert len(seed) == len(full)
                        assert np.all(seed[hits].astype(np.float64) > thr - rest), (W, st, frac, S)
                        n_checked += 1
                    n_hits += len(hits)
    assert n_checked == 6 * 6 * 2 * 2 * 2 and n_hits > 100


def test_seed_is_selective_on_the_c4_library():
    """The point of the engine: at thr 0.7 the seeds of a JASPAR-sized library pass ~0.1 windows per position and strand
    (12-base seeds), against 376 cells per position and strand for a per-cell prefilter."""
    from smeagol_b200 import synth

    pw = synth.synth_pwms(1000, widths=synth.jaspar_widths(1000))
    rng = np.random.default_rng(3)
    idx = rng.integers(0, 4, size=(20_000, 12))
    total = 0.0
    for w in pw.weights:
        T = np.asarray(w, dtype=np.float64)
        if T.shape[0] < 13:
            continue
        o, rest = choose_seed(T, 12)
        sc = T[o:o + 12][np.arange(12), idx].sum(axis=1)
        total += float((sc > 0.7 * T.max(axis=1).sum() - rest).mean())
    assert total < 0.25, total
