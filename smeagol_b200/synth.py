"""Synthetic inputs of the benchmark configurations (BASELINE.json configs c1..c5; generators as fixed in
SURVEY.md section 8(d)): numpy.random.default_rng(seed), ACGT only unless stated.  Sequences are returned as uint8
ASCII arrays (no Python strings are needed on the device path)."""
import numpy as np
import pandas as pd

_ACGT = np.frombuffer(b"ACGT", dtype=np.uint8)
JASPAR_WIDTHS = np.arange(6, 25)
# fixed discrete distribution on {6..24} with mean ~ 12 ("JASPAR-sized", config c4)
_JW = np.array([4, 6, 9, 11, 12, 12, 11, 9, 7, 5, 4, 3, 2, 1.5, 1.2, 1, 0.6, 0.4, 0.3])
JASPAR_P = _JW / _JW.sum()


def synth_pwms(n, wmin=7, wmax=12, seed_base=1000, widths=None, dtype=np.float64):
    """PWM m (seed seed_base+m): W ~ U{wmin..wmax} (or widths[m]); PPM rows ~ Dirichlet(0.5), mixed (p+0.01)/1.04;
    PWM = log2(PPM/0.25) (smeagol/matrices.py:123-137); ids M0000..."""
    ids, ws = [], []
    for m in range(n):
        rng = np.random.default_rng(seed_base + m)
        W = int(widths[m]) if widths is not None else int(rng.integers(wmin, wmax + 1))
        ppm = (rng.dirichlet([0.5] * 4, size=W) + 0.01) / 1.04
        ws.append(np.log2(ppm / 0.25).astype(dtype))
        ids.append("M%04d" % m)
    return pd.DataFrame({"Matrix_id": ids, "weights": ws})


def jaspar_widths(n, seed=4):
    return np.random.default_rng(seed).choice(JASPAR_WIDTHS, size=n, p=JASPAR_P)


def random_bases(rng, n):
    return _ACGT[rng.integers(0, 4, size=n, dtype=np.uint8)]


def genome_set(n_genomes=1000, total=10_000_000, lmin=5000, lmax=15000, seed=2):
    """c2: n_genomes lengths ~ U{lmin..lmax} rescaled to `total` bases."""
    rng = np.random.default_rng(seed)
    lens = rng.integers(lmin, lmax + 1, size=n_genomes).astype(np.float64)
    lens = np.maximum(1, np.floor(lens * (total / lens.sum()))).astype(np.int64)
    lens[-1] += total - lens.sum()
    return [random_bases(rng, int(L)) for L in lens]


def single_sequence(L, seed):
    return random_bases(np.random.default_rng(seed), L)


def mono_shuffles(seq, n, seed=3):
    """Fixed pre-generated backgrounds for c3 (seeded permutations of the genome; the reference's dinucleotide
    shuffle lives in deeplift and is out of scope -- BASELINE config 3 uses fixed pre-generated copies)."""
    rng = np.random.default_rng(seed)
    return [seq[rng.permutation(seq.shape[0])] for _ in range(n)]


def snvs(seq, n, wmax, seed=5):
    """c5: n SNVs at uniform positions >= wmax from either end, alt != ref uniform."""
    rng = np.random.default_rng(seed)
    pos = rng.integers(wmax, seq.shape[0] - wmax, size=n)
    ref_idx = np.searchsorted(_ACGT, seq[pos])
    alt_idx = (ref_idx + rng.integers(1, 4, size=n)) % 4
    return pos.astype(np.int32), (alt_idx + 1).astype(np.uint8)  # integer codes A=1..T=4
