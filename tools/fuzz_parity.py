"""Randomised parity sweep of the device scan against the numpy oracle (kernel-order fp32 scores bit for bit, hit
sets, counts): random PWM libraries (widths 1-40), IUPAC / 'Z' sequences, thresholds, strand modes, engine splits and
chunk lengths.  python tools/fuzz_parity.py [trials] [seed]   -- test infrastructure (imports oracle/)."""
import os
import sys

import numpy as np
import pandas as pd

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import test_gpu_parity as T  # noqa: E402

trials = int(sys.argv[1]) if len(sys.argv) > 1 else 40
rng = np.random.default_rng(int(sys.argv[2]) if len(sys.argv) > 2 else 12345)
sb = T._sb()
amb = "ACGTNWSMKRYBDHVZ"
for t in range(trials):
    n = int(rng.integers(1, 160))
    wmax = int(rng.choice([8, 12, 16, 24, 33, 40]))
    ws = []
    for _ in range(n):
        W = int(rng.integers(1, wmax + 1))
        ppm = (rng.dirichlet([0.5] * 4, size=W) + 0.01) / 1.04
        w = np.log2(ppm / 0.25)
        ws.append(w.astype(np.float32) if rng.random() < 0.2 else w)
    df = pd.DataFrame({"Matrix_id": ["M%d" % i for i in range(n)], "weights": ws})
    seqs = []
    for _ in range(int(rng.integers(1, 6))):
        L = int(rng.choice([0, 1, 7, 33, 257, 1000, 4097, 9000]))
        if rng.random() < 0.4:
            pa = np.array([0.22] * 4 + [0.12 / 12] * 12)
            seqs.append(T.rand_seq(rng, L, amb, pa / pa.sum()))
        else:
            seqs.append(T.rand_seq(rng, L))
    if all(len(s) == 0 for s in seqs):
        seqs.append("ACGTACGTAC")
    frac = float(rng.choice([0.35, 0.5, 0.6, 0.7, 0.8, 0.95]))
    rcomp = str(rng.choice(["both", "none", "only"]))
    split = int(rng.choice([0, 1, 5, 7, 9, 10, 13, 20, 33]))
    chunk = int(rng.choice([256, 512, 1024, 2048, 4096]))
    model = sb.models.PWMModel(df, tc_min_width=split)
    try:
        res, exp = T.check_against_oracle(model, seqs, frac, rcomp, chunk_len=chunk, atol=1e-5)
    except AssertionError as e:
        print("MISMATCH trial", t, dict(n=n, wmax=wmax, frac=frac, rcomp=rcomp, split=split, chunk=chunk, lens=[len(s) for s in seqs]))
        print(str(e)[:2000])
        sys.exit(1)
    print("trial %d ok: %d PWMs (W<=%d), %d rows, frac %.2f, %s, split %d, chunk %d, %d hits, engine %s" % (
        t, n, wmax, len(seqs), frac, rcomp, split, chunk, len(exp["row"]), res.stats["engine"]), flush=True)
print("all", trials, "trials passed")
