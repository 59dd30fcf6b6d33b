"""Golden vectors for the SURVEY 8(f) rank-4 rows and the threshold / window drivers of enrich.py (run ONCE in the
build container, where /root/reference exists):

    python tests/golden/make_golden_f4.py

Imports the UNMODIFIED reference modules through oracle/ref_harness.py (see make_golden.py) and records results of
  * matrices.ncorr / pairwise_ncorrs (matrices.py:356-433) on the reference's own fixtures (tests/test_matrices.py:176-213)
    and on seeded random PWMs of widths 3..20;
  * utils.get_tiling_windows_over_genome (utils.py:135-180), scan.count_in_sliding_windows (scan.py:381-423) and
    enrich.enrich_in_sliding_windows (enrich.py:285-374) on the sites of a recorded reference scan;
  * enrich.examine_thresholds' real_binned frame (enrich.py:206-284; the shuffled half depends on the random stream).
Output: tests/golden/reference_golden_f4.json.  Nothing at test time reads /root/reference.
"""
import json
import os
import sys

import numpy as np
import pandas as pd

HERE = os.path.dirname(os.path.abspath(__file__))
REPO = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, REPO)
sys.path.insert(0, HERE)

from oracle import ref_harness  # noqa: E402

ref_harness.install()

from Bio.Seq import Seq  # noqa: E402  (stand-ins)
from Bio.SeqRecord import SeqRecord  # noqa: E402
import smeagol.enrich as r_enrich  # noqa: E402
import smeagol.matrices as r_mat  # noqa: E402
import smeagol.models as r_models  # noqa: E402
import smeagol.scan as r_scan  # noqa: E402
import smeagol.utils as r_utils  # noqa: E402

from make_golden import jsonable, load_test_pwms, synth_pwms  # noqa: E402


def main():
    out = {"_generator": "tests/golden/make_golden_f4.py", "_reference": "gruber-sciencelab/SMEAGOL v0.1.1 (unmodified modules)"}
    # ---- PWM similarity
    X = np.array([[-1.30065948, -6.65821148, 1.668218, -1.30065948], [0.99284021, 0.99284021, -6.65821148, -6.65821148],
                  [0.26065175, 0.6727054, -1.30065948, -0.31836148]])
    Y = np.array([[-0.99106688, -1.97336487, -0.41205364, 1.31654155], [-7.33091688, 0.99551261, 0.80350944, -1.97336487],
                  [1.57897621, -0.41205364, -7.33091688, -1.97336487], [0.0, 0.0, 0.32013481, -0.41205364],
                  [-0.99106688, -1.97336487, 1.31654155, -0.41205364]])
    out["ncorr_kat"] = {"X": X.tolist(), "Y": Y.tolist(), "result": float(r_mat.ncorr(X, Y, min_overlap=3)),
                        "result_swapped": float(r_mat.ncorr(Y, X, min_overlap=3))}
    tp = load_test_pwms()
    out["pairwise_test_pwms"] = {"result": jsonable(r_mat.pairwise_ncorrs(list(tp.weights)))}
    rng = np.random.default_rng(401)
    rp = synth_pwms(rng, 14, 3, 20)
    rp_w = list(rp.weights) + [np.asarray(rp.weights[0]).copy()]  # a duplicate: similarity exactly 1
    out["pairwise_random"] = {"weights": [np.asarray(w).tolist() for w in rp_w], "result": jsonable(r_mat.pairwise_ncorrs(rp_w))}
    # ---- windows
    rng = np.random.default_rng(402)
    seqs = ["".join(rng.choice(list("ACGT"), size=n)) for n in (1300, 700)]
    genome = [SeqRecord(Seq(s), id="rec%d" % i, name="rec%d" % i) for i, s in enumerate(seqs)]
    pw = synth_pwms(rng, 6, 5, 9)
    model = r_models.PWMModel(pw)
    sites = r_scan.scan_sequences(genome, model, 0.6, sense="+", rcomp="both", outputs=["sites"], score=True)["sites"]
    mid = str(pw.Matrix_id[2])
    cases = []
    for width, shift in ((200, None), (150, 50), (64, 64), (500, 333)):
        w = r_utils.get_tiling_windows_over_genome(genome, width, shift)
        c = r_scan.count_in_sliding_windows(sites, genome, mid, width, shift)
        e = r_enrich.enrich_in_sliding_windows(sites, genome, width, shift, matrix_id=mid)
        cases.append({"width": width, "shift": shift, "windows": jsonable(w.reset_index(drop=True)), "counts": jsonable(c),
                      "enrich": jsonable(e)})
    out["windows"] = {"seqs": seqs, "ids": ["rec0", "rec1"], "pwms": {"Matrix_id": list(pw.Matrix_id),
                                                                      "weights": [np.asarray(w).tolist() for w in pw.weights]},
                      "threshold": 0.6, "matrix_id": mid, "n_sites": int(len(sites)), "cases": cases}
    # ---- examine_thresholds: the real half is deterministic
    r_enrich.shuffle_records = lambda records, simN, simK=2, **kw: [
        SeqRecord(Seq(str(r.seq)[::-1]), id="background_seq_%d" % i, name=r.id) for r in records for i in range(simN)]
    for combine in (True, False):
        res = r_enrich.examine_thresholds(genome, model, simN=2, simK=2, rcomp="both", min_threshold=0.5, threshold_shift=0.1,
                                          combine_seqs=combine)
        out["examine_thresholds_combine_%s" % combine] = {
            "real_binned": jsonable(res["real_binned"]), "shuf_binned_reversed_records": jsonable(res["shuf_binned"])}
    with open(os.path.join(HERE, "reference_golden_f4.json"), "w") as f:
        json.dump(out, f)
    print("wrote reference_golden_f4.json:", {k: (len(v) if hasattr(v, "__len__") else v) for k, v in out.items()})


if __name__ == "__main__":
    main()
