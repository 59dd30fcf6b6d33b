"""Golden-vector generator (run ONCE in the build container, where /root/reference exists):

    python tests/golden/make_golden.py

It imports the UNMODIFIED reference modules (smeagol/encode.py, models.py, scan.py, enrich.py, variant.py,
utils.py) through oracle/ref_harness.py (stand-ins only for the missing third-party packages; the Keras graph of
models.py:50-68 is evaluated with torch CPU conv1d in fp32) and records

  1. the reference's own test fixtures for the hot path (tests/data/test_pwms.hdf5 PWMs, real_sites.csv,
     real_counts.csv, shuf_stats.csv, enr.csv, the FASTA fixtures) -- data only, no source code;
  2. the results of the calls made by the reference's tests (test_models.py, test_scan.py, test_enrich.py,
     test_variant.py, test_encode.py);
  3. results of the reference's scan_sequences / predict / variant_effect_on_sites on seeded random inputs
     (ACGT, RNA, IUPAC ambiguity codes, several rcomp / group_by / combine / sep_ids / seq_batch settings).

Output: tests/golden/reference_golden.json.  Nothing at test time reads /root/reference.
"""
import gzip
import json
import os
import pickle
import sys

import numpy as np
import pandas as pd

HERE = os.path.dirname(os.path.abspath(__file__))
REPO = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, REPO)

from oracle import ref_harness  # noqa: E402

ref_harness.install()
REF_DATA = os.path.join(ref_harness.REFERENCE_ROOT, "tests", "data")

from Bio.Seq import Seq  # noqa: E402  (stand-ins)
from Bio.SeqRecord import SeqRecord  # noqa: E402
import smeagol.encode as r_encode  # noqa: E402
import smeagol.enrich as r_enrich  # noqa: E402
import smeagol.models as r_models  # noqa: E402
import smeagol.scan as r_scan  # noqa: E402
import smeagol.variant as r_variant  # noqa: E402
from smeagol.io import read_fasta  # noqa: E402


def jsonable(x):
    if isinstance(x, pd.DataFrame):
        return {c: jsonable(x[c].to_numpy(dtype=object) if x[c].dtype.kind not in "fiub" else x[c].to_numpy())
                for c in x.columns}
    if isinstance(x, np.ndarray):
        if x.dtype.kind in "fc":
            return [float(v) for v in x.ravel()] if x.ndim == 1 else [jsonable(v) for v in x]
        if x.dtype.kind in "iub":
            return [int(v) for v in x.ravel()] if x.ndim == 1 else [jsonable(v) for v in x]
        return [jsonable(v) for v in x.tolist()]
    if isinstance(x, (np.floating,)):
        return float(x)
    if isinstance(x, (np.integer,)):
        return int(x)
    if isinstance(x, (list, tuple)):
        return [jsonable(v) for v in x]
    if isinstance(x, dict):
        return {str(k): jsonable(v) for k, v in x.items()}
    if x is None or isinstance(x, (str, int, float, bool)):
        return x
    return str(x)


def load_test_pwms():
    """tests/data/test_pwms.hdf5 is a pandas/PyTables 'fixed' store whose payload is one pickled object array."""
    raw = open(os.path.join(REF_DATA, "test_pwms.hdf5"), "rb").read()
    start = raw.find(b"\x80\x04\x95")
    obj = pickle.loads(raw[start:])
    return pd.DataFrame({"Matrix_id": [o[0] for o in obj], "weights": [np.asarray(o[1]) for o in obj]})


def synth_pwms(rng, n, wmin, wmax, dtype=np.float64):
    """SURVEY section 8(d) generator: Dirichlet(0.5) PPM mixed with 0.01 pseudo-probability, PWM = log2(PPM/0.25)
    (matrices.py:123-137)."""
    ids, ws = [], []
    for m in range(n):
        W = int(rng.integers(wmin, wmax + 1))
        ppm = rng.dirichlet([0.5] * 4, size=W)
        ppm = (ppm + 0.01) / 1.04
        ws.append(np.log2(ppm / 0.25).astype(dtype))
        ids.append("M%04d" % m)
    return pd.DataFrame({"Matrix_id": ids, "weights": ws})


def pwms_json(df):
    return {"Matrix_id": list(df.Matrix_id), "dtype": str(df.weights[0].dtype),
            "weights": [[[float(v) for v in row] for row in w] for w in df.weights]}


def rand_seq(rng, L, alphabet="ACGT", p=None):
    return "".join(rng.choice(list(alphabet), size=L, p=p))


def recs(seqs, ids=None, names=None):
    ids = ids or ["s%d" % i for i in range(len(seqs))]
    names = names or ids
    return [SeqRecord(Seq(s), id=i, name=n) for s, i, n in zip(seqs, ids, names)]


def main():
    G = {"_generator": "tests/golden/make_golden.py", "_reference": "gruber-sciencelab/SMEAGOL v0.1.1 (/root/reference)"}
    df = load_test_pwms()
    G["test_pwms"] = pwms_json(df)
    model = r_models.PWMModel(df)

    # ---- tests/test_models.py:30-74
    enc = np.array([[1, 2, 3]])
    thr, sc = model.predict_with_threshold(enc, 0.1, score=True)
    G["test_models"] = {
        "Matrix_ids": list(model.Matrix_ids), "widths": jsonable(model.widths),
        "max_scores": jsonable(model.max_scores), "channels": int(model.channels),
        "max_width": int(model.max_width),
        "conv_kernel": jsonable(np.asarray(model.model.layers[4].weights[0])),
        "predict_in": [[1, 2, 3]], "predict_out": jsonable(model.predict(enc)),
        "pwt_frac": 0.1, "pwt_idx": jsonable([np.asarray(t) for t in thr]), "pwt_scores": jsonable(sc),
        "expected_in_reference_test": {"max_scores": [3.33376361, 5.52770673, 2.50816981],
                                       "preds_0_0": [-1.60847875, -7.32647115, 0.35611725000000005],
                                       "pwt_score": [0.35611725]},
    }

    # ---- tests/test_scan.py:16-79
    records = recs(["ATTAAA", "GCTATA"], ["S1.1", "S1.2"], ["S1", "S1"])
    encoding = r_encode.SeqEncoding(records, sense="+", rcomp="none")
    thresholded = [[1, 1], [0, 0], [0, 2]]
    scores = np.array([2.343, 1.929])
    thr2, sc2 = model.predict_with_threshold(encoding.seqs, 0.25, True)
    G["test_scan"] = {
        "seqs": ["ATTAAA", "GCTATA"], "ids": ["S1.1", "S1.2"], "names": ["S1", "S1"],
        "locate_sites": jsonable(r_scan._locate_sites(encoding, model, thresholded, scores)),
        "count_sites": jsonable(r_scan._count_sites(encoding, model, thresholded)),
        "pwt_025_idx": jsonable([np.asarray(t) for t in thr2]), "pwt_025_scores": jsonable(sc2),
        "bin_sites": jsonable(r_scan._bin_sites_by_score(encoding, model, thr2, sc2, [0.25, 0.5, 0.75])
                              .assign(bin=lambda d: d.bin.astype(str))),
    }

    # ---- tests/test_encode.py:13-85
    G["test_encode"] = {
        "alphabet": "ACGTNWSMKRYBDHVZ",
        "codes": jsonable(r_encode.integer_encode(SeqRecord(Seq("ACGTNWSMKRYBDHVZ"), id="id", name="name"))),
        "rc_AGCAU": jsonable(r_encode.integer_encode(SeqRecord(Seq("AGCAU")), rc=True)),
        "one_hot_ACGTN": jsonable(r_encode.one_hot_encode("ACGTN")),
        "test_fa": [[r.id, r.name, str(r.seq)] for r in read_fasta(os.path.join(REF_DATA, "test.fa.gz"))],
    }
    for rc in ("none", "only", "both"):
        e = r_encode.SeqEncoding(recs(["AGC", "ACA"], ["id1", "id2"], ["name1", "name2"]), sense="+", rcomp=rc)
        G["test_encode"]["SeqEncoding_" + rc] = {"ids": list(e.ids), "names": list(e.names),
                                                 "seqs": jsonable(e.seqs), "senses": list(e.senses), "len": e.len}

    # ---- tests/test_enrich.py:15-69
    genome_seqs = ["AGAGCGCATTTCCTACGCATGCTCGATCAAATGCTACGGATTCTAAAA", "CCCGGACGTTTGCTCGAGGG"]
    genome = recs(genome_seqs, ["seg1", "seg2"], ["seg1", "seg2"])
    shuf_path = os.path.join(REF_DATA, "shuf_genome.fa.gz")
    # pandas >= 2 maps np.std inside .agg([...]) to ddof=0 while the golden shuf_stats.csv is ddof=1
    # (SURVEY H6); record both the live result and the CSV fixtures.
    res = r_enrich.enrich_in_genome(genome, model, simN=10, simK=2, rcomp="both", sense="+", threshold=0.65,
                                    background="binomial", shuf_seqs=shuf_path)
    G["test_enrich"] = {
        "genome": genome_seqs, "ids": ["seg1", "seg2"], "threshold": 0.65, "simN": 10,
        "shuf_genome": [[r.id, str(r.seq)] for r in read_fasta(shuf_path)],
        "real_sites": jsonable(res["real_sites"]), "real_counts": jsonable(res["real_counts"]),
        "shuf_counts": jsonable(res["shuf_counts"]), "shuf_stats_live_pandas3": jsonable(res["shuf_stats"]),
        "enrichment_live": jsonable(res["enrichment"]),
        "csv": {n: jsonable(pd.read_csv(os.path.join(REF_DATA, n + ".csv")))
                for n in ("real_sites", "real_counts", "shuf_stats", "enr")},
    }

    # ---- tests/test_variant.py:16-63
    vrec = recs(["GACAAA", "GCTATA"], ["seq0", "seq1"], ["seq0", "seq1"])
    sites = pd.DataFrame({"name": ["seq0", "seq1"], "start": [0, 0], "Matrix_id": ["x", "z"], "width": [3, 3],
                          "end": [3, 3]})
    variants = pd.DataFrame({"name": ["seq0", "seq0", "seq1", "seq1"], "pos": [2, 5, 0, 1],
                             "alt": ["G", "T", "A", "A"]})
    vres = r_variant.variant_effect_on_sites(sites, variants, vrec, df)
    G["test_variant"] = {"seqs": ["GACAAA", "GCTATA"], "ids": ["seq0", "seq1"], "sites": jsonable(sites),
                         "variants": jsonable(variants),
                         "result": jsonable(vres.drop(columns=["weights", "ref_scores"], errors="ignore"))}

    # ---- seeded random cases through the reference's own scan_sequences
    cases = []

    def add_case(name, seqs, ids, names, pw, threshold, **kw):
        m = r_models.PWMModel(pw)
        out = r_scan.scan_sequences(recs(seqs, ids, names), m, threshold, **kw)
        c = {"name": name, "seqs": seqs, "ids": ids, "names": names, "pwms": pwms_json(pw),
             "threshold": jsonable(threshold), "kwargs": jsonable(kw), "outputs": {}}
        for k, v in out.items():
            if k == "binned_counts":
                v = v.assign(bin=v.bin.astype(str))
            c["outputs"][k] = jsonable(v)
        cases.append(c)

    rng = np.random.default_rng(20261019)
    pw_small = synth_pwms(rng, 12, 4, 12)
    s = [rand_seq(rng, 400), rand_seq(rng, 250), rand_seq(rng, 333)]
    add_case("acgt_both_sites_counts", s, ["a", "b", "c"], ["a", "b", "c"], pw_small, 0.6, sense="+", rcomp="both",
             outputs=["sites", "counts"], score=True)
    add_case("acgt_none", s, ["a", "b", "c"], ["a", "b", "c"], pw_small, 0.55, sense="+", rcomp="none",
             outputs=["sites", "counts"], score=True)
    add_case("acgt_only_minus_sense", s, ["a", "b", "c"], ["a", "b", "c"], pw_small, 0.55, sense="-", rcomp="only",
             outputs=["sites", "counts"], score=True)
    add_case("acgt_combine", s, ["a", "b", "c"], ["a", "b", "c"], pw_small, 0.6, sense="+", rcomp="both",
             outputs=["counts"], combine_seqs=True)
    add_case("acgt_combine_sep_ids", s, ["a", "b", "c"], ["a", "b", "c"], pw_small, 0.6, sense="+", rcomp="both",
             outputs=["counts"], combine_seqs=True, sep_ids=True)
    # group_by='name' with equal-length members (the enrich background path, enrich.py:172-183)
    bg = [rand_seq(rng, 300) for _ in range(6)] + [rand_seq(rng, 180) for _ in range(6)]
    bg_ids = ["background_seq_%d" % (i % 6) for i in range(12)]
    bg_names = ["g1"] * 6 + ["g2"] * 6
    add_case("background_groupby_stats", bg, bg_ids, bg_names, pw_small, 0.6, sense="+", rcomp="both",
             outputs=["counts", "stats"], group_by="name", combine_seqs=True, sep_ids=True, seq_batch=4)
    add_case("background_groupby_nocombine", bg, bg_ids, bg_names, pw_small, 0.6, sense="+", rcomp="both",
             outputs=["sites", "counts"], group_by="name", score=True)  # per-group 'stats' (scan.py:198-206) raises
    # under pandas 3 (mean over the string 'id' column) -- reference version rot, not recorded
    # IUPAC ambiguity codes incl. Z, and an RNA sequence (U)
    amb = "ACGTNWSMKRYBDHVZ"
    pa = np.array([0.22, 0.22, 0.22, 0.22] + [0.01] * 12)
    pa = pa / pa.sum()
    s_amb = [rand_seq(rng, 500, amb, pa), rand_seq(rng, 200, amb, pa)]
    add_case("iupac_both", s_amb, ["n1", "n2"], ["n1", "n2"], pw_small, 0.5, sense="+", rcomp="both",
             outputs=["sites", "counts"], score=True)
    s_rna = [rand_seq(rng, 300, "ACGU")]
    add_case("rna_both", s_rna, ["r1"], ["r1"], pw_small, 0.6, sense="+", rcomp="both",
             outputs=["sites", "counts"], score=True)
    # wider PWMs (JASPAR-sized, 6..24) and float32 weights (read_pms_from_file dtype, io.py:134)
    pw_wide = synth_pwms(rng, 10, 6, 24, dtype=np.float32)
    add_case("wide_float32_weights", [rand_seq(rng, 2000)], ["w1"], ["w1"], pw_wide, 0.65, sense="+", rcomp="both",
             outputs=["sites", "counts"], score=True)
    # binned counts (scan.py:77-111) and a sequence shorter than the widest PWM
    add_case("binned_counts", s[:2], ["a", "b"], ["a", "b"], pw_small, np.array([0.5, 0.6, 0.8]), sense="+",
             rcomp="both", outputs=["binned_counts"])
    add_case("short_sequence", ["ACGTAC", rand_seq(rng, 40)], ["t1", "t2"], ["t1", "t2"], pw_small, 0.5,
             sense="+", rcomp="both", outputs=["sites", "counts"], score=True)
    add_case("no_hits", [rand_seq(rng, 100)], ["z1"], ["z1"], pw_small, 1.0, sense="+", rcomp="both",
             outputs=["sites", "counts"], score=True)
    G["scan_cases"] = cases

    # dense predict on IUPAC codes (ambiguity through the conv is unpinned by the reference tests)
    m_small = r_models.PWMModel(pw_small)
    codes = r_encode.integer_encode(s_amb[1])
    G["predict_iupac"] = {"seq": s_amb[1], "pwms": pwms_json(pw_small),
                          "out": jsonable(m_small.predict(codes[None, :]))}

    # variant effects on random sites
    vs = [rand_seq(rng, 120), rand_seq(rng, 90)]
    vrecs = recs(vs, ["v0", "v1"], ["v0", "v1"])
    vsites = r_scan.scan_sequences(vrecs, m_small, 0.5, sense="+", rcomp="none", outputs=["sites"], score=False)["sites"]
    nv = 60
    vdf = pd.DataFrame({"name": rng.choice(["v0", "v1"], size=nv), "pos": rng.integers(0, 90, size=nv),
                        "alt": rng.choice(list("ACGT"), size=nv)})
    vout = r_variant.variant_effect_on_sites(vsites, vdf, vrecs, pw_small)
    G["variant_random"] = {"seqs": vs, "ids": ["v0", "v1"], "pwms": pwms_json(pw_small), "sites": jsonable(vsites),
                           "variants": jsonable(vdf),
                           "result": jsonable(vout.drop(columns=["weights", "ref_scores"], errors="ignore"))}

    out_path = os.path.join(HERE, "reference_golden.json")
    with open(out_path, "w") as f:
        json.dump(G, f)
    print("wrote", out_path, os.path.getsize(out_path), "bytes")


if __name__ == "__main__":
    main()
