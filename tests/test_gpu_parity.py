"""GPU parity tests (run on the B200 box with -m gpu): the CUDA path, called through the C-ABI, against the CPU
oracle on the same seeded inputs and against the committed golden fixtures of the reference.

Bars: hit sets / positions / counts identical (bit-exact integer work); fp32 scores BIT-EXACT against the oracle's
kernel-order fp32 leg and within 1e-5 relative (SCORE_RTOL, north_star) of the reference-order leg and of the fp64
leg; sites inside the 1e-5 near-threshold band are decided in fp64 on both sides.
"""
import os

import numpy as np
import pandas as pd
import pytest

from golden_util import (SCORE_RTOL, assert_frame_matches, case_kwargs, fix_stats_ddof, golden, pwm_df,
                         pwms_from_json)
from oracle import frames as of
from oracle import scan_oracle as so

pytestmark = pytest.mark.gpu


def _sb():
    import smeagol_b200

    return smeagol_b200


def synth_pwms(rng, n, wmin, wmax, dtype=np.float64):
    ws = []
    for _ in range(n):
        W = int(rng.integers(wmin, wmax + 1))
        ppm = (rng.dirichlet([0.5] * 4, size=W) + 0.01) / 1.04
        ws.append(np.log2(ppm / 0.25).astype(dtype))
    return pd.DataFrame({"Matrix_id": ["M%04d" % i for i in range(n)], "weights": ws})


def rand_seq(rng, L, alphabet="ACGT", p=None):
    return "".join(rng.choice(list(alphabet), size=L, p=p))


ENGINE = None  # set by the engine-parametrised tests


def device_scan(model, seqs, frac, rcomp="both", chunk_len=None, segments=None, adjudicate=True):
    """Scan forward strings as ONE group -> (row, pos, motif, score) with rows [fwd..., rc...] like the oracle."""
    from smeagol_b200 import device as dev

    n = len(seqs)
    batch = dev.SeqBatch(seqs, segments=segments)
    if rcomp == "none":
        out_rows = np.stack([np.arange(n), np.zeros(n)], axis=1)
    elif rcomp == "only":
        out_rows = np.stack([np.zeros(n), np.arange(n)], axis=1)
    else:
        out_rows = np.stack([np.arange(n), n + np.arange(n)], axis=1)
    res = dev.scan(model.device_set, batch, model.thresholds(frac), dev.STRANDS[rcomp], out_rows.astype(np.int32),
                   2 * n if rcomp == "both" else n, adjudicate=adjudicate, chunk_len=chunk_len, engine=ENGINE)
    return res


def check_against_oracle(model, seqs, frac, rcomp="both", chunk_len=None, atol=0.0):
    """atol: only for thresholds near 0, where hits have scores near 0 and a relative bound is meaningless for any
    fp32 sum; 1e-5 absolute is 1e-5 relative to the PWM score scale there."""
    res = device_scan(model, seqs, frac, rcomp, chunk_len)
    row, pos, motif, sc = res.hits_host()
    exp = so.scan_rows(seqs, model.weights, frac, rcomp=rcomp, order="kernel", adjudicate=True)
    np.testing.assert_array_equal(row, exp["row"])
    np.testing.assert_array_equal(pos, exp["pos"])
    np.testing.assert_array_equal(motif, exp["motif"])
    # fp32 scores: bit-exact against the same summation order
    np.testing.assert_array_equal(sc.view(np.uint32), exp["score"].view(np.uint32))
    np.testing.assert_allclose(sc, exp["score64"], rtol=SCORE_RTOL, atol=atol)
    ref = so.scan_rows(seqs, model.weights, frac, rcomp=rcomp, order="ref", adjudicate=True)
    np.testing.assert_array_equal(pos, ref["pos"])
    np.testing.assert_allclose(sc, ref["score"], rtol=SCORE_RTOL, atol=atol)
    # count table (device) == counts of the hit list
    n_rows = len(exp["rows"])
    ct = so.count_table(exp, n_rows, len(model.weights))
    dev_counts = res.counts.cpu().numpy()
    n = len(seqs)
    if rcomp == "both":
        got = np.concatenate([dev_counts[:, :, 0], dev_counts[:, :, 1]], axis=0)
    elif rcomp == "none":
        got = dev_counts[:, :, 0]
    else:
        got = dev_counts[:, :, 1]
    np.testing.assert_array_equal(got, ct)
    assert res.stats["n_hits_total"] == len(exp["row"])
    return res, exp


def test_models_kat():
    """tests/test_models.py:30-74 of the reference, through the device path."""
    sb = _sb()
    G = golden()
    model = sb.models.PWMModel(pwm_df(G["test_pwms"]))
    assert np.all(model.Matrix_ids == ["x", "y", "z"])
    assert np.all(model.widths == [3, 5, 3])
    assert sb.utils._equals(model.max_scores, np.array([3.33376361, 5.52770673, 2.50816981]))
    assert model.channels == 3 and model.max_width == 5
    preds = model.predict(np.array([[1, 2, 3]]))
    assert preds.shape == (1, 3, 3)
    assert sb.utils._equals(preds[0, 0, :], np.array([-1.60847875, -7.32647115, 0.35611725000000005]))
    np.testing.assert_allclose(preds, np.array(G["test_models"]["predict_out"]), rtol=1e-6, atol=1e-6)
    (r, p, m), sc = model.predict_with_threshold(np.array([[1, 2, 3]]), 0.1, score=True)
    assert list(r) == [0] and list(p) == [0] and list(m) == [2]
    assert sb.utils._equals(sc, np.array([0.35611725]))
    with pytest.raises(AssertionError):
        model.predict_with_threshold(np.array([[1, 2, 3]]), 1.5)


def test_predict_dense_iupac():
    sb = _sb()
    P = golden()["predict_iupac"]
    model = sb.models.PWMModel(pwm_df(P["pwms"]))
    codes = sb.encode.integer_encode(P["seq"])
    got = model.predict(codes[None, :])
    np.testing.assert_allclose(got, np.array(P["out"]), rtol=1e-5, atol=2e-6)
    exp = so.predict(codes[None, :].astype(np.uint8), model.weights)
    np.testing.assert_array_equal(got.view(np.uint32), exp.view(np.uint32))  # same order -> bit-exact


def test_encode_errors_and_pack():
    sb = _sb()
    from smeagol_b200 import device as dev

    with pytest.raises(KeyError):
        dev.SeqBatch(["ACGTACGTACGTACGTAE"])
    with pytest.raises(KeyError):
        dev.SeqBatch(["acgt"])  # the reference's alphabet is upper-case only
    b = dev.SeqBatch(["ACGTUNZ", "WSMKRYBDHV" * 3])
    assert b.has_ambiguity
    packed = b.packed.cpu().numpy().view(np.uint32)
    amask = b.amask.cpu().numpy().view(np.uint16)
    codes = b.codes.cpu().numpy()
    assert packed[0] & 0x3FF == (0 | 1 << 2 | 2 << 4 | 3 << 6 | 3 << 8)
    assert amask[0] == 0xFFFF & ~0x1F
    assert list(codes[:7]) == [1, 2, 3, 4, 5, 6, 0]
    w1 = int(b.rows_host["word_off"][1])
    assert list(codes[w1 * 16:w1 * 16 + 10]) == [7, 8, 9, 10, 11, 12, 13, 14, 15, 16]
    assert not dev.SeqBatch(["ACGTU" * 10]).has_ambiguity


@pytest.mark.parametrize("i", range(13))
def test_reference_scan_cases(i):
    """scan_sequences of this framework vs the reference's recorded outputs and vs the adjudicated oracle."""
    sb = _sb()
    case = golden()["scan_cases"][i]
    model = sb.models.PWMModel(pwm_df(case["pwms"]))
    thr, kw = case_kwargs(case)
    recs = [sb.SeqRecord(sb.Seq(s), id=i_, name=n_) for s, i_, n_ in zip(case["seqs"], case["ids"], case["names"])]
    out = sb.scan.scan_sequences(recs, model, thr, **kw)
    ids, ws = pwms_from_json(case["pwms"])
    exp = of.scan_sequences(case["seqs"], case["ids"], case["names"], ids, ws, thr, order="ref", adjudicate=True, **kw)
    exp.pop("_n_flipped")
    assert set(out.keys()) == set(exp.keys())
    for k in exp:
        assert_frame_matches(out[k], exp[k], case["name"] + ":" + k)
    # against the reference's own recorded output: identical unless a site sits in the near-threshold band
    pure = of.scan_sequences(case["seqs"], case["ids"], case["names"], ids, ws, thr, order="ref", adjudicate=False, **kw)
    if pure.pop("_n_flipped") == 0:
        for k, e in case["outputs"].items():
            assert_frame_matches(out[k], fix_stats_ddof(k, e), case["name"] + ":golden:" + k)
    else:
        assert case["name"] == "no_hits"


def test_enrich_golden(tmp_path):
    """tests/test_enrich.py:28-69 of the reference end to end (sites, counts, background stats, p / fdr)."""
    sb = _sb()
    G = golden()
    E = G["test_enrich"]
    model = sb.models.PWMModel(pwm_df(G["test_pwms"]))
    genome = [sb.SeqRecord(sb.Seq(s), id=i, name=i) for s, i in zip(E["genome"], E["ids"])]
    fa = os.path.join(str(tmp_path), "shuf_genome.fa.gz")
    sb.io.write_fasta([sb.SeqRecord(sb.Seq(s), id=i, name=i, description="") for i, s in E["shuf_genome"]], fa)
    res = sb.enrich.enrich_in_genome(genome, model, simN=10, simK=2, rcomp="both", sense="+", threshold=0.65,
                                     background="binomial", shuf_seqs=fa)
    csv = E["csv"]
    sites = res["real_sites"]
    assert np.all(sites.start == np.array(csv["real_sites"]["start"]) - 1)
    assert np.all(sites.sense == csv["real_sites"]["sense"])
    assert list(sites.Matrix_id) == csv["real_sites"]["m"]
    counts = res["real_counts"]
    assert list(counts.Matrix_id) == csv["real_counts"]["Matrix_id"]
    assert list(counts.sense) == csv["real_counts"]["sense"]
    assert list(counts.num) == csv["real_counts"]["num"]
    st = res["shuf_stats"]
    assert list(st.Matrix_id) == csv["shuf_stats"]["Matrix_id"]
    assert list(st.sense) == csv["shuf_stats"]["sense"]
    assert np.all(st.avg == csv["shuf_stats"]["avg"])
    assert all(sb.utils._equals(a, b) for a, b in zip(st.sd, csv["shuf_stats"]["sd"]))
    enr = res["enrichment"]
    assert list(enr.Matrix_id) == csv["enr"]["Matrix_id"]
    assert list(enr.sense) == csv["enr"]["sense"]
    assert all(sb.utils._equals(a, b) for a, b in zip(enr.p, csv["enr"]["p"]))
    assert all(sb.utils._equals(a, b) for a, b in zip(enr.fdr, csv["enr"]["fdr"]))
    assert list(enr.adj_len) == csv["enr"]["adj_len"]


def test_all_width_buckets_bit_exact():
    """Every register-table instantiation (padded widths 2..32, dual and single strand) plus the wide fallback."""
    sb = _sb()
    rng = np.random.default_rng(7)
    ws = []
    for W in list(range(1, 41)) * 2:  # widths 1..40, two of each -> partial warps in every bucket
        ppm = (rng.dirichlet([0.5] * 4, size=W) + 0.01) / 1.04
        ws.append(np.log2(ppm / 0.25))
    model = sb.models.PWMModel(pd.DataFrame({"Matrix_id": ["W%02d_%d" % (w.shape[0], i) for i, w in enumerate(ws)],
                                             "weights": ws}))
    seqs = [rand_seq(rng, 1500), rand_seq(rng, 700), rand_seq(rng, 31), rand_seq(rng, 5)]
    check_against_oracle(model, seqs, 0.55, "both", chunk_len=256)
    check_against_oracle(model, seqs, 0.55, "none", chunk_len=4096)
    check_against_oracle(model, seqs, 0.55, "only", chunk_len=256)


def test_many_motifs_random_and_iupac():
    sb = _sb()
    rng = np.random.default_rng(11)
    model = sb.models.PWMModel(synth_pwms(rng, 150, 5, 24))
    amb = "ACGTNWSMKRYBDHVZ"  # (U together with T is "Mixed RNA/DNA" for Biopython's reverse_complement)
    pa = np.array([0.23] * 4 + [0.08 / 12] * 12)
    seqs = [rand_seq(rng, 3000), rand_seq(rng, 2111, amb, pa / pa.sum()), "N" * 300 + rand_seq(rng, 500) + "N" * 100]
    res, exp = check_against_oracle(model, seqs, 0.6, "both", chunk_len=512)
    assert len(exp["row"]) > 500


def test_thresholds_low_and_float32_weights():
    sb = _sb()
    rng = np.random.default_rng(13)
    model = sb.models.PWMModel(synth_pwms(rng, 40, 6, 12, dtype=np.float32))
    seqs = [rand_seq(rng, 800)]
    check_against_oracle(model, seqs, 0.3, "both", atol=1e-5)  # dense hits: staging flushes and buffer growth
    check_against_oracle(model, seqs, 0.0, "both", atol=1e-5)
    check_against_oracle(model, seqs, 1.0, "both")   # thr == max score: near-threshold band, fp64 adjudication


def test_pure_reference_compare_mode():
    """adjudicate=False reproduces `(double)score_f32 > thr` (models.py:108) on the kernel-order fp32 scores."""
    sb = _sb()
    rng = np.random.default_rng(17)
    model = sb.models.PWMModel(synth_pwms(rng, 60, 7, 12), adjudicate=False)
    seqs = [rand_seq(rng, 2000), rand_seq(rng, 999)]
    res = device_scan(model, seqs, 0.55, "both", adjudicate=False)
    row, pos, motif, sc = res.hits_host()
    exp = so.scan_rows(seqs, model.weights, 0.55, rcomp="both", order="kernel", adjudicate=False)
    np.testing.assert_array_equal(row, exp["row"])
    np.testing.assert_array_equal(pos, exp["pos"])
    np.testing.assert_array_equal(motif, exp["motif"])
    np.testing.assert_array_equal(sc.view(np.uint32), exp["score"].view(np.uint32))


def test_segments_halo_ownership_equals_whole():
    """A long sequence scanned as segments with a W_max-1 right halo (window owned by the segment containing its
    start) gives exactly the hit set of the whole sequence, on both strands."""
    sb = _sb()
    rng = np.random.default_rng(19)
    model = sb.models.PWMModel(synth_pwms(rng, 70, 6, 24))
    L = 10000
    seq = rand_seq(rng, L)
    whole = device_scan(model, [seq], 0.6, "both")
    halo = int(model.max_width) - 1
    bounds = [0, 1777, 4096, 4100, 9000, L]
    parts, segs = [], []
    for a, b in zip(bounds[:-1], bounds[1:]):
        e = min(L, b + halo)
        parts.append(seq[a:e])
        segs.append((a, L, b - a))
    from smeagol_b200 import device as dev

    batch = dev.SeqBatch(parts, segments=segs)
    n = len(parts)
    out_rows = np.stack([np.zeros(n), np.ones(n)], axis=1).astype(np.int32)  # all segments -> output rows 0 (+) / 1 (-)
    res = dev.scan(model.device_set, batch, model.thresholds(0.6), 3, out_rows, 2, chunk_len=256)
    a = whole.hits_host()
    b = res.hits_host()
    for x, y in zip(a[:3], b[:3]):
        np.testing.assert_array_equal(x, y)
    np.testing.assert_array_equal(a[3].view(np.uint32), b[3].view(np.uint32))
    np.testing.assert_array_equal(whole.counts.cpu().numpy().sum(axis=0), res.counts.cpu().numpy().sum(axis=0))
    assert whole.key_checksum() == res.key_checksum()


def test_size_independent_properties_large():
    """Config-2-like size (1,000,000 bases x 200 PWMs): properties that need no oracle -- chunk-size invariance,
    strand symmetry (scanning the reverse complement swaps the strands), counts == histogram of the hit list."""
    sb = _sb()
    rng = np.random.default_rng(23)
    model = sb.models.PWMModel(synth_pwms(rng, 200, 7, 12))
    L = 1_000_000
    seq = "".join(np.array(list("ACGT"))[rng.integers(0, 4, size=L)])
    a = device_scan(model, [seq], 0.7, "both", chunk_len=4096)
    b = device_scan(model, [seq], 0.7, "both", chunk_len=1024)
    assert a.n_hits == b.n_hits and a.key_checksum() == b.key_checksum()
    row, pos, motif, sc = a.hits_host()
    assert np.all(np.diff((row << 40) | (pos << 12) | motif) > 0)  # sorted, unique
    cnt = a.counts.cpu().numpy()[0]
    hist = np.zeros((200, 2), dtype=np.int64)
    np.add.at(hist, (motif, row), 1)
    np.testing.assert_array_equal(cnt, hist)
    rc = so.reverse_complement_str(seq)
    c = device_scan(model, [rc], 0.7, "both", chunk_len=4096)
    row_c, pos_c, motif_c, sc_c = c.hits_host()
    # '+' hits of rc(seq) at position p  <->  '-' hits of seq at position p (the reference scans the rc string)
    np.testing.assert_array_equal(pos_c[row_c == 0], pos[row == 1])
    np.testing.assert_array_equal(motif_c[row_c == 0], motif[row == 1])
    np.testing.assert_allclose(sc_c[row_c == 0], sc[row == 1], rtol=SCORE_RTOL)
    # a 1 Mb parity slice against the oracle is too slow in numpy for 200 PWMs; check the first 20 kb exactly
    check_against_oracle(model, [seq[:20000]], 0.7, "both")


def test_variant_golden_and_random():
    """tests/test_variant.py:47-63 of the reference + random sites/variants recorded from the reference."""
    sb = _sb()
    G = golden()
    V = G["test_variant"]
    recs = [sb.SeqRecord(sb.Seq(s), id=i, name=i) for s, i in zip(V["seqs"], V["ids"])]
    result = sb.variant.variant_effect_on_sites(pd.DataFrame(V["sites"]), pd.DataFrame(V["variants"]), recs,
                                                pwm_df(G["test_pwms"]))
    assert len(result) == 3
    assert list(result["name"]) == ["seq0", "seq1", "seq1"] and list(result["pos"]) == [2, 0, 1]
    assert list(result["alt"]) == ["G", "A", "A"] and list(result["Matrix_id"]) == ["x", "z", "z"]
    assert list(result["seq"]) == ["GAC", "GCT", "GCT"]
    assert sb.utils._equals(result.max_score.values, np.array([3.33376361, 2.50816981, 2.50816981]))
    assert sb.utils._equals(result.score.values, np.array([1.0, 0.769149, 0.769149]))
    assert sb.utils._equals(result.variant_score.values, np.array([0.4080668, 0.1419828, 0.769149]))
    R = G["variant_random"]
    recs = [sb.SeqRecord(sb.Seq(s), id=i, name=i) for s, i in zip(R["seqs"], R["ids"])]
    got = sb.variant.variant_effect_on_sites(pd.DataFrame(R["sites"]), pd.DataFrame(R["variants"]), recs, pwm_df(R["pwms"]))
    exp = R["result"]
    assert len(got) == len(exp["pos"])
    for c in ("name", "pos", "alt", "start", "Matrix_id", "end", "seq"):
        assert list(got[c]) == list(exp[c]), c
    # device tables are the fp32 casts of the fp64 weights: 1e-5 relative (absolute floor for scores near 0)
    np.testing.assert_allclose(got["score"], exp["score"], rtol=1e-5, atol=1e-6)
    np.testing.assert_allclose(got["variant_score"], exp["variant_score"], rtol=1e-5, atol=1e-6)


def test_variant_windows_batched():
    """BASELINE config 5 restatement vs the oracle's loop version."""
    sb = _sb()
    rng = np.random.default_rng(29)
    pw = synth_pwms(rng, 40, 5, 14)
    seq = rand_seq(rng, 400)
    pos = rng.integers(0, 400, size=50)
    pos[:4] = [0, 1, 398, 399]
    alts = [("ACGT".replace(seq[p], ""))[rng.integers(0, 3)] for p in pos]
    ref_max, alt_max = sb.variant.variant_window_scores([seq], np.zeros(50, dtype=np.int32), pos, alts, pw)
    e_ref, e_alt = so.variant_window_scan(seq, pos, alts, list(pw.weights))
    np.testing.assert_allclose(ref_max, e_ref, rtol=1e-5, atol=2e-5)
    np.testing.assert_allclose(alt_max, e_alt, rtol=1e-5, atol=2e-5)
    # wide PWMs (generic kernel), IUPAC codes near the SNVs and an ambiguous alternate base (flagged -> generic kernel)
    pw2 = synth_pwms(rng, 30, 3, 22)
    seq2 = rand_seq(rng, 150) + "NNRYACGTN" + rand_seq(rng, 141)
    pos2 = np.concatenate([np.arange(140, 170), rng.integers(0, 300, size=20)])
    alts2 = ["ACGT"[rng.integers(0, 4)] for _ in pos2]
    alts2[3] = "N"
    r2, a2 = sb.variant.variant_window_scores([seq2], np.zeros(len(pos2), dtype=np.int32), pos2, alts2, pw2)
    e2r, e2a = so.variant_window_scan(seq2, pos2, alts2, list(pw2.weights))
    np.testing.assert_allclose(r2, e2r, rtol=1e-5, atol=2e-5)
    np.testing.assert_allclose(a2, e2a, rtol=1e-5, atol=2e-5)


def test_benchmark_slice_full_parity_vs_c_oracle():
    """BASELINE config c2 inputs (first 60 of the 1,000 genomes x all 1,000 PWMs, thr 0.7, both strands): the whole
    count table, the hit count and an order-independent checksum of every hit key against the C oracle."""
    sb = _sb()
    from oracle import c_oracle
    from smeagol_b200 import device as dev
    from smeagol_b200 import synth

    pw = synth.synth_pwms(1000, 7, 12)
    genomes = synth.genome_set(1000, 10_000_000, 5000, 15000, seed=2)[:60]
    model = sb.models.PWMModel(pw)
    n = len(genomes)
    out_rows = np.stack([2 * np.arange(n), 2 * np.arange(n) + 1], axis=1).astype(np.int32)
    batch = dev.SeqBatch(genomes)
    res = dev.scan(model.device_set, batch, model.thresholds(0.7), 3, out_rows, 2 * n)
    lut = np.zeros(256, dtype=np.uint8)
    lut[list(b"ACGT")] = [1, 2, 3, 4]
    exp = c_oracle.scan([lut[g] for g in genomes], list(pw.weights), 0.7, strands=3, out_rows=out_rows, n_out_rows=2 * n)
    assert (res.pos_bits, res.motif_bits) == (exp["pos_bits"], exp["motif_bits"])
    np.testing.assert_array_equal(res.counts.cpu().numpy().astype(np.int64), exp["counts"])
    assert res.n_hits == exp["n_hits"] and res.stats["n_hits_total"] == exp["n_hits"]
    ksum, kxor = res.key_checksum()
    assert ksum % (1 << 64) == exp["key_sum"] and kxor % (1 << 64) == exp["key_xor"]
    assert res.stats["n_near"] >= exp["n_near"]  # the kernel's fp32 band is a superset of the oracle's fp64 band


def test_bucketed_sort_equals_radix_sort(monkeypatch):
    """csrc/sort_buckets.cu (scatter into (row, position block) buckets + shared-memory sort per bucket) against
    cub::DeviceRadixSort on the same unsorted hit lists: many rows, one long row, a hit-dense low-complexity stretch that
    overflows a bucket (the call must fall back, never truncate), tiny inputs."""
    import torch

    sb = _sb()
    from smeagol_b200 import device as dev

    rng = np.random.default_rng(59)
    model = sb.models.PWMModel(synth_pwms(rng, 120, 6, 14))
    cases = [[rand_seq(rng, int(n)) for n in rng.integers(50, 4000, size=40)], [rand_seq(rng, 300_000)],
             [rand_seq(rng, 20_000) + "A" * 30_000 + rand_seq(rng, 20_000)], ["ACGTACGTAC"], [rand_seq(rng, 64)]]
    for seqs in cases:
        n = len(seqs)
        batch = dev.SeqBatch(seqs)
        out_rows = np.stack([np.arange(n), n + np.arange(n)], axis=1).astype(np.int32)
        plan = dev.ScanPlan(model.device_set, batch, model.thresholds(0.55), 3, out_rows, 2 * n)
        plan.launch()
        res = plan.finish(sort=False)
        if res.n_hits == 0:
            continue
        monkeypatch.setenv("SMEAGOL_B200_SORT", "radix")
        k0, s0 = dev.sort_hits(res.keys, res.scores, res.n_hits, plan.key_bits)
        monkeypatch.setenv("SMEAGOL_B200_SORT", "bucket")
        k1, s1 = dev.sort_hits(res.keys, res.scores, res.n_hits, plan.key_bits, motif_bits=plan.motif_bits,
                               total_positions=2 * batch.total_own)
        assert torch.equal(k0, k1) and torch.equal(s0.view(torch.int32), s1.view(torch.int32))
        assert bool((k1[1:] > k1[:-1]).all().item()) if res.n_hits > 1 else True
    monkeypatch.delenv("SMEAGOL_B200_SORT")


def test_buffer_overflow_is_detected_and_retried():
    """Hit / near-threshold buffers that are too small are regrown and the scan repeated -- never truncated."""
    sb = _sb()
    from smeagol_b200 import device as dev

    rng = np.random.default_rng(31)
    model = sb.models.PWMModel(synth_pwms(rng, 30, 6, 10))
    seqs = [rand_seq(rng, 3000)]
    batch = dev.SeqBatch(seqs)
    out_rows = np.array([[0, 1]], dtype=np.int32)
    thr = model.thresholds(1.0)  # every window at the maximum score is near-threshold
    full = dev.scan(model.device_set, batch, model.thresholds(0.4), 3, out_rows, 2)
    tiny = dev.scan(model.device_set, batch, model.thresholds(0.4), 3, out_rows, 2, hit_cap=7, near_cap=1)
    assert tiny.stats["hit_cap"] >= full.n_hits > 7
    for a, b in zip(full.hits_host(), tiny.hits_host()):
        np.testing.assert_array_equal(a, b)
    near = dev.scan(model.device_set, batch, thr, 3, out_rows, 2, near_cap=1)
    exp = so.scan_rows(seqs, model.weights, 1.0, rcomp="both", order="kernel")
    np.testing.assert_array_equal(near.hits_host()[1], exp["pos"])


def test_edge_inputs():
    """Empty and tiny rows, a row made only of N, sequences shorter than every PWM, a single PWM."""
    sb = _sb()
    rng = np.random.default_rng(37)
    model = sb.models.PWMModel(synth_pwms(rng, 5, 6, 9))
    seqs = ["", "A", "ACGTA", "N" * 50, rand_seq(rng, 200)]
    check_against_oracle(model, seqs, 0.5, "both")
    one = sb.models.PWMModel(synth_pwms(rng, 1, 8, 8))
    check_against_oracle(one, [rand_seq(rng, 500)], 0.5, "both")
    recs = [sb.SeqRecord(sb.Seq("ACG"), id="s", name="s")]
    out = sb.scan.scan_sequences(recs, model, 0.5, rcomp="both", outputs=["sites", "counts"], score=True)
    assert len(out["sites"]) == 0 and len(out["counts"]) == 0
    assert list(out["sites"].columns) == ["id", "name", "sense", "start", "Matrix_id", "width", "end", "score", "max_score",
                                          "frac_score"]
    assert list(out["counts"].columns) == ["Matrix_id", "width", "id", "name", "sense", "num"]
    with pytest.raises(KeyError):
        sb.scan.scan_sequences([sb.SeqRecord(sb.Seq("ACGTX"), id="s", name="s")], model, 0.5)
    with pytest.raises(ValueError):
        sb.scan.scan_sequences([sb.SeqRecord(sb.Seq("ACGT"), id="a", name="g"), sb.SeqRecord(sb.Seq("ACG"), id="b", name="g")],
                               model, 0.5, group_by="name")
    with pytest.raises(AssertionError):
        sb.scan.scan_sequences(recs, model, 1.5)


@pytest.fixture(params=["hmma", "tc5"])
def hmma_engine(request):
    """'hmma': warp-level mma.sync prefilter (scan_hmma.cuh); 'tc5': tcgen05.mma + TMEM prefilter (scan_tc5.cuh)."""
    global ENGINE
    ENGINE = request.param
    yield request.param
    ENGINE = None


def test_tensor_core_engine_bit_identical(hmma_engine):
    """The tensor-core prefilters + exact second pass must give the same hits, fp32 scores and counts as the oracle
    (and hence as the register-table engine): widths 1..32, IUPAC codes, every rcomp mode, chunk boundaries."""
    sb = _sb()
    rng = np.random.default_rng(41)
    ws = []
    for W in list(range(1, 33)) * 2:
        ppm = (rng.dirichlet([0.5] * 4, size=W) + 0.01) / 1.04
        ws.append(np.log2(ppm / 0.25))
    model = sb.models.PWMModel(pd.DataFrame({"Matrix_id": ["W%02d_%d" % (w.shape[0], i) for i, w in enumerate(ws)],
                                             "weights": ws}), tc_min_width=0)
    amb = "ACGTNWSMKRYBDHVZ"
    pa = np.array([0.23] * 4 + [0.08 / 12] * 12)
    seqs = [rand_seq(rng, 1500), rand_seq(rng, 777, amb, pa / pa.sum()), rand_seq(rng, 31), "ACGTA", ""]
    res, exp = check_against_oracle(model, seqs, 0.55, "both", chunk_len=256)
    assert res.stats["engine"] == "hmma" and res.stats["tc_kind"] == hmma_engine
    assert res.stats["n_candidates"] >= len(exp["row"])
    check_against_oracle(model, seqs, 0.55, "none", chunk_len=4096)
    check_against_oracle(model, seqs, 0.55, "only", chunk_len=256)
    model2 = sb.models.PWMModel(synth_pwms(rng, 150, 5, 24), tc_min_width=0)
    check_against_oracle(model2, [rand_seq(rng, 3000), "N" * 300 + rand_seq(rng, 500) + "N" * 100], 0.6, "both", chunk_len=512)
    check_against_oracle(model2, [rand_seq(rng, 900)], 1.0, "both")
    check_against_oracle(model2, [rand_seq(rng, 900)], 0.2, "both", atol=1e-5)
    # many tiles per CTA, several column blocks and column groups, rows that end inside a tile, default chunking
    model3 = sb.models.PWMModel(synth_pwms(rng, 700, 6, 24), tc_min_width=0)
    long_seqs = [rand_seq(rng, 40000), rand_seq(rng, 5000, "ACGTN", [0.24, 0.24, 0.24, 0.24, 0.04]), rand_seq(rng, 513),
                 rand_seq(rng, 511), rand_seq(rng, 1024)]
    res3, exp3 = check_against_oracle(model3, long_seqs, 0.7, "both")
    assert len(exp3["row"]) > 10000


def test_hybrid_engine_split_bit_identical():
    """Hybrid engine: PWMs narrower than the split on the register-table kernel, the others on the tensor-core
    prefilter + exact second pass, both writing the same hit / count buffers -- results identical to the oracle."""
    sb = _sb()
    rng = np.random.default_rng(43)
    ws = []
    for W in list(range(1, 41)) * 2:  # includes PWMs wider than 32 (generic kernel on the smg_scan side)
        ppm = (rng.dirichlet([0.5] * 4, size=W) + 0.01) / 1.04
        ws.append(np.log2(ppm / 0.25))
    df = pd.DataFrame({"Matrix_id": ["W%02d_%d" % (w.shape[0], i) for i, w in enumerate(ws)], "weights": ws})
    amb = "ACGTNWSMKRYBDHVZ"
    pa = np.array([0.23] * 4 + [0.08 / 12] * 12)
    seqs = [rand_seq(rng, 1500), rand_seq(rng, 777, amb, pa / pa.sum()), rand_seq(rng, 31), ""]
    # any width may be the split, also inside a K-depth class and above every PWM; split 2 has the width-1 PWMs (padded
    # to width 2 in their lane group) on the register-table side -- they must not fall between the engines
    for split in (2, 5, 10, 13, 22, 33):
        model = sb.models.PWMModel(df, tc_min_width=split)
        res, exp = check_against_oracle(model, seqs, 0.55, "both", chunk_len=256)
        assert res.stats["engine"] == "auto"
        if split == 5:
            assert res.stats["n_candidates"] > 0  # the tensor-core side really ran
        check_against_oracle(model, seqs, 0.55, "only", chunk_len=4096)
    plain = sb.models.PWMModel(df, tc_min_width=0)
    res, _ = check_against_oracle(plain, seqs, 0.55, "both")
    assert res.stats["engine"] == "regtab"


def test_kmer_index_engine_bit_identical(monkeypatch):
    """K-mer index engine (csrc/kmer.cu): every k-mer of the width classes up to 12 scored once per call, hits emitted by
    table lookup -- identical hits, fp32 scores (bit for bit), counts and near-threshold tallies as the oracle.  Covers
    every width class 1..12 (+ wider PWMs on the other engines of the same scan), IUPAC / 'Z' windows (generic path),
    all strand modes, rows ending inside a window, thresholds at both ends, and a forced class limit below a width that
    is present."""
    sb = _sb()
    rng = np.random.default_rng(53)
    ws = []
    for W in list(range(1, 15)) * 3:
        ppm = (rng.dirichlet([0.5] * 4, size=W) + 0.01) / 1.04
        ws.append(np.log2(ppm / 0.25))
    df = pd.DataFrame({"Matrix_id": ["W%02d_%d" % (w.shape[0], i) for i, w in enumerate(ws)], "weights": ws})
    amb = "ACGTNWSMKRYBDHVZ"
    pa = np.array([0.23] * 4 + [0.08 / 12] * 12)
    seqs = [rand_seq(rng, 6000), rand_seq(rng, 1500, amb, pa / pa.sum()), rand_seq(rng, 31), "ACGTA", "",
            "N" * 40 + rand_seq(rng, 300) + "N" * 40]
    for wk in ("12", "11", "7", "3"):
        monkeypatch.setenv("SMEAGOL_B200_KMER_MAX_WIDTH", wk)
        # the seed engine is chosen by a cost model (long scans only): force it here, with 12- and 11-base seeds
        if wk in ("12", "11"):
            monkeypatch.setenv("SMEAGOL_B200_SEED_LEN", wk)
        else:
            monkeypatch.delenv("SMEAGOL_B200_SEED_LEN", raising=False)
        model = sb.models.PWMModel(df)  # default split: engine 'auto' = k-mer index + register tables + tensor cores
        res, exp = check_against_oracle(model, seqs, 0.6, "both", chunk_len=512)
        assert res.stats["engine"] == "auto" and res.stats["kmer_max_width"] == int(wk)
        # a 12-mer index also carries the SEEDS of the wider PWMs (W 13, 14 here) instead of the tensor-core prefilter
        assert res.stats["tc_kind"] == ("seed" if wk in ("12", "11") else "tc5")
        assert res.stats["kmer_index_entries"] > 0
        check_against_oracle(model, seqs, 0.6, "none")
        check_against_oracle(model, seqs, 0.6, "only", chunk_len=4096)
    monkeypatch.delenv("SMEAGOL_B200_SEED_LEN", raising=False)
    monkeypatch.setenv("SMEAGOL_B200_KMER_MAX_WIDTH", "10")
    model = sb.models.PWMModel(df)
    check_against_oracle(model, seqs, 1.0, "both")
    check_against_oracle(model, [rand_seq(rng, 2000)], 0.15, "both", atol=1e-5)  # dense lists: many entries per k-mer
    # segments with halo: ownership and reverse-complement coordinates come from the row descriptor
    whole = rand_seq(rng, 5000)
    res_w = device_scan(model, [whole], 0.6)
    L, halo = len(whole), model.max_width - 1
    cuts = [0, 1600, 3328, L]
    parts = [whole[a:min(L, b + halo)] for a, b in zip(cuts[:-1], cuts[1:])]
    segs = [(a, L, b - a) for a, b in zip(cuts[:-1], cuts[1:])]
    from smeagol_b200 import device as dev

    batch = dev.SeqBatch(parts, segments=segs)
    out_rows = np.array([[0, 1]] * 3, dtype=np.int32)
    res_s = dev.scan(model.device_set, batch, model.thresholds(0.6), 3, out_rows, 2)
    a, b = res_w.hits_host(), res_s.hits_host()
    for x, y in zip(a, b):
        np.testing.assert_array_equal(x, y)
    monkeypatch.delenv("SMEAGOL_B200_KMER_MAX_WIDTH")
    # automatic choice: 4^W <= window starts of the scan
    model = sb.models.PWMModel(df)
    res, _ = check_against_oracle(model, [rand_seq(rng, 70000)], 0.7, "both")
    assert res.stats["kmer_max_width"] == 8


@pytest.mark.gpu
def test_nonfinite_weights_never_hit_like_the_reference():
    """log2 of a zero probability gives -inf weights.  The reference contracts the full one-hot vector with every
    weight of a row (models.py:52-68): 0 * -inf = NaN, NaN > thr is False (models.py:108), so such a PWM has no site
    anywhere.  Both engines must agree (they add only the selected weight and would otherwise report hits)."""
    sb = _sb()
    rng = np.random.default_rng(47)
    ws = []
    for i, W in enumerate([6, 8, 9, 11, 12, 14, 17, 24]):
        w = np.log2((rng.dirichlet([0.5] * 4, size=W) + 0.01) / 1.04 / 0.25)
        if i % 2 == 1:
            w[rng.integers(0, W), rng.integers(0, 4)] = -np.inf
        ws.append(w)
    df = pd.DataFrame({"Matrix_id": ["M%d" % i for i in range(len(ws))], "weights": ws})
    seqs = [rand_seq(rng, 3000), rand_seq(rng, 500, "ACGTN", [0.24, 0.24, 0.24, 0.24, 0.04])]
    with np.errstate(invalid="ignore"):
        for split in (0, 5, 10):
            model = sb.models.PWMModel(df, tc_min_width=split)
            res, exp = check_against_oracle(model, seqs, 0.5, "both")
            assert set(np.unique(exp["motif"])) == {0, 2, 4, 6}
            assert res.stats["n_near"] < 1000  # the dead PWMs are switched off, not sent to the adjudicator


@pytest.mark.gpu
def test_fasta_device_ingest_matches_host_parser(tmp_path):
    """FASTA text -> packed SeqBatch on the device (csrc/fasta.cu) against the host parser that
    mirrors io.read_fasta (io.py:16-41): same records, same de-wrapped sequences, same scan results.  Edge cases: text
    before the first header, CRLF line ends, blank lines, blanks inside lines, '>' inside a sequence line, a header as
    the last line, no trailing newline, short and long lines, many records (headers straddling 16-byte loads), gzip."""
    sb = _sb()
    from smeagol_b200 import device as dev
    from smeagol_b200 import io as sio

    rng = np.random.default_rng(53)
    texts = []
    texts.append("junk line\n>a first record\nACGT ACGT\r\nAC\n\n>b\nTTTT\n>c  two  spaces\nGATTACA")
    parts = []
    for i in range(300):  # many short records, line widths 1..90
        L = int(rng.integers(1, 400))
        s = rand_seq(rng, L, "ACGTN", [0.24, 0.24, 0.24, 0.24, 0.04])
        w = int(rng.integers(1, 91))
        eol = "\r\n" if i % 7 == 0 else "\n"
        parts.append(">r%d len=%d%s" % (i, L, eol) + eol.join(s[j:j + w] for j in range(0, L, w)) + eol)
    texts.append("".join(parts))
    big = rand_seq(rng, 300_000)
    texts.append(">chrBig\n" + "\n".join(big[j:j + 60] for j in range(0, len(big), 60)) + "\n>tail\nACGTACGTAC\n")
    model = sb.models.PWMModel(synth_pwms(rng, 40, 5, 14))
    for t, text in enumerate(texts):
        for gz in (False, True):
            path = str(tmp_path / ("f%d.fa%s" % (t, ".gz" if gz else "")))
            if gz:
                import gzip
                with gzip.open(path, "wt", newline="") as fh:
                    fh.write(text)
            else:
                with open(path, "w", newline="") as fh:
                    fh.write(text)
            host = sio.read_fasta(path)
            fb = sio.read_fasta_device(path)
            assert fb.ids == [r.id for r in host] and fb.names == [r.name for r in host]
            assert fb.descriptions == [r.description for r in host]
            assert list(fb.lengths) == [len(r.seq) for r in host]
            assert [str(r.seq) for r in fb.records()] == [str(r.seq) for r in host]
            if not gz:
                seqs = [str(r.seq) for r in host]
                keep = [i for i, s in enumerate(seqs) if len(s) > 0]
                n = len(seqs)
                out_rows = np.stack([np.arange(n), n + np.arange(n)], axis=1).astype(np.int32)
                thr = model.thresholds(0.6)
                a = dev.scan(model.device_set, fb.batch, thr, 3, out_rows, 2 * n)
                b = dev.scan(model.device_set, dev.SeqBatch(seqs), thr, 3, out_rows, 2 * n)
                assert a.n_hits == b.n_hits and len(keep) > 0
                np.testing.assert_array_equal(a.keys[:a.n_hits].cpu().numpy(), b.keys[:b.n_hits].cpu().numpy())
                np.testing.assert_array_equal(a.counts.cpu().numpy(), b.counts.cpu().numpy())
    # lower case: KeyError like the reference unless upper=True
    path = str(tmp_path / "lower.fa")
    with open(path, "w") as fh:
        fh.write(">x\nACGTacgtNNnn\nacgt\n")
    with pytest.raises(KeyError):
        sio.read_fasta_device(path)
    fb = sio.read_fasta_device(path, upper=True)
    assert [str(r.seq) for r in fb.records()] == ["ACGTACGTNNNNACGT"]
    with pytest.raises(ValueError):
        sio.read_fasta_device(None, data=b"no header here\nACGT\n")
    with pytest.raises(KeyError, match=">"):  # '>' inside a line is sequence text: the encoder rejects it (encode.py:66)
        sio.read_fasta_device(None, data=b">b\nTT>TT\n")


@pytest.mark.gpu
def test_kmer_shuffle_device_properties_and_distribution():
    """utils.shuffle_records on the device.  Every copy must have the properties of the reference's
    shuffles (utils.py:36-83; simK = 2 is deeplift.dinuc_shuffle's algorithm: all dinucleotide counts, first and last
    base preserved; simK = 1: base counts), be reproducible from the seed, differ between copies and seeds, and sample
    the same distribution as the CPU restatement (oracle/shuffle_oracle.py): compared on the per-position base
    frequencies of a short sequence over many copies."""
    sb = _sb()
    from oracle import shuffle_oracle as sho
    from smeagol_b200 import utils as su

    rng = np.random.default_rng(59)
    recs = [sb.SeqRecord(sb.Seq(rand_seq(rng, 3000)), id="g0", name="g0"),
            sb.SeqRecord(sb.Seq(rand_seq(rng, 257, "ACGTN", [0.3, 0.2, 0.2, 0.25, 0.05])), id="g1", name="g1"),
            sb.SeqRecord(sb.Seq("A"), id="g2", name="g2"), sb.SeqRecord(sb.Seq("ACGTTTGA"), id="g3", name="g3")]
    for k in (1, 2):
        out = su.shuffle_records(recs, 6, simK=k, seed=7)
        assert [r.id for r in out] == ["background_seq_%d" % i for i in range(6)] * 4
        assert [r.name for r in out] == ["g0"] * 6 + ["g1"] * 6 + ["g2"] * 6 + ["g3"] * 6
        for i, r in enumerate(out):
            sho.check_shuffle(str(recs[i // 6].seq), str(r.seq), k)
        assert len({str(r.seq) for r in out[:6]}) == 6                       # copies differ
        again = su.shuffle_records(recs, 6, simK=k, seed=7)
        assert [str(r.seq) for r in again] == [str(r.seq) for r in out]      # seeded
        other = su.shuffle_records(recs, 6, simK=k, seed=8)
        assert [str(r.seq) for r in other[:6]] != [str(r.seq) for r in out[:6]]
    # a long sequence takes the global-scratch path (> 200 KiB)
    big = sb.SeqRecord(sb.Seq(rand_seq(rng, 300_000)), id="big", name="big")
    for k in (1, 2):
        o = su.shuffle_records([big], 2, simK=k, seed=3)
        for r in o:
            sho.check_shuffle(str(big.seq), str(r.seq), k)
        assert str(o[0].seq) != str(o[1].seq)
    # distribution: per-position base frequencies over many copies, device vs the CPU restatement of the algorithm
    seq = "ACGTTGCAAGCTTAGGCTAACGTCAGT"
    n = 4000
    dev_copies = [str(r.seq) for r in su.shuffle_records([sb.SeqRecord(sb.Seq(seq), id="s", name="s")], n, simK=2, seed=11)]
    rs = np.random.RandomState(5)
    cpu_copies = [sho.dinuc_shuffle(seq, rs) for _ in range(n)]
    for c in dev_copies[:50]:
        sho.check_shuffle(seq, c, 2)

    def freq(copies):
        a = np.frombuffer("".join(copies).encode(), dtype=np.uint8).reshape(len(copies), len(seq))
        return np.stack([(a == ord(b)).mean(axis=0) for b in "ACGT"])

    assert np.abs(freq(dev_copies) - freq(cpu_copies)).max() < 0.05   # ~4 sigma of a binomial at n = 4000
    assert len(set(dev_copies)) > 0.5 * len(set(cpu_copies))          # comparable diversity of Eulerian paths
    # the packed batch scans like the read-back records
    ids, names, batch, _ = su.shuffle_batch_device(recs[:2], 3, 2, seed=7)
    back = su.shuffle_records(recs[:2], 3, simK=2, seed=7)
    from smeagol_b200 import device as dev
    model = sb.models.PWMModel(synth_pwms(rng, 20, 5, 12))
    nrow = len(ids)
    out_rows = np.stack([np.arange(nrow), nrow + np.arange(nrow)], axis=1).astype(np.int32)
    a = dev.scan(model.device_set, batch, model.thresholds(0.6), 3, out_rows, 2 * nrow)
    b = dev.scan(model.device_set, dev.SeqBatch([str(r.seq) for r in back]), model.thresholds(0.6), 3, out_rows, 2 * nrow)
    np.testing.assert_array_equal(a.keys[:a.n_hits].cpu().numpy(), b.keys[:b.n_hits].cpu().numpy())


@pytest.mark.gpu
def test_binned_counts_on_device_equal_host_route():
    """outputs=['binned_counts'] alone is histogrammed on the device (smg_bin_hits, no hit leaves the
    GPU); asking for the sites as well takes the reference's per-hit pd.cut + crosstab on the host.  Same frames."""
    sb = _sb()
    rng = np.random.default_rng(61)
    model = sb.models.PWMModel(synth_pwms(rng, 30, 5, 14))
    recs = [sb.SeqRecord(sb.Seq(rand_seq(rng, int(L))), id="s%d" % i, name="g%d" % (i % 2))
            for i, L in enumerate([1500, 900, 1500, 900])]  # records sharing a name must have equal lengths (encode.py:172)
    bins = np.array([0.5, 0.6, 0.7, 0.8, 0.95])
    for kw in (dict(sense="+", rcomp="both"), dict(sense="-", rcomp="only", group_by="name"),
               dict(sense="+", rcomp="both", group_by="name", combine_seqs=True, sep_ids=True),
               dict(sense="+", rcomp="none", combine_seqs=True)):
        a = sb.scan.scan_sequences(recs, model, bins, outputs=["binned_counts"], **kw)["binned_counts"]
        b = sb.scan.scan_sequences(recs, model, bins, outputs=["binned_counts", "sites"], **kw)["binned_counts"]
        assert list(a.columns) == list(b.columns) and len(a) == len(b) and len(a) > 0
        for col in a.columns:
            assert [str(x) for x in a[col]] == [str(x) for x in b[col]], (kw, col)
        assert int(a.num.sum()) == int(b.num.sum())


@pytest.mark.gpu
def test_device_resident_records_through_the_reference_api(tmp_path):
    """Records produced on the device (GPU shuffles, device FASTA ingest) carry DeviceSeq sequences: scan_sequences /
    enrich_in_genome pack them without a host round trip and return the same frames as for host strings; str(), len(),
    slicing and reverse_complement() still work for callers that want the text."""
    sb = _sb()
    from smeagol_b200 import io as sio
    from smeagol_b200 import utils as su
    from smeagol_b200.seqrecord import DeviceSeq

    rng = np.random.default_rng(67)
    model = sb.models.PWMModel(synth_pwms(rng, 25, 5, 13))
    genome = [sb.SeqRecord(sb.Seq(rand_seq(rng, 4000)), id="virus", name="virus")]
    shuf = su.shuffle_records(genome, 12, simK=2, seed=5)
    assert all(isinstance(r.seq, DeviceSeq) for r in shuf) and len(shuf[0]) == 4000
    host = [sb.SeqRecord(sb.Seq(str(r.seq)), id=r.id, name=r.name) for r in shuf]
    assert str(shuf[3].seq[10:20]) == str(host[3].seq)[10:20]
    assert str(shuf[3].seq.reverse_complement()) == str(host[3].seq.reverse_complement())
    kw = dict(sense="+", rcomp="both", outputs=["sites", "counts", "stats"], group_by="name", combine_seqs=True, sep_ids=True)
    a = sb.scan.scan_sequences(shuf, model, 0.7, **kw)
    b = sb.scan.scan_sequences(host, model, 0.7, **kw)
    for k in ("sites", "counts", "stats"):
        pd.testing.assert_frame_equal(a[k].reset_index(drop=True), b[k].reset_index(drop=True))
    # enrichment with backgrounds generated on the device == the same call on their host copies
    e1 = sb.enrich.enrich_in_genome(genome, model, "both", 0.7, simN=12, simK=2, shuf_seqs=shuf)
    e2 = sb.enrich.enrich_in_genome(genome, model, "both", 0.7, simN=12, simK=2, shuf_seqs=host)
    pd.testing.assert_frame_equal(e1["enrichment"].reset_index(drop=True), e2["enrichment"].reset_index(drop=True))
    e3 = sb.enrich.enrich_in_genome(genome, model, "both", 0.7, simN=12, simK=2)  # default path: shuffles on the GPU
    assert list(e3["enrichment"].columns) == list(e1["enrichment"].columns) and len(e3["shuf_stats"]) > 0
    # device FASTA ingest -> records -> scan_sequences
    path = str(tmp_path / "g.fa")
    sio.write_fasta(genome + host[:3], path, gz=False)
    fb = sio.read_fasta_device(path)
    recs = fb.records()
    assert all(isinstance(r.seq, DeviceSeq) for r in recs)
    c = sb.scan.scan_sequences(recs, model, 0.7, sense="+", rcomp="both", outputs=["sites"], score=True)["sites"]
    d = sb.scan.scan_sequences(sio.read_fasta(path), model, 0.7, sense="+", rcomp="both", outputs=["sites"], score=True)["sites"]
    pd.testing.assert_frame_equal(c.reset_index(drop=True), d.reset_index(drop=True))
