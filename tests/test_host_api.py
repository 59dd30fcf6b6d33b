"""CPU tests of the host-side mirror of the reference API (no GPU): encode / SeqEncoding / SeqGroups / FASTA /
threshold bounds / key layout / count melting.  Restates tests/test_encode.py and the schema tests of
tests/test_scan.py of the reference against this package."""
import gzip
import os

import numpy as np
import pandas as pd
import pytest

import smeagol_b200 as sb
from golden_util import assert_frame_matches, golden, pwm_df
from smeagol_b200 import device as dev
from smeagol_b200.encode import SeqEncoding, SeqGroups, integer_encode, one_hot_encode
from smeagol_b200.seqrecord import Seq, SeqRecord


def test_integer_encode():
    """reference tests/test_encode.py:13-23"""
    record = SeqRecord(Seq("ACGTNWSMKRYBDHVZ"), id="id", name="name")
    assert np.all(integer_encode(record, rc=False) == [1, 2, 3, 4, 6, 7, 8, 9, 10, 11, 12, 13, 14, 15, 16, 0])
    assert integer_encode(record).dtype == np.float32
    with pytest.raises(KeyError):
        integer_encode(SeqRecord(Seq("ACGE")), rc="none")
    assert np.all(integer_encode(SeqRecord(Seq("AGCAU")), rc=True) == [1, 5, 3, 2, 5])
    assert np.all(integer_encode("ACGT", rc=True) == [1, 2, 3, 4])
    with pytest.raises(TypeError):
        integer_encode(5)
    G = golden()["test_encode"]
    np.testing.assert_array_equal(one_hot_encode("ACGTN"), np.array(G["one_hot_ACGTN"], dtype=np.float32))


def test_SeqEncoding():
    """reference tests/test_encode.py:26-51"""
    records = [SeqRecord(Seq("AGC"), id="id1", name="name1"), SeqRecord(Seq("ACA"), id="id2", name="name2")]
    result = SeqEncoding(records, sense="+")
    assert result.len == 3
    assert np.all(result.ids == ["id1", "id2"]) and np.all(result.names == ["name1", "name2"])
    assert np.all(result.seqs == np.array([[1, 3, 2], [1, 2, 1]])) and np.all(result.senses == ["+", "+"])
    assert type(result.senses) == np.ndarray and type(result.ids) == np.ndarray and type(result.names) == np.ndarray
    result = SeqEncoding(records, sense="+", rcomp="only")
    assert np.all(result.seqs == np.array([[3, 2, 4], [4, 3, 4]])) and np.all(result.senses == ["-", "-"])
    result = SeqEncoding(records, sense="+", rcomp="both")
    assert np.all(result.ids == ["id1", "id2", "id1", "id2"])
    assert np.all(result.names == ["name1", "name2", "name1", "name2"])
    assert np.all(result.seqs == np.array([[1, 3, 2], [1, 2, 1], [3, 2, 4], [4, 3, 4]]))
    assert np.all(result.senses == ["+", "+", "-", "-"])
    with pytest.raises(ValueError):
        SeqEncoding([SeqRecord(Seq("AGC"), id="a", name="a"), SeqRecord(Seq("AG"), id="b", name="b")], sense="+")
    with pytest.raises(AssertionError):
        SeqEncoding(records, sense="+", rcomp="sideways")
    for rc in ("none", "only", "both"):
        e = SeqEncoding(records, sense="+", rcomp=rc)
        g = golden()["test_encode"]["SeqEncoding_" + rc]
        assert list(e.ids) == g["ids"] and list(e.senses) == g["senses"] and e.len == g["len"]
        np.testing.assert_array_equal(e.seqs, np.array(g["seqs"]))


def test_SeqGroups_and_fasta(tmp_path):
    """reference tests/test_encode.py:54-85 (test.fa.gz content is replayed from the golden file)"""
    fa = os.path.join(str(tmp_path), "test.fa.gz")
    with gzip.open(fa, "wt") as f:
        for rid, name, seq in golden()["test_encode"]["test_fa"]:
            f.write(">%s Test sequence\n%s\n" % (rid, seq))
    result = SeqGroups(fa, sense="+")
    assert [g.records[0] for g in result.seqs] == ["ATTAAATA", "CAAAATCTTTAGGATTAGCAC"]
    assert [g.ids[0] for g in result.seqs] == ["Seg1", "Seg2"] and [g.names[0] for g in result.seqs] == ["Seg1", "Seg2"]
    np.testing.assert_array_equal(result.seqs[0].seqs, [[1, 4, 4, 1, 1, 1, 4, 1]])
    records = [SeqRecord(Seq("AGC"), id="id1", name="name"), SeqRecord(Seq("ACA"), id="id2", name="name")]
    result = SeqGroups(records, sense="-", rcomp="both", group_by="name")
    assert len(result.seqs) == 1
    expected = SeqEncoding(records, sense="-", rcomp="both")
    assert np.all(result.seqs[0].seqs == expected.seqs) and np.all(result.seqs[0].senses == ["-", "-", "+", "+"])
    out = os.path.join(str(tmp_path), "o.fa")
    sb.io.write_fasta(records, out)
    back = sb.io.read_fasta(out)
    assert [str(r.seq) for r in back] == ["AGC", "ACA"] and [r.id for r in back] == ["id1", "id2"]


def test_model_constants_and_thresholds():
    """reference tests/test_models.py:38-43"""
    model = sb.models.PWMModel(pwm_df(golden()["test_pwms"]))
    assert np.all(model.Matrix_ids == ["x", "y", "z"]) and np.all(model.widths == [3, 5, 3])
    assert sb.utils._equals(model.max_scores, np.array([3.33376361, 5.52770673, 2.50816981]))
    assert model.channels == 3 and model.max_width == 5
    with pytest.raises(AssertionError):
        model.thresholds(1.2)
    thr = model.thresholds(0.65)
    lo, hi = dev.threshold_bounds(thr, np.array([10.0, 12.0, 8.0]), adjudicate=False)
    assert np.all(lo == hi) and np.all(lo.astype(np.float64) <= thr)
    assert np.all(np.nextafter(lo, np.float32(np.inf)).astype(np.float64) > thr)
    lo, hi = dev.threshold_bounds(thr, np.array([10.0, 12.0, 8.0]), adjudicate=True)
    assert np.all(lo.astype(np.float64) <= thr - 1e-5 * np.array([10.0, 12.0, 8.0]))
    assert np.all(hi.astype(np.float64) >= thr + 1e-5 * np.array([10.0, 12.0, 8.0]))


def test_locate_and_count_sites_schema():
    """reference tests/test_scan.py:28-47, 73-79"""
    G = golden()["test_scan"]
    model = sb.models.PWMModel(pwm_df(golden()["test_pwms"]))
    recs = [SeqRecord(Seq(s), id=i, name=n) for s, i, n in zip(G["seqs"], G["ids"], G["names"])]
    encoding = SeqEncoding(recs, sense="+", rcomp="none")
    thresholded = [[1, 1], [0, 0], [0, 2]]
    res = sb.scan._locate_sites(encoding, model, thresholded, np.array([2.343, 1.929]))
    assert_frame_matches(res, G["locate_sites"], "locate_sites")
    assert sb.utils._equals(res.frac_score.values, np.array([0.70280927926, 0.76908668]))
    cnt = sb.scan._count_sites(encoding, model, thresholded)
    assert_frame_matches(cnt, G["count_sites"], "count_sites")
    empty = sb.scan._count_sites(encoding, model, ([], [], []))
    assert list(empty.columns) == ["Matrix_id", "width", "id", "name", "sense", "num"] and len(empty) == 0


def test_chunks_and_row_layout_host_side():
    lens = np.array([100, 4096, 5000, 1], dtype=np.int64)
    rows = np.zeros(4, dtype=dev.ROW_DTYPE)
    rows["n_own"] = lens

    class B:
        rows_host = rows
        n_rows = 4
    ch = dev.SeqBatch.chunks(B, 4096)
    assert list(ch["row"]) == [0, 1, 2, 2, 3] and list(ch["start"]) == [0, 0, 0, 4096, 0]
    assert dev.pick_chunk_len(10_000_000, 32) == 4096 and dev.pick_chunk_len(30_000, 4) == 256
    assert dev._bits(1000) == 10 and dev._bits(1024) == 10 and dev._bits(1025) == 11 and dev._bits(1) == 1


def test_read_fasta_host_semantics(tmp_path):
    """io.read_fasta mirrors Biopython's FastaIterator (reference io.py:16-41): text before the first header is skipped,
    id = name = first token, description = the whole header line, line breaks / blanks / carriage returns removed."""
    from smeagol_b200 import io as sio

    p = tmp_path / "a.fa"
    p.write_bytes(b"junk\n>a first record\nACGT ACGT\r\nAC\n\n>b\nTTTT\n>c  two  spaces\nGATTACA")
    recs = sio.read_fasta(str(p))
    assert [(r.id, r.name, r.description, str(r.seq)) for r in recs] == [
        ("a", "a", "a first record", "ACGTACGTAC"), ("b", "b", "b", "TTTT"), ("c", "c", "c  two  spaces", "GATTACA")]
    sio.write_fasta(recs, str(tmp_path / "b.fa.gz"))
    again = sio.read_fasta(str(tmp_path / "b.fa.gz"))
    assert [str(r.seq) for r in again] == [str(r.seq) for r in recs] and [r.id for r in again] == ["a", "b", "c"]


def test_shuffle_oracle_and_successor_buckets():
    """The CPU restatement of deeplift.dinuc_shuffle's algorithm keeps its defining properties, and the host pre-pass
    of the device shuffle (successors bucketed by symbol, original order) matches a direct construction."""
    from oracle import shuffle_oracle as sho
    from smeagol_b200 import utils as su
    from smeagol_b200.encode import integer_encoding_dict

    rng = np.random.default_rng(3)
    seq = "".join(rng.choice(list("ACGTN"), size=500, p=[0.3, 0.2, 0.2, 0.25, 0.05]))
    rs = np.random.RandomState(1)
    copies = [sho.dinuc_shuffle(seq, rs) for _ in range(5)]
    for c in copies:
        sho.check_shuffle(seq, c, 2)
    assert len(set(copies)) == 5 and seq not in copies
    tok = np.array([integer_encoding_dict[b] for b in seq], dtype=np.uint8)
    succ, off = su._successor_buckets(tok)
    assert off[0] == 0 and off[-1] == len(seq) - 1
    for t in range(17):
        want = [tok[i + 1] for i in range(len(seq) - 1) if tok[i] == t]
        assert list(succ[off[t]:off[t + 1]]) == want
    with pytest.raises(AssertionError):
        sho.check_shuffle(seq, seq[::-1] if seq[::-1] != seq else seq + "A", 2)


def test_melt_binned_equals_crosstab_of_hits():
    """The dense-histogram route of the binned counts (device smg_bin_hits -> scan._melt_binned) builds the same frame
    as the reference's per-hit pd.cut + crosstab + melt (scan.py:77-111), including which rows / columns exist."""
    import pandas as pd

    from smeagol_b200 import scan as sscan

    class M:
        pass

    rng = np.random.default_rng(5)
    model = M()
    model.channels = 7
    model.Matrix_ids = np.array(["m%d" % (i % 5) for i in range(7)], dtype=object)
    model.widths = np.array([5, 7, 5, 9, 6, 8, 5])
    model.max_scores = rng.uniform(5, 12, size=7)

    class E:
        pass

    enc = E()
    enc.ids = np.array(["a", "b", "a", "c", "b", "d"], dtype=object)
    enc.names = np.array(["a", "b", "a", "c", "b", "d"], dtype=object)
    enc.senses = np.array(["+", "+", "-", "+", "-", "+"], dtype=object)
    bins = np.array([0.5, 0.6, 0.75, 0.9])
    edges = np.concatenate([bins, [1.0]])
    for n_hits, motifs in ((400, range(7)), (12, (1, 4)), (0, ())):
        seq_idx = rng.integers(0, 5, size=n_hits)          # row 5 ('d') never hit
        pwm_idx = rng.choice(list(motifs), size=n_hits) if n_hits else np.zeros(0, dtype=np.int64)
        frac = rng.uniform(0.5001, 0.8999, size=n_hits)    # last bin never observed
        scores = (frac * model.max_scores[pwm_idx]).astype(np.float32)
        frac_used = scores / model.max_scores[pwm_idx]
        hist = np.zeros((6, 7, 4), dtype=np.int64)
        b = np.searchsorted(edges, frac_used, side="left") - 1
        ok = (frac_used > edges[0]) & (frac_used <= edges[-1])
        np.add.at(hist, (seq_idx[ok], pwm_idx[ok], b[ok]), 1)
        got = sscan._melt_binned(hist, (enc.ids, enc.names, enc.senses), model, edges)
        if n_hits == 0:
            assert len(got) == 0 and list(got.columns) == ["Matrix_id", "width", "id", "name", "sense", "bin", "num"]
            continue
        want = sscan._bin_sites_by_score(enc, model, (seq_idx, np.zeros(n_hits, dtype=np.int64), pwm_idx), scores, bins)
        assert list(got.columns) == list(want.columns) and len(got) == len(want)
        for col in ("Matrix_id", "width", "id", "name", "sense", "num"):
            assert list(got[col]) == list(want[col]), col
        assert [str(x) for x in got["bin"]] == [str(x) for x in want["bin"]]
        assert list(got["bin"].cat.categories) == list(want["bin"].cat.categories)


def test_melt_routes_with_inexact_edges_and_duplicate_pwm_keys():
    """(1) bin edges that are not exactly representable (np.arange(0.5, 1.0, 0.1), the reference's standard grid): the
    dense route must carry pd.cut's own interval labels ((0.7, 0.8], not (0.7, 0.7999999999999999]); (2) PWMs sharing
    (Matrix_id, width) collapse into one summed crosstab row in counts and binned counts (scan.py:137-140, :100-111)."""
    import pandas as pd

    from smeagol_b200 import scan as sscan

    class M:
        pass

    rng = np.random.default_rng(11)
    model = M()
    model.channels = 6
    model.Matrix_ids = np.array(["m0", "m1", "m0", "m2", "m1", "m0"], dtype=object)
    model.widths = np.array([5, 7, 5, 9, 7, 6])                 # keys: (m0,5) x2, (m1,7) x2, (m2,9), (m0,6)
    model.max_scores = rng.uniform(5, 12, size=6)

    class E:
        pass

    enc = E()
    enc.ids = np.array(["a", "b", "a", "c"], dtype=object)
    enc.names = np.array(["a", "b", "a", "c"], dtype=object)
    enc.senses = np.array(["+", "+", "-", "+"], dtype=object)
    bins = np.arange(0.5, 1.0, 0.1)
    edges = np.concatenate([bins, [1.0]])
    n_hits = 300
    seq_idx = rng.integers(0, 4, size=n_hits)
    pwm_idx = rng.integers(0, 6, size=n_hits)
    frac = rng.uniform(0.5001, 0.9999, size=n_hits)
    scores = (frac * model.max_scores[pwm_idx]).astype(np.float32)
    frac_used = scores / model.max_scores[pwm_idx]
    hist = np.zeros((4, 6, len(bins)), dtype=np.int64)
    b = np.searchsorted(edges, frac_used, side="left") - 1
    ok = (frac_used > edges[0]) & (frac_used <= edges[-1])
    np.add.at(hist, (seq_idx[ok], pwm_idx[ok], b[ok]), 1)
    got = sscan._melt_binned(hist, (enc.ids, enc.names, enc.senses), model, edges)
    want = sscan._bin_sites_by_score(enc, model, (seq_idx, np.zeros(n_hits, dtype=np.int64), pwm_idx), scores, bins)
    assert got["bin"].cat.categories.equals(want["bin"].cat.categories)
    assert "(0.7, 0.8]" in [str(c) for c in got["bin"].cat.categories]
    assert len(got) == len(want)
    for col in ("Matrix_id", "width", "id", "name", "sense", "num"):
        assert list(got[col]) == list(want[col]), col
    assert [str(x) for x in got["bin"]] == [str(x) for x in want["bin"]]
    # counts: the reference's crosstab over the same hits
    table = np.zeros((4, 6), dtype=np.int64)
    np.add.at(table, (seq_idx, pwm_idx), 1)
    gotc = sscan._melt_counts(table, (enc.ids, enc.names, enc.senses), model)
    ct = pd.crosstab(index=[model.Matrix_ids[pwm_idx], model.widths[pwm_idx]],
                     columns=[enc.ids[seq_idx], enc.names[seq_idx], enc.senses[seq_idx]])
    wantc = ct.melt(ignore_index=False).reset_index()
    wantc.columns = ["Matrix_id", "width", "id", "name", "sense", "num"]
    assert len(gotc) == len(wantc) == 4 * 4                      # 4 distinct (Matrix_id, width) keys x 4 label columns
    for col in wantc.columns:
        assert list(gotc[col]) == list(wantc[col]), col
    # many groups in ONE frame (_counts_frame) == the per-group frames concatenated, incl. a group without any hit
    groups = [(table[:2], (enc.ids[:2], enc.names[:2], enc.senses[:2])), (np.zeros((1, 6), dtype=np.int64), (["z"], ["z"], ["+"])),
              (table[2:], (enc.ids[2:], enc.names[2:], enc.senses[2:]))]
    one = sscan._counts_frame([sscan._melt_counts_parts(t, lab, model) for t, lab in groups], model)
    per = pd.concat([sscan._melt_counts(t, lab, model) for t, lab in groups]).reset_index(drop=True)
    assert list(one.columns) == list(per.columns) and len(one) == len(per) > 0
    for col in per.columns:
        assert list(one[col]) == list(per[col]), col
    empty = sscan._counts_frame([None], model)
    assert list(empty.columns) == ["Matrix_id", "width", "id", "name", "sense", "num"] and len(empty) == 0


def test_device_seq_behaves_like_a_seq():
    """seqrecord.DeviceSeq (sequences that live in a device ASCII buffer) keeps the Seq behaviour the reference API
    relies on; exercised here on a CPU tensor, the buffer type does not matter for the host-side logic."""
    import torch

    import smeagol_b200 as sb
    from smeagol_b200.seqrecord import DeviceSeq
    from smeagol_b200.utils import seq_str

    flat = torch.from_numpy(np.frombuffer(b"NNACGTTGCAAGNN", dtype=np.uint8).copy())
    s = DeviceSeq(flat, 2, 10)
    assert len(s) == 10 and str(s) == "ACGTTGCAAG" and s[0] == "A" and str(s[2:5]) == "GTT"
    assert str(s.reverse_complement()) == "CTTGCAACGT" and s == sb.Seq("ACGTTGCAAG") and list(s)[:3] == ["A", "C", "G"]
    rec = sb.SeqRecord(s, id="x", name="x")
    assert len(rec) == 10 and seq_str(rec) == "ACGTTGCAAG" and seq_str(rec, keep_device=True) is s
    enc = sb.encode.SeqEncoding([rec, sb.SeqRecord(sb.Seq("ACGTACGTAC"), id="y", name="y")], rcomp="both", sense="+")
    assert enc.len == 10 and enc.seqs.shape == (4, 10) and enc.records[0] is s
    assert list(enc.seqs[0][:4]) == [1, 2, 3, 4]
    with pytest.raises(ValueError):
        sb.encode.SeqEncoding([rec, sb.SeqRecord(sb.Seq("ACG"), id="z", name="z")], rcomp="none", sense="+")


def _golden_f4():
    import json
    import os

    with open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "reference_golden_f4.json")) as f:
        return json.load(f)


def test_window_and_similarity_host_functions_replay_the_reference():
    """Host halves of the SURVEY 8(f) rank-4 rows against outputs recorded from the unmodified reference
    (tests/golden/make_golden_f4.py): matrices.ncorr and its helpers, utils.get_tiling_windows_over_genome,
    scan.count_in_sliding_windows on a sites DataFrame, enrich.enrich_in_sliding_windows (Fisher + FDR)."""
    import pandas as pd

    import smeagol_b200 as sb
    from oracle import frames as of

    G = _golden_f4()
    k = G["ncorr_kat"]
    X, Y = np.array(k["X"]), np.array(k["Y"])
    assert sb.utils._equals(sb.matrices.ncorr(X, Y, min_overlap=3), 0.25069190635147276)  # tests/test_matrices.py:193-196
    assert abs(sb.matrices.ncorr(Y, X, min_overlap=3) - k["result_swapped"]) < 1e-12
    assert list(sb.matrices.align_pms(X, Y)) == [(0, 3, 0, 3), (0, 3, 1, 4), (0, 3, 2, 5)]
    rc = sb.matrices.reverse_complement(Y)
    assert np.array_equal(rc, Y[:, ::-1])  # the reference's function as written: columns swapped, rows kept
    with pytest.raises(ValueError):
        sb.matrices.matrix_correlation(X, Y)
    Wd = G["windows"]
    genome = [sb.SeqRecord(sb.Seq(s), id=i, name=i) for s, i in zip(Wd["seqs"], Wd["ids"])]
    sites = of.scan_sequences(Wd["seqs"], Wd["ids"], Wd["ids"], Wd["pwms"]["Matrix_id"], [np.array(w) for w in Wd["pwms"]["weights"]],
                              Wd["threshold"], sense="+", rcomp="both", outputs=["sites"], score=True)["sites"]
    for case in Wd["cases"]:
        w = sb.utils.get_tiling_windows_over_genome(genome, case["width"], case["shift"]).reset_index(drop=True)
        for c in ("id", "start", "end"):
            assert list(w[c]) == list(case["windows"][c])
        cnt = sb.scan.count_in_sliding_windows(sites, genome, Wd["matrix_id"], case["width"], case["shift"])
        assert list(cnt.columns) == ["id", "start", "end", "count"] and list(cnt["count"]) == list(case["counts"]["count"])
        enr = sb.enrich.enrich_in_sliding_windows(sites, genome, case["width"], case["shift"], matrix_id=Wd["matrix_id"])
        exp = case["enrich"]
        assert list(enr.columns) == list(exp.keys())
        for c in ("id", "start", "end", "len", "count", "tot_count"):
            assert list(enr[c]) == list(exp[c]), c
        for c in ("expected", "odds", "p", "padj"):
            np.testing.assert_allclose(enr[c].astype(float), np.array(exp[c], dtype=float), rtol=1e-9, atol=1e-12, err_msg=c)
        one = sb.enrich.enrich_in_window(w.iloc[1], sites, genome, matrix_id=Wd["matrix_id"])
        assert one["count"] == exp["count"][1] and abs(float(one["p"]) - exp["p"][1]) < 1e-12


def test_choice_of_the_per_call_index_engines():
    """device.choose_index_engines (host logic of ScanPlan): which width classes the k-mer index takes (4^W against the
    scan length) and whether the wide PWMs go through the seed engine or the tensor-core prefilter (cost model)."""
    from smeagol_b200 import synth

    w4 = np.array(synth.jaspar_widths(1000))                     # c4: W 6..24
    assert dev.choose_index_engines(w4, 7, 248_000_000) == (12, 12, 12)          # whole chromosome on one GPU: 12-base seeds
    wk, kw, sl = dev.choose_index_engines(w4, 7, 31_000_000)                     # one eighth of it: a cheaper 11-base seed index
    assert (wk, kw) == (12, 12) and sl in (11, 12)
    assert dev.choose_index_engines(w4, 7, 30_000) == (7, 7, 0)                  # short scan: no seeds, classes up to 4^7 <= 45,000
    w2 = np.array([7, 8, 9, 10, 11, 12] * 100)                                   # c2-like: W 7..12, 10 Mb
    assert dev.choose_index_engines(w2, 7, 10_000_000) == (11, 11, 11)           # W=12 PWMs seeded with 11-mers
    assert dev.choose_index_engines(np.full(1024, 16), 7, 4_000_000)[2] == 0     # 4 Mb x 1,024 wide PWMs: the prefilter is cheaper
    assert dev.choose_index_engines(np.full(8, 30), 7, 248_000_000)[2] == 0      # W = 30 > seed + 14: not seedable
    assert dev.choose_index_engines(w4, 7, 248_000_000, allow_seed=False)[2] == 0
    assert dev.choose_index_engines(w4, 7, 1000, forced_kmer="12", forced_seed="12") == (12, 12, 12)  # test switches
    assert dev.choose_index_engines(w4, 7, 248_000_000, forced_kmer="0") == (0, 0, 0)
