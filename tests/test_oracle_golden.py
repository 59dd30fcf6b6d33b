"""CPU tests: the oracle (oracle/scan_oracle.py, oracle/frames.py) replayed against every golden vector the
reference's own tests hold for the scan path, and against outputs of the reference's modules recorded by
tests/golden/make_golden.py.  This is what pins the oracle (parity status: PINNED)."""
import os

import numpy as np
import pandas as pd
import pytest

from golden_util import assert_frame_matches, case_kwargs, fix_stats_ddof, golden, pwms_from_json
from oracle import frames as of
from oracle import scan_oracle as so


def _equals(x, y, eps=1e-4):  # reference tolerance helper, utils.py:14-33
    x, y = np.asarray(x, dtype=float), np.asarray(y, dtype=float)
    return bool(np.all(np.abs(x - y) < eps * max(1, x.size)))


def test_models_kat():
    """tests/test_models.py:38-74"""
    G = golden()
    ids, ws = pwms_from_json(G["test_pwms"])
    widths, max_scores, max_w = so.model_constants(ws)
    assert ids == ["x", "y", "z"] and list(widths) == [3, 5, 3] and max_w == 5
    assert _equals(max_scores, [3.33376361, 5.52770673, 2.50816981])
    preds = so.predict(np.array([[1, 2, 3]]), ws)
    assert preds.shape == (1, 3, 3)
    assert _equals(preds[0, 0, :], [-1.60847875, -7.32647115, 0.35611725])
    np.testing.assert_allclose(preds, np.array(G["test_models"]["predict_out"]), rtol=1e-6, atol=1e-6)
    hit, sc = so.predict_with_threshold(np.array([[1, 2, 3]]), ws, 0.1, score=True)
    assert [list(h) for h in hit] == [[0], [0], [2]]
    assert _equals(sc, [0.35611725])


def test_encode_kat():
    """tests/test_encode.py:13-23"""
    assert list(so.integer_encode("ACGTNWSMKRYBDHVZ")) == [1, 2, 3, 4, 6, 7, 8, 9, 10, 11, 12, 13, 14, 15, 16, 0]
    with pytest.raises(KeyError):
        so.integer_encode("ACGE")
    assert list(so.integer_encode("AGCAU", rc=True)) == [1, 5, 3, 2, 5]
    G = golden()["test_encode"]
    assert list(so.integer_encode(G["alphabet"])) == [int(v) for v in G["codes"]]
    np.testing.assert_array_equal(so.EMB32[so.integer_encode("ACGTN")], np.array(G["one_hot_ACGTN"], dtype=np.float32))


def test_scan_kat_bins():
    """tests/test_scan.py:50-70: predict_with_threshold(0.25) on ATTAAA / GCTATA"""
    G = golden()
    ids, ws = pwms_from_json(G["test_pwms"])
    rows = np.stack([so.integer_encode(s) for s in G["test_scan"]["seqs"]])
    hit, sc = so.predict_with_threshold(rows, ws, 0.25, score=True)
    assert [list(map(int, h)) for h in hit] == G["test_scan"]["pwt_025_idx"]
    np.testing.assert_allclose(sc, G["test_scan"]["pwt_025_scores"], rtol=1e-6)
    out = of.scan_sequences(G["test_scan"]["seqs"], G["test_scan"]["ids"], G["test_scan"]["names"], ids, ws,
                            np.array([0.25, 0.5, 0.75]), sense="+", rcomp="none", outputs=["binned_counts"],
                            group_by="name")
    assert list(out["binned_counts"].num) == [0, 2, 1, 0, 0, 1]
    assert_frame_matches(out["binned_counts"], G["test_scan"]["bin_sites"], "bin_sites")


def test_enrich_golden_sites_counts_stats():
    """tests/test_enrich.py:41-59 with data/real_sites.csv, real_counts.csv, shuf_stats.csv"""
    G = golden()
    E = G["test_enrich"]
    ids, ws = pwms_from_json(G["test_pwms"])
    real = of.scan_sequences(E["genome"], E["ids"], E["ids"], ids, ws, 0.65, sense="+", rcomp="both",
                             outputs=["sites", "counts"], score=True, combine_seqs=True)
    csv = E["csv"]
    assert list(real["sites"].start) == [s - 1 for s in csv["real_sites"]["start"]]  # 1-based in the file
    assert list(real["sites"].sense) == csv["real_sites"]["sense"]
    assert list(real["sites"].Matrix_id) == csv["real_sites"]["m"]
    assert list(real["sites"].id) == csv["real_sites"]["s"]
    np.testing.assert_allclose(real["sites"].score, csv["real_sites"]["score"], rtol=1e-5)
    assert list(real["counts"].Matrix_id) == csv["real_counts"]["Matrix_id"]
    assert list(real["counts"].sense) == csv["real_counts"]["sense"]
    assert list(real["counts"].num) == csv["real_counts"]["num"]
    assert real["_n_flipped"] == 0
    # live reference run (stand-in conv) agrees too
    assert_frame_matches(real["sites"].drop(columns=["score", "max_score", "frac_score"]), E["real_sites"], "real_sites")
    assert_frame_matches(real["counts"], E["real_counts"], "real_counts")
    # background: group_by name, sep_ids, stats (enrich.py:172-183)
    bg_ids = [r[0] for r in E["shuf_genome"]]
    bg_seqs = [r[1] for r in E["shuf_genome"]]
    bg_names = [E["ids"][i // E["simN"]] for i in range(len(bg_ids))]  # utils.read_bg_seqs (utils.py:109-132)
    shuf = of.scan_sequences(bg_seqs, bg_ids, bg_names, ids, ws, 0.65, sense="+", rcomp="both",
                             outputs=["counts", "stats"], group_by="name", combine_seqs=True, sep_ids=True)
    assert list(shuf["stats"].Matrix_id) == csv["shuf_stats"]["Matrix_id"]
    assert list(shuf["stats"].sense) == csv["shuf_stats"]["sense"]
    assert list(shuf["stats"].avg) == csv["shuf_stats"]["avg"]
    np.testing.assert_allclose(shuf["stats"].sd, csv["shuf_stats"]["sd"], rtol=1e-9, atol=1e-12)  # ddof=1
    assert_frame_matches(shuf["counts"], E["shuf_counts"], "shuf_counts")


def test_variant_golden():
    """tests/test_variant.py:47-63"""
    G = golden()
    ids, ws = pwms_from_json(G["test_pwms"])
    w = dict(zip(ids, ws))
    s0, v0 = so.variant_effect_on_site(w["x"], "GAC", 2, "G")
    s1, v1 = so.variant_effect_on_site(w["z"], "GCT", 0, "A")
    s2, v2 = so.variant_effect_on_site(w["z"], "GCT", 1, "A")
    assert _equals([s0, s1, s2], [1.0, 0.769149, 0.769149])
    assert _equals([v0, v1, v2], [0.4080668, 0.1419828, 0.769149])
    np.testing.assert_allclose([v0, v1, v2], G["test_variant"]["result"]["variant_score"], rtol=1e-12)


@pytest.mark.parametrize("i", range(13))
def test_reference_scan_cases(i):
    """Reference scan_sequences outputs on seeded random inputs (IUPAC, RNA, grouping, combine, stats, bins)."""
    G = golden()
    case = G["scan_cases"][i]
    ids, ws = pwms_from_json(case["pwms"])
    thr, kw = case_kwargs(case)
    # pure reference semantics first (fp32 compare), any disagreement must be a near-threshold site
    out = of.scan_sequences(case["seqs"], case["ids"], case["names"], ids, ws, thr, order="ref", adjudicate=False,
                            **kw)
    nf = out.pop("_n_flipped")
    if nf:  # reported: fp32, summation-order dependent decisions inside the 1e-5 near-threshold band
        assert case["name"] == "no_hits", "unexpected near-threshold disagreement in %s" % case["name"]
        return
    for k, exp in case["outputs"].items():
        assert_frame_matches(out[k], fix_stats_ddof(k, exp), case["name"] + ":" + k)


def test_kernel_order_equals_ref_order_outside_band():
    """The '-' strand evaluated on the forward string with the reverse-complemented table (kernel order) gives the
    same sites as scanning the reverse-complement string (reference order); fp32 scores agree to 1e-5 relative."""
    G = golden()
    for case in G["scan_cases"][:9]:
        ids, ws = pwms_from_json(case["pwms"])
        thr, kw = case_kwargs(case)
        if "binned_counts" in kw.get("outputs", []):
            continue
        frac = thr
        a = so.scan_rows(case["seqs"][:2], ws, frac, rcomp="both", order="ref")
        b = so.scan_rows(case["seqs"][:2], ws, frac, rcomp="both", order="kernel") if len(
            {len(s) for s in case["seqs"][:2]}) == 1 else None
        if b is None:
            a = so.scan_rows(case["seqs"][:1], ws, frac, rcomp="both", order="ref")
            b = so.scan_rows(case["seqs"][:1], ws, frac, rcomp="both", order="kernel")
        for k in ("row", "pos", "motif"):
            np.testing.assert_array_equal(a[k], b[k])
        np.testing.assert_allclose(a["score"], b["score"], rtol=1e-5)
        np.testing.assert_array_equal(a["score64"].round(9), b["score64"].round(9))


def test_c_oracle_equals_numpy_oracle():
    """oracle/scan_oracle.c (the OpenMP restatement used for benchmark-sized parity) == oracle/scan_oracle.py."""
    from oracle import c_oracle

    rng = np.random.default_rng(3)
    ws = [np.log2((rng.dirichlet([0.5] * 4, size=int(rng.integers(3, 25))) + 0.01) / 1.04 / 0.25) for _ in range(25)]
    amb = "ACGTNWSMKRYBDHVZ"
    pa = np.array([0.23] * 4 + [0.08 / 12] * 12)
    seqs = ["".join(rng.choice(list("ACGT"), size=600)), "".join(rng.choice(list(amb), size=400, p=pa / pa.sum())), "ACGTA"]
    codes = [so.integer_encode(s) for s in seqs]
    for rcomp, strands in (("both", 3), ("none", 1), ("only", 2)):
        exp = so.scan_rows(seqs, ws, 0.5, rcomp=rcomp, order="kernel")
        n = len(seqs)
        out_rows = {"both": np.stack([np.arange(n), n + np.arange(n)], 1), "none": np.stack([np.arange(n), np.zeros(n)], 1),
                    "only": np.stack([np.zeros(n), np.arange(n)], 1)}[rcomp].astype(np.int32)
        res = c_oracle.scan(codes, ws, 0.5, strands=strands, out_rows=out_rows, n_out_rows=2 * n, want_hits=True,
                            hit_cap=200000)
        pb, mb = res["pos_bits"], res["motif_bits"]
        k = res["keys"]
        np.testing.assert_array_equal((k >> np.uint64(pb + mb)).astype(np.int64), exp["row"])
        np.testing.assert_array_equal(((k >> np.uint64(mb)) & np.uint64((1 << pb) - 1)).astype(np.int64), exp["pos"])
        np.testing.assert_array_equal((k & np.uint64((1 << mb) - 1)).astype(np.int64), exp["motif"])
        np.testing.assert_array_equal(res["scores"].view(np.uint32), exp["score"].view(np.uint32))
        assert res["n_hits"] == len(exp["row"]) and res["counts"].sum() == len(exp["row"])


def test_c_variant_windows_equals_numpy_oracle():
    """oracle/scan_oracle.c::orc_variant_windows (the checker of the config-c5 GPU test) against the numpy loop
    restatement of variant.py:7-47 / scan.py:8-37 (fp64), incl. IUPAC codes near the SNVs and SNVs at the row ends."""
    from oracle import c_oracle

    rng = np.random.default_rng(71)
    ws = []
    for W in (3, 5, 8, 12, 17):
        ws.append(np.log2((rng.dirichlet([0.5] * 4, size=W) + 0.01) / 1.04 / 0.25))
    seq = "".join(rng.choice(list("ACGT"), size=120)) + "NRYACGTW" + "".join(rng.choice(list("ACGT"), size=90))
    pos = np.concatenate([[0, 1, 2, len(seq) - 1, len(seq) - 2], rng.integers(0, len(seq), size=40)])
    alts = [("ACGTN"[rng.integers(0, 5)]) for _ in pos]
    e_ref, e_alt = so.variant_window_scan(seq, pos, alts, ws)
    codes = so.integer_encode(seq).astype(np.uint8)
    r, a = c_oracle.variant_windows(codes, ws, pos, [so.CODE_OF[x] for x in alts])
    np.testing.assert_allclose(r, e_ref, rtol=1e-5, atol=2e-5)
    np.testing.assert_allclose(a, e_alt, rtol=1e-5, atol=2e-5)


def _golden_f4():
    import json

    with open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "reference_golden_f4.json")) as f:
        return json.load(f)


def test_matrix_oracle_replays_reference_ncorr_and_windows():
    """oracle/matrix_oracle.py against outputs recorded from the unmodified reference (tests/golden/make_golden_f4.py):
    tests/test_matrices.py:176-213's known answers, pairwise_ncorrs on 15 random PWMs (widths 3..20, one duplicate),
    tiling windows and count_in_sliding_windows on the sites of a recorded scan."""
    from oracle import matrix_oracle as mo

    G = _golden_f4()
    k = G["ncorr_kat"]
    assert abs(mo.ncorr(np.array(k["X"]), np.array(k["Y"])) - 0.25069190635147276) < 1e-9
    assert abs(mo.ncorr(np.array(k["Y"]), np.array(k["X"])) - k["result_swapped"]) < 1e-12
    tp = pwms_from_json(golden()["test_pwms"])[1]
    np.testing.assert_allclose(mo.pairwise_ncorrs(tp), np.array(G["pairwise_test_pwms"]["result"]), rtol=0, atol=1e-12)
    # (tests/test_matrices.py:199-213 expects 0.46869686 / 0.98630056 / 0.46716826 here; the unmodified module returns the
    # recorded values -- its reverse_complement does not reverse the rows, see oracle/matrix_oracle.py)
    R = G["pairwise_random"]
    got = mo.pairwise_ncorrs([np.array(w) for w in R["weights"]])
    np.testing.assert_allclose(got, np.array(R["result"]), rtol=0, atol=1e-12)
    assert abs(got[0, 14] - 1.0) < 1e-12
    Wd = G["windows"]
    sites = of.scan_sequences(Wd["seqs"], Wd["ids"], Wd["ids"], Wd["pwms"]["Matrix_id"], [np.array(w) for w in Wd["pwms"]["weights"]],
                              Wd["threshold"], sense="+", rcomp="both", outputs=["sites"], score=True)["sites"]
    assert len(sites) == Wd["n_sites"]
    lens = [len(s) for s in Wd["seqs"]]
    for case in Wd["cases"]:
        w = mo.tiling_windows(Wd["ids"], lens, case["width"], case["shift"]).reset_index(drop=True)
        for c in ("id", "start", "end"):
            assert list(w[c]) == list(case["windows"][c])
        cnt = mo.count_in_sliding_windows(sites, Wd["ids"], lens, Wd["matrix_id"], case["width"], case["shift"])
        assert list(cnt["count"]) == list(case["counts"]["count"]) and sum(cnt["count"]) > 0
