"""GPU parity tests on the shapes of BASELINE.json's configurations c3 / c4 / c5 and of the multi-GPU split
(run on the B200 box with -m gpu).  The checker is oracle/c_oracle.py (the plain-C restatement, pinned to the numpy
oracle and through it to the reference's goldens by tests/test_oracle_golden.py); bars as in test_gpu_parity.py:
integer results (hits, positions, counts, checksums of the hit keys) identical, fp32 scores within 1e-5 relative.
"""
import os
import socket

import numpy as np
import pandas as pd
import pytest

pytestmark = pytest.mark.gpu

_LUT = np.zeros(256, dtype=np.uint8)
_LUT[list(b"ACGTUN")] = [1, 2, 3, 4, 5, 6]


def _c4_slice(L=1_200_000):
    """c4's generator (synth.single_sequence seed 4, JASPAR-sized widths seed 4) on a 1.2 Mb slice, with a stretch of
    guaranteed hits planted across the first segment boundary: the consensus sequences of 60 PWMs back to back."""
    from smeagol_b200 import synth

    pw = synth.synth_pwms(1000, widths=synth.jaspar_widths(1000, seed=4))
    seq = synth.single_sequence(L, 4).copy()
    cut = min(399_984, (L // 2) // 16 * 16)  # a multiple of 16: first segment boundary
    cons = b"".join(bytes(b"ACGT"[i] for i in np.argmax(w, axis=1)) for w in list(pw.weights)[:60])
    cons = np.frombuffer(cons, dtype=np.uint8)
    seq[cut - len(cons) // 2:cut - len(cons) // 2 + len(cons)] = cons
    return pw, seq, cut


def _scan_parts(model, seq, cuts):
    """Scan [cuts[i], cuts[i+1]) as separate segment rows with halo; returns (ScanResult, total counts (M, 2))."""
    from smeagol_b200 import device as dev

    L, halo = len(seq), int(model.max_width) - 1
    parts = [seq[a:min(L, b + halo)] for a, b in zip(cuts[:-1], cuts[1:])]
    segs = [(a, L, b - a) for a, b in zip(cuts[:-1], cuts[1:])]
    batch = dev.SeqBatch(parts, segments=segs)
    out_rows = np.array([[0, 1]] * len(parts), dtype=np.int32)
    res = dev.scan(model.device_set, batch, model.thresholds(0.7), 3, out_rows, 2)
    return res, res.counts.sum(dim=0).cpu().numpy().astype(np.int64)


def test_c4_slice_segments_with_halo_vs_c_oracle():
    """Config c4 shape: 1.2 Mb of the c4 generator x the 1,000 JASPAR-sized PWMs (W 6-24) at thr 0.7, scanned as THREE
    segments with a (W_max - 1)-base halo, one boundary inside a hit-dense stretch; all three engines take part (k-mer
    index, register tables, tcgen05 prefilter).  Count table, hit count and the sum / xor checksums of every hit key
    (global coordinates, reverse-complement positions from the global length) against the C oracle on the whole row."""
    import smeagol_b200 as sb
    from oracle import c_oracle

    pw, seq, cut = _c4_slice()
    model = sb.models.PWMModel(pw)
    res, counts = _scan_parts(model, seq, [0, cut, 800_000, len(seq)])
    assert res.stats["engine"] == "auto" and res.stats["kmer_max_width"] >= 9 and res.stats["n_candidates"] > 0
    exp = c_oracle.scan([_LUT[seq]], list(pw.weights), 0.7, strands=3, out_rows=np.array([[0, 1]]), n_out_rows=2)
    assert (res.pos_bits, res.motif_bits) == (exp["pos_bits"], exp["motif_bits"])
    np.testing.assert_array_equal(counts, exp["counts"][0])
    assert res.n_hits == exp["n_hits"] and res.stats["n_hits_total"] == exp["n_hits"]
    ksum, kxor = res.key_checksum()
    assert ksum % (1 << 64) == exp["key_sum"] and kxor % (1 << 64) == exp["key_xor"]
    # the planted stretch really is hit dense, and it straddles the boundary
    row, pos, motif, sc = res.hits_host()
    fwd = pos[row == 0]
    dense = np.sum((fwd >= cut - 300) & (fwd < cut + 300))
    assert dense > 600 * len(fwd) / len(seq)  # denser than the random background (0.67 hits per start and strand)
    assert np.any((fwd >= cut - 24) & (fwd < cut)) and np.any((fwd >= cut) & (fwd < cut + 24))
    assert np.all(np.diff((row << 50) | (pos << 12) | motif) > 0)  # sorted, unique


def test_rank_partitions_equal_single_gpu():
    """N-rank parity of the c4 split, emulated as partitions on ONE GPU (B200_PROFILING.md: with fewer GPUs than ranks,
    run the ranks' data through the same kernels one after the other): for world sizes 2, 3 and 8 the ranges of
    dist.shard_sequence are scanned separately; the summed count tables, hit counts and key checksums must equal the
    single-range scan, and the concatenation of the per-rank sorted hit lists must be the globally sorted list."""
    import smeagol_b200 as sb
    from smeagol_b200 import dist as sdist

    pw, seq, cut = _c4_slice(400_000 + 64)
    pw = pw.iloc[::4].reset_index(drop=True)  # 250 PWMs: every width class stays present
    model = sb.models.PWMModel(pw)
    L, halo = len(seq), int(model.max_width) - 1
    whole, counts = _scan_parts(model, seq, [0, L])
    w_row, w_pos, w_motif, w_sc = whole.hits_host()
    for world in (2, 3, 8):
        tot = np.zeros_like(counts)
        n_hits, ksum, kxor = 0, 0, 0
        fwd_lists, rc_lists = [], []
        for rank in range(world):
            a, b, e = sdist.shard_sequence(L, world, halo)[rank]
            slices, segs = sdist.segment_rows(a, b, e, L, seg_len=50_000, halo=halo)
            from smeagol_b200 import device as dev

            batch = dev.SeqBatch([seq[lo:hi] for lo, hi in slices], segments=segs)
            out_rows = np.array([[0, 1]] * len(segs), dtype=np.int32)
            res = dev.scan(model.device_set, batch, model.thresholds(0.7), 3, out_rows, 2)
            tot += res.counts.sum(dim=0).cpu().numpy().astype(np.int64)
            n_hits += res.n_hits
            s, x = res.key_checksum()
            ksum, kxor = (ksum + s) % (1 << 64), kxor ^ x
            r, p, m, sc = res.hits_host()
            fwd_lists.append((p[r == 0], m[r == 0], sc[r == 0]))
            rc_lists.append((p[r == 1], m[r == 1], sc[r == 1]))
        np.testing.assert_array_equal(tot, counts)
        ws, wx = whole.key_checksum()
        assert n_hits == whole.n_hits and ksum == ws % (1 << 64) and kxor == wx
        # forward hits: rank order is position order; reverse-complement coordinates run backwards over the ranks
        for lists, sel in ((fwd_lists, w_row == 0), (rc_lists[::-1], w_row == 1)):
            np.testing.assert_array_equal(np.concatenate([t[0] for t in lists]), w_pos[sel])
            np.testing.assert_array_equal(np.concatenate([t[1] for t in lists]), w_motif[sel])
            np.testing.assert_array_equal(np.concatenate([t[2] for t in lists]).view(np.uint32), w_sc[sel].view(np.uint32))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _nccl_worker(rank, world, port, q):
    import torch
    import torch.distributed as dist

    import smeagol_b200 as sb
    from smeagol_b200 import dist as sdist

    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    pw, seq, cut = _c4_slice(400_000 + 64)
    pw = pw.iloc[::4].reset_index(drop=True)
    model = sb.models.PWMModel(pw)
    res, counts = sdist.scan_long_sequence(torch.from_numpy(seq).pin_memory(), model, 0.7, seg_len=50_000)
    s, x = res.key_checksum()
    t = torch.tensor([res.n_hits, s & ((1 << 62) - 1)], dtype=torch.int64, device="cuda")
    dist.all_reduce(t)
    xs = [None] * world
    dist.all_gather_object(xs, x)
    if rank == 0:
        kx = 0
        for v in xs:
            kx ^= v
        q.put((int(t[0]), int(t[1]) & ((1 << 62) - 1), kx, counts.cpu().numpy().astype(np.int64)))
    dist.destroy_process_group()


def test_two_nccl_ranks_equal_single_gpu():
    """The same invariance over real NCCL ranks (dist.scan_long_sequence on 2 GPUs: halo sharding, rank-local hits,
    ONE all_reduce of the count table); skipped on a one-GPU box, where the emulated test above stands in."""
    import torch
    import torch.multiprocessing as mp

    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs (the partition test covers the same split on one)")
    import smeagol_b200 as sb

    pw, seq, cut = _c4_slice(400_000 + 64)
    pw = pw.iloc[::4].reset_index(drop=True)
    model = sb.models.PWMModel(pw)
    whole, counts = _scan_parts(model, seq, [0, len(seq)])
    ws, wx = whole.key_checksum()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_nccl_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    n_hits, ksum, kxor, cnt = q.get(timeout=300)
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    assert n_hits == whole.n_hits and kxor == wx
    assert ksum == ws & ((1 << 62) - 1)  # the ranks' sums are reduced modulo 2^62
    np.testing.assert_array_equal(cnt[0], counts)


def test_c3_shuffled_backgrounds_counts_and_stats():
    """Config c3 shape: 64 dinucleotide-preserving shuffles of a 30 kb genome generated on the device
    (utils.shuffle_records -> smg_kmer_shuffle) x 500 PWMs through the PUBLIC call enrich_in_genome makes for its
    background (enrich.py:172-183): scan_sequences(group_by='name', combine_seqs=True, sep_ids=True,
    outputs=['counts', 'stats']).  Expected frames are built from the C oracle's count table with the reference's
    crosstab / groupby semantics (scan.py:137-146, :268-310; std with ddof=1 as in the golden shuf_stats.csv)."""
    import smeagol_b200 as sb
    from oracle import c_oracle
    from smeagol_b200 import synth

    pw = synth.synth_pwms(500, 7, 12)
    genome = synth.single_sequence(30_000, 3)
    rec = sb.SeqRecord(sb.Seq(genome.tobytes().decode()), id="virus", name="virus")
    bg = sb.utils.shuffle_records([rec], 64, 2, seed=3)
    assert len(bg) == 64 and bg[0].id == "background_seq_0" and bg[5].name == "virus"
    model = sb.models.PWMModel(pw)
    out = sb.scan.scan_sequences(bg, model, 0.7, sense="+", rcomp="both", outputs=["counts", "stats"], group_by="name",
                                 combine_seqs=True, sep_ids=True)
    texts = [str(r.seq) for r in bg]
    assert all(len(t) == 30_000 for t in texts) and len(set(texts)) == 64
    exp = c_oracle.scan([_LUT[np.frombuffer(t.encode(), dtype=np.uint8)] for t in texts], list(pw.weights), 0.7, strands=3)
    cnt = exp["counts"]  # (64, 500, 2)
    assert np.all(cnt.sum(axis=(0, 2)) > 0)  # every PWM has a hit somewhere: the crosstab rectangle is complete
    ids = np.array([r.id for r in bg], dtype=object)
    widths = np.array([w.shape[0] for w in pw.weights])
    rows = []
    for st, sense in ((0, "+"), (1, "-")):
        for i in range(64):
            rows.append(pd.DataFrame({"Matrix_id": np.asarray(pw.Matrix_id, dtype=object), "width": widths, "sense": sense,
                                      "id": ids[i], "num": cnt[i, :, st]}))
    full = pd.concat(rows)
    exp_counts = full.groupby(["Matrix_id", "width", "sense", "id"]).num.sum().reset_index()
    g = exp_counts.groupby(["Matrix_id", "width", "sense"]).num
    exp_stats = pd.DataFrame({"len": g.size(), "avg": g.mean(), "sd": g.std(ddof=1)}).reset_index()
    got_c, got_s = out["counts"], out["stats"]
    assert list(got_c.columns) == list(exp_counts.columns) and len(got_c) == len(exp_counts) == 500 * 2 * 64
    for c in exp_counts.columns:
        assert list(got_c[c]) == list(exp_counts[c]), c
    assert list(got_s.columns) == list(exp_stats.columns) and len(got_s) == 1000
    for c in ("Matrix_id", "width", "sense", "len"):
        assert list(got_s[c]) == list(exp_stats[c]), c
    np.testing.assert_allclose(got_s["avg"], exp_stats["avg"], rtol=1e-12)
    np.testing.assert_allclose(got_s["sd"], exp_stats["sd"], rtol=1e-9)


@pytest.mark.parametrize("exact", [False, True])
def test_c5_variant_windows_10k_snvs_vs_c_oracle(exact):
    """Config c5 shape: 10,000 SNVs from the c5 generator (synth.snvs) on a 200 kb reference x 500 PWMs (W 7-12):
    maximum reference / alternate window score over the covering windows of both strands, against the C oracle.
    Reference maxima: same fp32 summation order, bit-identical.  Alternate maxima: bit-identical in the exact (forked
    sums) mode; in the default delta mode (ref - T[k, ref] + T[k, alt]) within 1e-5 of the PWM's score scale
    sum_k max_b |T[k, b]| -- the tolerance north_star states for floating-point scores."""
    import smeagol_b200 as sb
    from oracle import c_oracle
    from smeagol_b200 import device as dev
    from smeagol_b200 import synth, variant

    pw = synth.synth_pwms(500, 7, 12)
    seq = synth.single_sequence(200_000, 5)
    pos, alt = synth.snvs(seq, 10_000, 12, seed=5)
    model = sb.models.PWMModel(pw)
    batch = dev.SeqBatch([seq])
    ref_max, alt_max = variant.variant_window_scores_on_batch(batch, np.zeros(len(pos), dtype=np.int32), pos, alt, model,
                                                              exact=exact)
    e_ref, e_alt = c_oracle.variant_windows(_LUT[seq], list(pw.weights), pos, alt)
    assert np.array_equal(ref_max, e_ref)
    if exact:
        assert np.array_equal(alt_max, e_alt)
    else:
        scale = np.array([np.abs(np.asarray(w)).max(axis=1).sum() for w in pw.weights], dtype=np.float64)
        assert np.all(np.abs(alt_max.astype(np.float64) - e_alt) <= 1e-5 * scale[None, :])
        assert np.mean(alt_max == e_alt) > 0.2  # many windows still round the same way
    # the delta the config reports: an SNV changes at least one PWM's best window for nearly every SNV
    assert np.mean(np.any(alt_max != ref_max, axis=1)) > 0.99
    # the [n_snv][n_motifs] layout of the C ABI holds the same numbers
    r2, a2 = variant.variant_window_scores_on_batch(batch, np.zeros(len(pos), dtype=np.int32), pos, alt, model, exact=exact,
                                                    motif_major=False)
    assert r2.shape == ref_max.shape and np.array_equal(r2, ref_max) and np.array_equal(a2, alt_max)


def test_shuffle_fixture_of_the_reference():
    """tests/test_utils.py:15-34 of the reference, replayed on the device shuffle: its fixture genome, its call
    shuffle_records(genome, 3, 2) and its assertions (6 records; every copy keeps its source's dinucleotide counts),
    plus the naming the reference gives the copies (utils.py:68-71)."""
    from collections import Counter

    import smeagol_b200 as sb

    genome = [sb.SeqRecord(sb.Seq("AGAGCGCATTTCCTACGCATGCTCGATCAAATGCTACGGATTCTAAAA"), id="seg1", name="seg1"),
              sb.SeqRecord(sb.Seq("CCCGGACGTTTGCTCGAGGG"), id="seg2", name="seg2")]

    def dinuc_freqs(x):
        return Counter(sorted([x[i:i + 2] for i in range(len(x) - 1)]))

    ref_freqs = [dinuc_freqs(str(x.seq)) for x in genome]
    shuf = sb.utils.shuffle_records(genome, 3, 2)
    assert len(shuf) == 6
    shuf_freqs = [dinuc_freqs(str(x.seq)) for x in shuf]
    assert shuf_freqs[0] == shuf_freqs[1] == shuf_freqs[2] == ref_freqs[0]
    assert shuf_freqs[3] == shuf_freqs[4] == shuf_freqs[5] == ref_freqs[1]
    assert [r.id for r in shuf] == ["background_seq_0", "background_seq_1", "background_seq_2"] * 2
    assert [r.name for r in shuf] == ["seg1"] * 3 + ["seg2"] * 3
    # k = 1: a permutation of the bases
    shuf1 = sb.utils.shuffle_records(genome, 2, 1)
    assert sorted(str(shuf1[0].seq)) == sorted(str(genome[0].seq)) and sorted(str(shuf1[3].seq)) == sorted(str(genome[1].seq))


def _golden_f4():
    import json

    with open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "reference_golden_f4.json")) as f:
        return json.load(f)


def test_pairwise_ncorrs_device_replays_the_reference():
    """matrices.pairwise_ncorrs on the device (smg_pairwise_ncorrs, one thread per pair, fp64) against
    outputs recorded from the unmodified reference (test PWMs; 15 random PWMs of widths 3..20 incl. a duplicate) and
    against the CPU oracle on 120 random PWMs."""
    import smeagol_b200 as sb
    from golden_util import golden, pwms_from_json
    from oracle import matrix_oracle as mo

    G = _golden_f4()
    tp = pwms_from_json(golden()["test_pwms"])[1]
    np.testing.assert_allclose(sb.matrices.pairwise_ncorrs(tp), np.array(G["pairwise_test_pwms"]["result"]), rtol=0, atol=1e-12)
    R = G["pairwise_random"]
    got = sb.matrices.pairwise_ncorrs([np.array(w) for w in R["weights"]])
    np.testing.assert_allclose(got, np.array(R["result"]), rtol=0, atol=1e-12)
    assert got[0, 14] == pytest.approx(1.0, abs=1e-12) and np.all(np.diag(got) == 1.0) and np.array_equal(got, got.T)
    rng = np.random.default_rng(91)
    mats = [np.log2((rng.dirichlet([0.5] * 4, size=int(rng.integers(1, 25))) + 0.01) / 1.04 / 0.25) for _ in range(120)]
    mats[7] = np.zeros((4, 4))  # zero variance: NaN correlations, Python max() semantics
    with np.errstate(invalid="ignore", divide="ignore"):
        exp = mo.pairwise_ncorrs(mats)
    got = sb.matrices.pairwise_ncorrs(mats)
    np.testing.assert_allclose(got, exp, rtol=0, atol=1e-12, equal_nan=True)
    assert np.isnan(exp[7, 3])


def test_window_counts_and_thresholds_on_device_replay_the_reference(monkeypatch):
    """scan.count_in_sliding_windows / enrich.enrich_in_sliding_windows on sites that never leave the device
    (scan.scan_sequences_device + smg_window_counts), and enrich.examine_thresholds (device binned counts), against
    outputs recorded from the unmodified reference (tests/golden/make_golden_f4.py)."""
    import smeagol_b200 as sb
    from golden_util import assert_frame_matches

    G = _golden_f4()
    Wd = G["windows"]
    genome = [sb.SeqRecord(sb.Seq(s), id=i, name=i) for s, i in zip(Wd["seqs"], Wd["ids"])]
    pw = pd.DataFrame({"Matrix_id": Wd["pwms"]["Matrix_id"], "weights": [np.array(w) for w in Wd["pwms"]["weights"]]})
    model = sb.models.PWMModel(pw)
    dsites = sb.scan.scan_sequences_device(genome, model, Wd["threshold"], sense="+", rcomp="both")
    assert len(dsites) == Wd["n_sites"]
    for case in Wd["cases"]:
        cnt = sb.scan.count_in_sliding_windows(dsites, genome, Wd["matrix_id"], case["width"], case["shift"])
        for c in ("id", "start", "end", "count"):
            assert list(cnt[c]) == list(case["counts"][c]), c
        enr = sb.enrich.enrich_in_sliding_windows(dsites, genome, case["width"], case["shift"], matrix_id=Wd["matrix_id"])
        exp = case["enrich"]
        assert list(enr.columns) == list(exp.keys())
        for c in ("id", "start", "end", "len", "count", "tot_count"):
            assert list(enr[c]) == list(exp[c]), c
        for c in ("expected", "odds", "p", "padj"):
            np.testing.assert_allclose(enr[c].astype(float), np.array(exp[c], dtype=float), rtol=1e-9, atol=1e-12, err_msg=c)
    # every PWM at once (matrix_id=None) against the host route on the DataFrame of the same scan
    sites = sb.scan.scan_sequences(genome, model, Wd["threshold"], sense="+", rcomp="both", outputs=["sites"])["sites"]
    a = sb.scan.count_in_sliding_windows(dsites, genome, None, 150, 50)
    b = sb.scan.count_in_sliding_windows(sites, genome, None, 150, 50)
    assert list(a["count"]) == list(b["count"]) and a["count"].sum() > len(sites)
    # examine_thresholds: the recorded run used reversed records as its "shuffles" (deterministic stand-in)
    def fake_shuffle(records, simN, simK=2, **kw):
        return [sb.SeqRecord(sb.Seq(str(r.seq)[::-1]), id="background_seq_%d" % i, name=r.id) for r in records for i in range(simN)]

    monkeypatch.setattr(sb.enrich, "shuffle_records", fake_shuffle)
    # A perfect match has frac_score = fp32 score / fp64 max_score = 1 +- 1 ulp: whether it is inside the last bin
    # (0.9, 1.0] depends on the fp32 summation order, which the reference leaves to its conv backend.  Cells of that bin
    # may therefore differ by at most the number of such sites (north_star: scores within 1e-5); everything else is exact.
    def perfect_sites(recs, keys, **kw):
        st = sb.scan.scan_sequences(recs, model, 0.5, sense="+", rcomp="both", outputs=["sites"], score=True, **kw)["sites"]
        st = st[np.abs(st.frac_score - 1.0) < 1e-6]
        return st.groupby(keys).size().to_dict() if len(st) else {}

    def check_binned(got, exp, name, slack):
        exp = pd.DataFrame(exp)
        got = got.copy()
        got["bin"] = got["bin"].astype(str)
        assert list(got.columns) == list(exp.columns) and len(got) == len(exp), name
        keys = [c for c in exp.columns if c not in ("num", "bin")]
        for c in exp.columns:
            if c != "num":
                assert list(got[c]) == list(exp[c]), (name, c)
        diff = np.flatnonzero(got["num"].to_numpy() != exp["num"].to_numpy())
        for i in diff:
            assert got["bin"].iloc[i] == "(0.9, 1.0]", (name, got.iloc[i].to_dict())
            k = tuple(got[c].iloc[i] for c in slack[0])
            assert abs(int(got["num"].iloc[i]) - int(exp["num"].iloc[i])) <= slack[1].get(k if len(k) > 1 else k[0], 0), (name, k)
        return len(diff)

    n_diff = 0
    for combine in (True, False):
        res = sb.enrich.examine_thresholds(genome, model, simN=2, simK=2, rcomp="both", min_threshold=0.5, threshold_shift=0.1,
                                           combine_seqs=combine)
        exp = G["examine_thresholds_combine_%s" % combine]
        rkeys = ["Matrix_id", "sense"] + ([] if combine else ["id"])
        skeys = ["Matrix_id", "sense", "id"] + ([] if combine else ["name"])
        n_diff += check_binned(res["real_binned"], exp["real_binned"], "real_binned", (rkeys, perfect_sites(genome, rkeys)))
        n_diff += check_binned(res["shuf_binned"], exp["shuf_binned_reversed_records"], "shuf_binned",
                               (skeys, perfect_sites(fake_shuffle(genome, 2), skeys, group_by="name")))
    assert n_diff <= 8  # a handful of perfect-match sites in 4 x 240 cells
