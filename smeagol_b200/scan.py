"""Scan orchestration -- API mirror of smeagol/scan.py.

The reference loops over SeqEncoding groups, runs one Keras predict per group and builds DataFrames with pandas
fancy indexing / crosstab (scan.py:150-312).  Here ALL groups go through one fused device scan (both strands in the
same kernel, hits sorted on the device into the reference's (group, row, pos, motif) order, per-(row, motif,
strand) counts accumulated in the kernel); the DataFrames are then assembled with vectorised numpy from the hit
list / count table, keeping the reference's schemas, ordering and sparse-crosstab semantics.
"""
import numpy as np
import pandas as pd

from . import device as dev
from .encode import SeqGroups, one_hot_dict, one_hot_encode
from .encode import sense_complement_dict


def score_site(pwm, seq, position_wise=False):
    """PWM match score of a site sequence as long as the PWM (scan.py:8-24), host fp64."""
    seq = one_hot_encode(seq, rc=False)
    assert seq.shape == pwm.shape
    position_wise_scores = np.sum(np.multiply(seq, pwm), axis=1)
    return position_wise_scores if position_wise else np.sum(position_wise_scores)


def _score_base(pwm, base, pos):
    """Match score of one base at one PWM position (scan.py:27-37)."""
    return np.sum(np.multiply(pwm[pos, :], one_hot_dict[base]))


def _take(labels, idx):
    """labels[idx] as a DataFrame column.  pandas >= 3 converts object arrays of str element by element; converting the
    (short) label array once and gathering with take() keeps large site lists cheap."""
    labels = np.asarray(labels)
    idx = np.asarray(idx, dtype=np.int64)
    if labels.dtype.kind not in "OUS":
        return labels[idx]
    try:
        return pd.array(labels).take(idx)
    except Exception:
        return labels.astype(object)[idx]


def _sites_frame(ids, names, senses, row_idx, pos_idx, pwm_idx, model, scores=None):
    """One DataFrame construction with the reference's column order (scan.py:59-73)."""
    pos_idx = np.asarray(pos_idx)
    width = model.widths[pwm_idx]
    cols = {"id": _take(ids, row_idx), "name": _take(names, row_idx), "sense": _take(senses, row_idx), "start": pos_idx,
            "Matrix_id": _take(model.Matrix_ids, pwm_idx), "width": width, "end": pos_idx + width}
    if scores is not None:
        max_score = model.max_scores[pwm_idx]
        cols["score"] = scores
        cols["max_score"] = max_score
        cols["frac_score"] = scores / max_score
    return pd.DataFrame(cols, copy=False)


def _locate_sites(encoding, model, thresholded, scores=None):
    """DataFrame of binding sites from a thresholded triple (scan.py:40-74); same columns and order."""
    seq_idx, pos_idx, pwm_idx = (np.asarray(t) for t in thresholded)
    return _sites_frame(encoding.ids, encoding.names, encoding.senses, seq_idx, pos_idx, pwm_idx, model, scores)


_EMPTY_COUNTS = ["Matrix_id", "width", "id", "name", "sense", "num"]


def _melt_counts_parts(table, row_labels, model):
    """Index form of one group's melted crosstab: (m_idx, col_idx, labels, num) with labels = sorted unique
    (id, name, sense) tuples; row j of the frame is PWM m_idx[j] x label col_idx[j] with count num[j].  None if empty."""
    ids, names, senses = row_labels
    table = _collapse_motifs(table, model, 1)
    # rows of the encoding sharing (id, name, sense) collapse into one crosstab column; columns are sorted tuples
    keys = list(zip(ids, names, senses))
    col_uniques = sorted(set(keys))
    n_cols = len(col_uniques)
    if n_cols == len(keys):
        if n_cols == 1:
            agg = table.astype(np.int64, copy=False)
        else:
            pos_of = {k: i for i, k in enumerate(col_uniques)}
            col_codes = np.fromiter((pos_of[k] for k in keys), dtype=np.int64, count=len(keys))
            agg = np.empty((n_cols, table.shape[1]), dtype=np.int64)
            agg[col_codes] = table
    else:
        pos_of = {k: i for i, k in enumerate(col_uniques)}
        col_codes = np.fromiter((pos_of[k] for k in keys), dtype=np.int64, count=len(keys))
        agg = np.zeros((n_cols, table.shape[1]), dtype=np.int64)
        np.add.at(agg, col_codes, table)
    keep_cols = np.flatnonzero(agg.sum(axis=1) > 0)
    keep_m = np.flatnonzero(agg.sum(axis=0) > 0)
    if keep_cols.size == 0:
        return None
    m_order = keep_m[np.argsort(_motif_rank(model)[keep_m], kind="stable")]
    sub = agg[np.ix_(keep_cols, m_order)]
    return np.tile(m_order, len(keep_cols)), np.repeat(keep_cols, len(m_order)), col_uniques, sub.reshape(-1)


def _counts_frame(parts, model):
    """ONE DataFrame for the melted crosstabs of many groups (in group order): a frame per group costs ~1 ms of pandas
    overhead, which for 1,000 single-sequence groups was the whole cost of a `counts` call."""
    parts = [p for p in parts if p is not None]
    if not parts:
        return pd.DataFrame({k: [] for k in _EMPTY_COUNTS})
    labels, offs, base = [], [], 0
    for _, _, lab, _ in parts:
        offs.append(base)
        labels.extend(lab)
        base += len(lab)
    m_idx = np.concatenate([p[0] for p in parts])
    c_idx = np.concatenate([p[1] + o for p, o in zip(parts, offs)])
    return pd.DataFrame({
        "Matrix_id": _take(model.Matrix_ids, m_idx),
        "width": model.widths[m_idx],
        "id": _take([c[0] for c in labels], c_idx),
        "name": _take([c[1] for c in labels], c_idx),
        "sense": _take([c[2] for c in labels], c_idx),
        "num": np.concatenate([p[3] for p in parts]),
    })


def _melt_counts(table, row_labels, model, extra=None):
    """Sparse-crosstab semantics of scan._count_sites (scan.py:137-146): `table` is (n_rows, n_motifs) of counts for
    one group; rows kept = PWMs with >=1 hit, columns kept = (id, name, sense) labels with >=1 hit; output iterates
    sorted columns (outer) x sorted (Matrix_id, width) (inner)."""
    return _counts_frame([_melt_counts_parts(table, row_labels, model)], model)


def _crosstab_rows(model):
    """The crosstab's row index is (Matrix_id, width): PWMs sharing both collapse into ONE summed row (scan.py:137-140).
    Returns (rep, inverse): rep[j] = first PWM of unique key j (keys in order of first appearance), inverse[m] = key of
    PWM m; (None, None) when every key is unique (the usual case)."""
    cached = getattr(model, "_crosstab_rows", None)
    if cached is None:
        seen, rep, inverse = {}, [], np.empty(model.channels, dtype=np.int64)
        for m in range(model.channels):
            k = (model.Matrix_ids[m], int(model.widths[m]))
            if k not in seen:
                seen[k] = len(rep)
                rep.append(m)
            inverse[m] = seen[k]
        cached = (None, None) if len(rep) == model.channels else (np.array(rep, dtype=np.int64), inverse)
        model._crosstab_rows = cached
    return cached


def _collapse_motifs(table, model, axis):
    """Sum the PWM axis of a count table over PWMs with equal (Matrix_id, width); entries of the collapsed axis sit at
    the index of the key's first PWM (the other duplicates become all-zero and drop out as unobserved rows)."""
    rep, inverse = _crosstab_rows(model)
    if rep is None:
        return table
    out = np.zeros_like(table)
    idx = [slice(None)] * table.ndim
    idx[axis] = rep[inverse]
    np.add.at(out, tuple(idx), table)
    return out


def _motif_rank(model):
    """Rank of every PWM in the crosstab's row order: sorted by (Matrix_id, width) (scan.py:137-140)."""
    rank = getattr(model, "_crosstab_rank", None)
    if rank is None:
        order = sorted(range(model.channels), key=lambda i: (model.Matrix_ids[i], model.widths[i]))
        rank = np.empty(model.channels, dtype=np.int64)
        rank[order] = np.arange(model.channels)
        model._crosstab_rank = rank
    return rank


def _count_sites(encoding, model, thresholded):
    """Number of matches per PWM and sequence from a thresholded triple (scan.py:114-147)."""
    seq_idx = np.asarray(thresholded[0])
    if len(seq_idx) == 0:
        return pd.DataFrame({k: [] for k in _EMPTY_COUNTS})
    pwm_idx = np.asarray(thresholded[2])
    table = np.zeros((len(encoding.ids), model.channels), dtype=np.int64)
    np.add.at(table, (seq_idx, pwm_idx), 1)
    return _melt_counts(table, (encoding.ids, encoding.names, encoding.senses), model)


def _bin_sites_by_score(encoding, model, thresholded, scores, bins):
    """Binned site counts (scan.py:77-111); host pandas on the device-produced hit list."""
    seq_idx = np.asarray(thresholded[0])
    pwm_idx = np.asarray(thresholded[2])
    frac_scores = scores / model.max_scores[pwm_idx]
    binned_scores = pd.cut(frac_scores, np.concatenate([bins, [1]]))
    result = pd.crosstab(
        index=[model.Matrix_ids[pwm_idx], model.widths[pwm_idx]],
        columns=[encoding.ids[seq_idx], encoding.names[seq_idx], encoding.senses[seq_idx], binned_scores],
    )
    result = result.melt(ignore_index=False).reset_index()
    result.columns = ["Matrix_id", "width", "id", "name", "sense", "bin", "num"]
    return result


def _device_scan_groups(encoded, model, frac_threshold, want_hits, want_counts, sort=True):
    """One fused device scan over every SeqEncoding group.  Returns (ScanResult, layout) where layout maps the
    device rows / output rows back to groups."""
    groups = encoded.seqs
    rcomp = groups[0].rcomp
    fwd = []          # forward strings, one device row each
    out_rows = []     # (n_device_rows, 2): output row of (row, strand)
    group_first_out = []
    base_out = 0
    for g in groups:
        n = g.n_records
        group_first_out.append(base_out)
        for i in range(n):
            fwd.append(g.records[i])
            if rcomp == "none":
                out_rows.append((base_out + i, 0))
            elif rcomp == "only":
                out_rows.append((0, base_out + i))
            else:
                out_rows.append((base_out + i, base_out + n + i))
        base_out += len(g.ids)
    out_rows = np.array(out_rows, dtype=np.int32)
    from .seqrecord import DeviceSeq

    if fwd and all(isinstance(f, DeviceSeq) for f in fwd) and all(f.flat is fwd[0].flat for f in fwd):
        # device-resident records (GPU shuffles, device FASTA ingest): pack straight from their buffer
        batch = dev.SeqBatch(None, flat=(fwd[0].flat, np.array([f.offset for f in fwd], dtype=np.int64),
                                         np.array([f.length for f in fwd], dtype=np.int64)), device=fwd[0].flat.device)
    else:
        batch = dev.SeqBatch([str(f) if isinstance(f, DeviceSeq) else f for f in fwd], kind="ascii")
    res = dev.scan(model.device_set, batch, model.thresholds(frac_threshold), dev.STRANDS[rcomp], out_rows, base_out,
                   want_hits=want_hits, want_counts=want_counts, adjudicate=model.adjudicate, sort=sort)
    model.last_scan_stats = res.stats
    layout = {"group_first_out": np.array(group_first_out + [base_out], dtype=np.int64), "out_rows": out_rows,
              "n_out": base_out}
    return res, layout


def _melt_binned(hist, row_labels, model, edges):
    """Frame of scan._bin_sites_by_score (scan.py:77-111) from a dense histogram `hist` (n_rows, n_motifs, n_bins) of one
    group: the crosstab keeps the OBSERVED (Matrix_id, width) rows and the OBSERVED (id, name, sense, bin) columns,
    zero-fills inside that rectangle, and melt walks sorted columns (outer) x sorted rows (inner); `bin` is the
    ordered categorical of pd.cut's right-closed intervals."""
    ids, names, senses = row_labels
    hist = _collapse_motifs(hist, model, 1)
    keys = list(zip(ids, names, senses))
    col_uniques = sorted(set(keys))
    pos_of = {k: i for i, k in enumerate(col_uniques)}
    col_codes = np.fromiter((pos_of[k] for k in keys), dtype=np.int64, count=len(keys))
    agg = np.zeros((len(col_uniques),) + hist.shape[1:], dtype=np.int64)
    np.add.at(agg, col_codes, hist)
    # the labels must be pd.cut's own (it rounds the breaks to 3 significant digits: (0.7, 0.8], not
    # (0.7, 0.7999999999999999]) so that `bin` equals the reference's and the host route's column
    cats = pd.cut(np.zeros(0), np.asarray(edges, dtype=np.float64)).categories
    cols = np.argwhere(agg.sum(axis=1) > 0)          # observed (label, bin) pairs, already in sorted order
    keep_m = np.flatnonzero(agg.sum(axis=(0, 2)) > 0)
    if cols.shape[0] == 0:
        out = pd.DataFrame({k: [] for k in ["Matrix_id", "width", "id", "name", "sense"]})
        out["bin"] = pd.Categorical.from_codes([], categories=cats, ordered=True)
        out["num"] = np.zeros(0, dtype=np.int64)
        return out
    m_order = keep_m[np.argsort(_motif_rank(model)[keep_m], kind="stable")]
    nm = len(m_order)
    lab = np.repeat(cols[:, 0], nm)
    out = pd.DataFrame({
        "Matrix_id": _take(model.Matrix_ids, np.tile(m_order, cols.shape[0])),
        "width": np.tile(model.widths[m_order], cols.shape[0]),
        "id": _take([c[0] for c in col_uniques], lab),
        "name": _take([c[1] for c in col_uniques], lab),
        "sense": _take([c[2] for c in col_uniques], lab),
    })
    out["bin"] = pd.Categorical.from_codes(np.repeat(cols[:, 1], nm), categories=cats, ordered=True)
    out["num"] = agg[cols[:, 0]][:, m_order, :][np.arange(cols.shape[0]), :, cols[:, 1]].reshape(-1)
    return out


def _binned_counts_device(res, layout, groups, model, bins):
    """Binned counts of every group from the device hit list (smg_bin_hits): no per-hit data leaves the GPU."""
    import torch

    from . import _lib

    edges = np.concatenate([np.asarray(bins, dtype=np.float64), [1.0]])
    n_bins = len(edges) - 1
    d = res.counts.device if res.counts is not None else res.keys.device
    hist = torch.zeros((layout["n_out"], model.channels, n_bins), dtype=torch.int32, device=d)
    if res.n_hits:
        d_max = torch.from_numpy(np.asarray(model.max_scores, dtype=np.float64)).to(d)
        d_edges = torch.from_numpy(edges).to(d)
        with torch.cuda.device(d):
            _lib.check(_lib.load().smg_bin_hits(dev._ptr(res.keys), dev._ptr(res.scores), res.n_hits, res.pos_bits, res.motif_bits,
                                                dev._ptr(d_max), dev._ptr(d_edges), n_bins, model.channels, dev._ptr(hist),
                                                dev._stream()), "smg_bin_hits")
    h = hist.cpu().numpy().astype(np.int64)
    gfo = layout["group_first_out"]
    frames = [_melt_binned(h[gfo[gi]:gfo[gi + 1]], (g.ids, g.names, g.senses), model, edges) for gi, g in enumerate(groups)]
    return pd.concat(frames).reset_index(drop=True)


def _find_sites_in_groups(encoding, model, threshold, outputs=["sites"], score=False, combine_seqs=False, sep_ids=False,
                          seq_batch=0):
    """Predict binding sites on a SeqGroups object (scan.py:211-312): same outputs, one device scan."""
    outputs = list(outputs)
    binned = "binned_counts" in outputs
    if binned:
        assert type(threshold) == np.ndarray
        score = True
        frac = min(threshold)
    else:
        frac = threshold
    groups = encoding.seqs
    want_sites = ("sites" in outputs) or binned
    want_counts = ("counts" in outputs) or ("stats" in outputs)
    device_bins = binned and "sites" not in outputs  # histogram on the device: the hit list never leaves the GPU
    res, layout = _device_scan_groups(encoding, model, frac, want_hits=want_sites, want_counts=want_counts,
                                      sort=not device_bins)
    ids_all = np.concatenate([g.ids for g in groups]).astype(object)
    names_all = np.concatenate([g.names for g in groups]).astype(object)
    senses_all = np.concatenate([g.senses for g in groups]).astype(object)
    output = {}
    if device_bins:
        output["binned_counts"] = _binned_counts_device(res, layout, groups, model, threshold)
    elif want_sites:
        row, pos, motif, sc = res.hits_host()
        if binned:
            gfo = layout["group_first_out"]
            frames = []
            for gi, g in enumerate(groups):
                sel = (row >= gfo[gi]) & (row < gfo[gi + 1])
                frames.append(_bin_sites_by_score(g, model, (row[sel] - gfo[gi], pos[sel], motif[sel]), sc[sel], threshold))
            output["binned_counts"] = pd.concat(frames).reset_index(drop=True)
        if "sites" in outputs:
            output["sites"] = _sites_frame(ids_all, names_all, senses_all, row, pos, motif, model, sc if score else None)
    if want_counts:
        counts_dev = res.counts.cpu().numpy().astype(np.int64)  # (n_device_rows, M, 2)
        table = np.zeros((layout["n_out"], model.channels), dtype=np.int64)
        orow = layout["out_rows"]
        strands = dev.STRANDS[groups[0].rcomp]
        if strands & 1:
            table[orow[:, 0]] = counts_dev[:, :, 0]
        if strands & 2:
            table[orow[:, 1]] = counts_dev[:, :, 1]
        gfo = layout["group_first_out"]
        parts = [_melt_counts_parts(table[gfo[gi]:gfo[gi + 1]], (g.ids, g.names, g.senses), model) for gi, g in enumerate(groups)]
        counts = _counts_frame(parts, model)
        frames = [_counts_frame([pt], model) for pt in parts] if ("stats" in outputs and not combine_seqs) else None
        if "counts" in outputs or combine_seqs:
            output["counts"] = counts
        if "stats" in outputs and not combine_seqs:
            # per-group stats of the reference (scan.py:198-206): groupby over Matrix_id, width, sense, name
            st = []
            for f in frames:
                g = f.groupby(["Matrix_id", "width", "sense", "name"]).num
                st.append(pd.DataFrame({"len": g.size(), "avg": g.mean(), "sd": g.std(ddof=1)}).reset_index())
            output["stats"] = pd.concat(st).reset_index(drop=True)
    if combine_seqs:
        keys = ["Matrix_id", "width", "sense"] + (["id"] if sep_ids else [])
        if "counts" in output:
            output["counts"] = output["counts"].groupby(keys).num.sum().reset_index()
        if "binned_counts" in output:
            bkeys = ["Matrix_id", "width", "sense", "bin"] + (["id"] if sep_ids else [])
            output["binned_counts"] = output["binned_counts"].groupby(bkeys, observed=False).num.sum().reset_index()
        if "stats" in outputs:
            # len / mean / std with ddof=1 -- the value the reference's golden tests/data/shuf_stats.csv holds
            # (pandas >= 2 would give ddof=0 for the reference's .agg([len, np.mean, np.std]), scan.py:304)
            g = output["counts"].groupby(["Matrix_id", "width", "sense"]).num
            output["stats"] = pd.DataFrame({"len": g.size(), "avg": g.mean(), "sd": g.std(ddof=1)}).reset_index()
    if "counts" not in outputs and "counts" in output:
        del output["counts"]
    return output


def _find_sites_seq(encoding, model, threshold, outputs=["sites"], score=False, seq_batch=0):
    """Predict binding sites on one SeqEncoding (scan.py:150-208)."""
    class _One:
        seqs = [encoding]
    return _find_sites_in_groups(_One, model, threshold, outputs=outputs, score=score, seq_batch=seq_batch)


def scan_sequences(seqs, model, threshold, sense="+", rcomp="none", outputs=["sites"], score=False, group_by=None,
                   combine_seqs=False, sep_ids=False, seq_batch=0):
    """Encode given sequences and predict binding sites on them (scan.py:315-375; same signature and outputs).

    seqs: list of SeqRecord-like objects or a FASTA path; model: PWMModel; threshold: fraction of the maximum score
    (or np.ndarray of bin edges with outputs=['binned_counts']); sense: '+'/'-'; rcomp: 'none' | 'only' | 'both';
    outputs: any of 'sites', 'counts', 'binned_counts', 'stats'; group_by: record attribute to group by;
    combine_seqs / sep_ids: combine outputs over groups (optionally keeping ids apart); seq_batch: accepted for
    compatibility (the device scan needs no batching).
    """
    encoded = SeqGroups(seqs, rcomp=rcomp, sense=sense, group_by=group_by)
    return _find_sites_in_groups(encoded, model, threshold=threshold, outputs=outputs, score=score,
                                 combine_seqs=combine_seqs, sep_ids=sep_ids, seq_batch=seq_batch)


# ---------------------------------------------------------------------- analysis in windows (scan.py:381-423)
class DeviceSites:
    """The sites of one scan kept on the device (sorted hit keys + the row layout), for consumers that never need the
    DataFrame: count_in_sliding_windows / enrich_in_sliding_windows histogram them in place (smg_window_counts)."""

    def __init__(self, res, layout, groups, model):
        self.res, self.layout, self.groups, self.model = res, layout, groups, model
        self.row_ids = np.concatenate([g.ids for g in groups]).astype(object)  # id of every output row

    def __len__(self):
        return int(self.res.n_hits)


def scan_sequences_device(seqs, model, threshold, sense="+", rcomp="none", group_by=None):
    """scan_sequences(..., outputs=['sites']) that leaves the site list on the device (DeviceSites)."""
    encoded = SeqGroups(seqs, rcomp=rcomp, sense=sense, group_by=group_by)
    res, layout = _device_scan_groups(encoded, model, threshold, want_hits=True, want_counts=False, sort=False)
    return DeviceSites(res, layout, encoded.seqs, model)


def _count_in_window(window, sites, matrix_id=None):
    """Count of sites of one PWM whose start lies in a window (scan.py:381-401)."""
    sel = (sites.id == window.id) & (sites.start >= window.start) & (sites.start < window.end)
    if matrix_id is not None:
        sel &= sites.Matrix_id == matrix_id
    return int(sel.sum())


def _window_counts_device(dsites, windows, genome, matrix_id, width, shift):
    import torch

    from . import _lib

    res, model = dsites.res, dsites.model
    d = res.keys.device if res.keys is not None else dev._require_cuda()
    rec_of = {r.id: i for i, r in enumerate(genome)}
    n_win = np.array([len(np.arange(0, len(r), shift)) for r in genome], dtype=np.int64)
    win_off = np.concatenate([[0], np.cumsum(n_win)]).astype(np.int64)
    row_rec = np.array([rec_of.get(i, -1) for i in dsites.row_ids], dtype=np.int32)
    sel = np.ones(model.channels, dtype=np.uint8) if matrix_id is None else (np.asarray(model.Matrix_ids) == matrix_id).astype(np.uint8)
    counts = torch.zeros(max(int(win_off[-1]), 1), dtype=torch.int32, device=d)
    if res.n_hits:
        # named tensors: a temporary would be freed (and its block reused by the next upload) before the call is made
        d_sel, d_rec, d_off = torch.from_numpy(sel).to(d), torch.from_numpy(row_rec).to(d), torch.from_numpy(win_off).to(d)
        with torch.cuda.device(d):
            _lib.check(_lib.load().smg_window_counts(dev._ptr(res.keys), res.n_hits, res.pos_bits, res.motif_bits, dev._ptr(d_sel),
                                                     dev._ptr(d_rec), dev._ptr(d_off), int(width), int(shift), dev._ptr(counts),
                                                     dev._stream()), "smg_window_counts")
    return counts[:int(win_off[-1])].cpu().numpy().astype(np.int64)


def count_in_sliding_windows(sites, genome, matrix_id, width, shift=None):
    """Count binding sites in sliding windows across a genome (scan.py:404-423; same arguments and frame: id, start, end,
    count).  `sites` is the reference's sites DataFrame (counted on the host with two binary searches per record instead
    of one boolean mask per window) or a DeviceSites from scan_sequences_device (histogram on the device)."""
    from .utils import get_tiling_windows_over_genome

    shift_ = width if shift is None else shift
    windows = get_tiling_windows_over_genome(genome, width, shift).reset_index(drop=True)
    if isinstance(sites, DeviceSites):
        windows["count"] = _window_counts_device(sites, windows, genome, matrix_id, width, shift_)
        return windows
    sel = sites if matrix_id is None else sites[sites.Matrix_id == matrix_id]
    counts = np.zeros(len(windows), dtype=np.int64)
    for rid, idx in windows.groupby("id", sort=False).groups.items():
        st = np.sort(sel.start[sel.id == rid].to_numpy())
        w = windows.loc[idx]
        counts[np.asarray(idx)] = np.searchsorted(st, w.end.to_numpy(), "left") - np.searchsorted(st, w.start.to_numpy(), "left")
    windows["count"] = counts
    return windows
