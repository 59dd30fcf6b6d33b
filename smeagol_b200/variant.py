"""Variant effect prediction -- API mirror of smeagol/variant.py plus the batched form of BASELINE config 5.

variant_effect_on_sites keeps the reference's signature and output columns (variant.py:50-103).  The reference masks
every variant against every site with pandas (O(#sites x #variants)); here overlapping (site, variant) pairs are
found with a sorted search on the host and re-scored in fp64 on the device (smg_variant_sites).
variant_window_scores is the batched restatement: for every SNV and PWM, maximum reference / alternate score over all
windows covering the SNV on both strands (smg_variant_windows).
"""
import ctypes

import numpy as np
import pandas as pd
import torch

from . import _lib
from . import device as dev
from .encode import integer_encoding_dict
from .utils import read_fasta, seq_str


def _model_for(pwms):
    from .models import PWMModel

    return pwms if isinstance(pwms, PWMModel) else PWMModel(pwms)


def variant_effect_on_sites(sites, variants, seqs, pwms):
    """Predict variant effects on sites (variant.py:50-103).

    sites: DataFrame with columns name, start, end, Matrix_id (optionally score); variants: DataFrame with name, pos,
    alt; seqs: list of SeqRecord-like objects or FASTA path; pwms: DataFrame with Matrix_id and weights.
    Returns one row per overlapping (variant, site) pair: the variant columns, the site columns, then seq, max_score,
    score and variant_score (fractions of max_score), in the reference's site-major order.
    """
    if type(seqs) == str:
        seqs = read_fasta(seqs)
    model = _model_for(pwms)
    names = [r.name for r in seqs]
    strings = [seq_str(r) for r in seqs]
    first_row = {}
    for i, n in enumerate(names):
        first_row.setdefault(n, i)  # the reference takes the first record with a matching name (utils.py:194)
    mid_to_idx = {}
    for i, mid in enumerate(model.Matrix_ids):
        mid_to_idx.setdefault(mid, i)  # pwms.weights[pwms.Matrix_id == x].values[0] (variant.py:78-81)
    sites = sites.reset_index(drop=True)
    variants = variants.reset_index(drop=True)
    s_row = np.array([first_row[n] for n in sites["name"]], dtype=np.int32)
    s_start = sites["start"].to_numpy().astype(np.int64)
    s_end = sites["end"].to_numpy().astype(np.int64)
    s_motif = np.array([mid_to_idx[m] for m in sites["Matrix_id"]], dtype=np.int32)
    # overlapping pairs: per sequence name the variants are sorted by position once and every site's [start, end) is
    # located with two binary searches (O((sites + variants) log variants) instead of the reference's O(sites x variants)
    # boolean masks); pairs come out site-major and, within a site, in the variants' original order (variant.py:85-92)
    pair_site, pair_var = [], []
    s_names = sites["name"].to_numpy()
    v_pos_all = variants["pos"].to_numpy()
    for name, vidx in variants.groupby("name", sort=False).indices.items():
        sidx = np.flatnonzero(s_names == name)
        if sidx.size == 0:
            continue
        vidx = np.asarray(vidx, dtype=np.int64)
        order = np.argsort(v_pos_all[vidx], kind="stable")
        vps = v_pos_all[vidx][order]
        lo = np.searchsorted(vps, s_start[sidx], side="left")
        hi = np.searchsorted(vps, s_end[sidx], side="left")
        cnt = np.maximum(hi - lo, 0)
        tot = int(cnt.sum())
        if tot == 0:
            continue
        ps = np.repeat(sidx, cnt)
        within = np.arange(tot) - np.repeat(np.cumsum(cnt) - cnt, cnt)
        pv = vidx[order[np.repeat(lo, cnt) + within]]
        pair_site.append(ps)
        pair_var.append(pv)
    if pair_site:
        pair_site, pair_var = np.concatenate(pair_site), np.concatenate(pair_var)
        o = np.lexsort((pair_var, pair_site))
        pair_site, pair_var = [pair_site[o]], [pair_var[o]]
    site_seq = [strings[r][a:b] for r, a, b in zip(s_row, s_start, s_end)]
    max_score = np.array([np.sum(np.max(model.weights[m], axis=1)) for m in s_motif])
    if not pair_site:
        return pd.DataFrame()
    pair_site = np.concatenate(pair_site)
    pair_var = np.concatenate(pair_var)
    rel = (variants["pos"].to_numpy()[pair_var] - s_start[pair_site]).astype(np.int32)
    alt = np.array([integer_encoding_dict[a] for a in variants["alt"].to_numpy()[pair_var]], dtype=np.uint8)
    # device: fp64 reference and variant scores of every pair.  The reference uses the unrounded fp64 weights on the
    # host (scan.py:8-37); the device tables are their fp32 casts -> relative difference <= 6e-8, inside the 1e-5
    # tolerance of the path (stated in tests/test_gpu_parity.py::test_variant_golden_and_random).
    batch = dev.SeqBatch(strings, kind="ascii")
    dset = model.device_set
    d = batch.device
    t = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(d)  # noqa: E731
    out_ref = torch.empty(len(pair_site), dtype=torch.float64, device=d)
    out_alt = torch.empty(len(pair_site), dtype=torch.float64, device=d)
    a_row, a_start, a_motif, a_rel, a_alt = t(s_row[pair_site]), t(s_start[pair_site].astype(np.int32)), t(s_motif[pair_site]), t(rel), t(alt)
    _lib.check(_lib.load().smg_variant_sites(dset.handle, dev._ptr(batch.packed), dev._ptr(batch.amask), dev._ptr(batch.codes),
                                             dev._ptr(batch.d_rows), dev._ptr(a_row), dev._ptr(a_start), dev._ptr(a_motif),
                                             dev._ptr(a_rel), dev._ptr(a_alt), len(pair_site), dev._ptr(out_ref),
                                             dev._ptr(out_alt), dev._stream()), "smg_variant_sites")
    ref = out_ref.cpu().numpy()
    altv = out_alt.cpu().numpy()
    res = variants.iloc[pair_var].reset_index(drop=True)
    site_cols = sites.iloc[pair_site].reset_index(drop=True)
    for c in site_cols.columns:
        if c not in res.columns:
            res[c] = site_cols[c]
    res["seq"] = [site_seq[i] for i in pair_site]
    res["max_score"] = max_score[pair_site]
    if "score" not in sites.columns:
        res["score"] = ref / res["max_score"]
    res["variant_score"] = altv / res["max_score"]
    return res


def variant_window_scores(seqs, snv_row, snv_pos, snv_alt, pwms, to_host=True, exact=False):
    """Batched SNV scoring (BASELINE config 5): for every SNV (row index into seqs, 0-based position, alternate base)
    and every PWM, the maximum reference and alternate score over all windows covering the SNV on both strands.
    Returns (ref_max, alt_max) float32 arrays of shape (n_snv, n_motifs); delta = alt_max - ref_max.

    exact=False (default): alternate window scores are formed as (reference score - T[k, ref]) + T[k, alt], which is
    what makes the batch cheap; they agree with a fresh fp32 sum to a few ulp of the largest partial sum.  exact=True
    forks the alternate sums from the reference sums instead (bit-identical to the fresh ascending-k sum, ~1.3x slower).
    Reference maxima are bit-identical either way."""
    model = _model_for(pwms)
    strings = [seq_str(r) if not isinstance(r, (bytes, np.ndarray)) else r for r in seqs]
    batch = dev.SeqBatch(strings, kind="ascii")
    return variant_window_scores_on_batch(batch, snv_row, snv_pos, snv_alt, model, to_host, exact)


def variant_window_scores_on_batch(batch, snv_row, snv_pos, snv_alt, model, to_host=True, exact=False, motif_major=True):
    dset = model.device_set
    d = batch.device
    n = len(snv_pos)
    if isinstance(snv_alt, torch.Tensor) or (isinstance(snv_alt, np.ndarray) and snv_alt.dtype == np.uint8):
        alt = snv_alt
    else:
        alt = np.array([integer_encoding_dict[a] for a in snv_alt], dtype=np.uint8)
    def t(a, dt):  # device tensors are taken as they are (a resident batch of SNVs scored repeatedly)
        if isinstance(a, torch.Tensor):
            return a.to(device=d, dtype={np.int32: torch.int32, np.uint8: torch.uint8}[dt]).contiguous()
        return torch.from_numpy(np.ascontiguousarray(a, dtype=dt)).to(d)
    # motif-major storage (the layout the kernels can write as whole lines), handed back as (n_snv, n_motifs) views
    shape = (dset.n_motifs, n) if motif_major else (n, dset.n_motifs)
    ref_max = torch.empty(shape, dtype=torch.float32, device=d)
    alt_max = torch.empty(shape, dtype=torch.float32, device=d)
    a_row, a_pos, a_alt = t(snv_row, np.int32), t(snv_pos, np.int32), t(alt, np.uint8)
    _lib.check(_lib.load().smg_variant_windows_mode(dset.handle, dev._ptr(batch.packed), dev._ptr(batch.amask),
                                                    dev._ptr(batch.codes), dev._ptr(batch.d_rows), dev._ptr(a_row),
                                                    dev._ptr(a_pos), dev._ptr(a_alt), n, dev._ptr(ref_max), dev._ptr(alt_max),
                                                    1 if exact else 0, 1 if motif_major else 0, dev._stream()),
               "smg_variant_windows_mode")
    if to_host:
        ref_max, alt_max = ref_max.cpu().numpy(), alt_max.cpu().numpy()
    return (ref_max.T, alt_max.T) if motif_major else (ref_max, alt_max)
