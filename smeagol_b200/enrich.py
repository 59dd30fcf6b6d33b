"""Enrichment of PWM matches relative to shuffled backgrounds -- API mirror of smeagol/enrich.py.

Only the two scan_sequences calls are the accelerated path (enrich.py:151-160 and :172-183): the real genome
(sites + counts) and the background set (counts per background id, the expensive one -- simN x more bases) run on
the device with in-kernel per-(sequence, motif, strand) counting.  The statistics (binomial / normal test,
Benjamini-Hochberg FDR per sense) stay on the host as in the reference (enrich.py:18-90), restated with
scipy.stats.binomtest (binom_test was removed from SciPy) and an in-file BH procedure (statsmodels is optional).
"""
import numpy as np
import pandas as pd
import scipy.stats as stats

from .scan import scan_sequences
from .utils import read_bg_seqs, shuffle_records


def _fdrcorrection(pvals):
    """Benjamini-Hochberg adjusted p-values (statsmodels.stats.multitest.fdrcorrection, method 'indep')."""
    p = np.asarray(pvals, dtype=float)
    n = len(p)
    order = np.argsort(p)
    adj = p[order] * n / np.arange(1, n + 1)
    adj = np.minimum(np.minimum.accumulate(adj[::-1])[::-1], 1.0)
    out = np.empty(n)
    out[order] = adj
    return out


def _enrich_over_shuffled(real_counts, shuf_stats, background="binomial", records=None):
    """Test for enrichment of binding sites in real vs. shuffled genomes (enrich.py:18-90), host statistics."""
    on = ["Matrix_id", "width", "sense"] + (["name"] if "name" in shuf_stats.columns else [])
    # outer merge in the key order the reference's goldens (tests/data/enr.csv, pandas < 2) hold: keys of
    # real_counts first, in their order, then the PWMs seen only in the background
    left = real_counts.merge(shuf_stats, on=on, how="left")
    marked = shuf_stats.merge(real_counts[on].drop_duplicates(), on=on, how="left", indicator=True)
    right_only = marked[marked["_merge"] == "left_only"].drop(columns="_merge")
    enr = pd.concat([left, right_only], ignore_index=True)[list(left.columns)]
    enr["num"] = enr["num"].fillna(0)
    num_shuf = enr["len"][0]
    enr.loc[enr.avg == 0, "avg"] = 1 / num_shuf
    if background in ("normal", "both"):
        enr.loc[enr.avg == 0, "sd"] = np.std([1] + [0] * (int(num_shuf) - 1))
        enr["z"] = (enr.num - enr.avg) / enr.sd
        enr["z"] = enr.z.replace([-np.inf], -10)
    if background == "normal":
        enr["pnorm"] = stats.norm.sf(abs(enr.z)) * 2
    else:
        enr["adj_len"] = sum([len(record) - enr.width + 1 for record in records])
        enr["p"] = [stats.binomtest(int(k), int(n), a / n, alternative="two-sided").pvalue
                    for k, n, a in zip(enr["num"], enr["adj_len"], enr["avg"])]
    parts = []
    for sense in pd.unique(enr.sense):
        enr_x = enr[enr.sense == sense].copy()
        if background == "normal":
            enr_x["fdr_norm"] = _fdrcorrection(enr_x.pnorm)
        else:
            enr_x["fdr"] = _fdrcorrection(enr_x.p)
        parts.append(enr_x)
    enr_full = pd.concat(parts)
    # the reference sorts by a non-existent column 'fdr_binom' for the normal background (enrich.py:85)
    enr_full.sort_values(["sense", "fdr_norm" if background == "normal" else "fdr"], inplace=True, kind="stable")
    enr_full.reset_index(inplace=True, drop=True)
    return enr_full


def enrich_in_genome(records, model, rcomp, threshold, simN=1000, simK=2, sense="+", background="binomial",
                     verbose=False, combine_seqs=True, seq_batch=0, shuf_seqs=None):
    """Enrichment of motifs in the given sequence(s) relative to a background (enrich.py:93-203; same arguments
    and result keys: enrichment, real_sites, real_counts, shuf_counts, shuf_stats)."""
    real_preds = scan_sequences(records, model, threshold, sense, rcomp, outputs=["sites", "counts"], score=False,
                                combine_seqs=combine_seqs)
    if shuf_seqs is not None:
        if type(shuf_seqs) == str:
            shuf = read_bg_seqs(shuf_seqs, records, simN)
        elif type(shuf_seqs) == list:
            shuf = shuf_seqs
    else:
        shuf = shuffle_records(records, simN, simK)
    shuf_preds = scan_sequences(shuf, model, threshold, sense, rcomp, outputs=["counts", "stats"], group_by="name",
                                combine_seqs=combine_seqs, sep_ids=True, seq_batch=seq_batch)
    enr = _enrich_over_shuffled(real_preds["counts"], shuf_preds["stats"], background=background, records=records)
    results = {"enrichment": enr, "real_sites": real_preds["sites"], "real_counts": real_preds["counts"],
               "shuf_counts": shuf_preds["counts"], "shuf_stats": shuf_preds["stats"]}
    if verbose:
        results["shuf_seqs"] = shuf
    return results


def examine_thresholds(records, model, simN=300, simK=2, rcomp="only", min_threshold=0.5, threshold_shift=0.1, sense="+",
                       verbose=False, combine_seqs=True):
    """Number of binding sites at different cut-offs of the match score, for the real sequence(s) and for shuffled
    ones (enrich.py:206-284; same arguments and result keys 'real_binned', 'shuf_binned').  Both scans histogram the
    device hit list in place (smg_bin_hits), the shuffles are generated on the device."""
    thresholds = np.arange(min_threshold, 1.0, threshold_shift)
    real_binned = scan_sequences(records, model, thresholds, sense, rcomp, outputs=["binned_counts"],
                                 combine_seqs=combine_seqs)["binned_counts"]
    shuf = shuffle_records(records, simN, simK)
    shuf_binned = scan_sequences(shuf, model, thresholds, sense, rcomp, outputs=["binned_counts"], combine_seqs=combine_seqs,
                                 group_by="name", sep_ids=True)["binned_counts"]
    results = {"real_binned": real_binned, "shuf_binned": shuf_binned}
    if verbose:
        results["shuf_seqs"] = shuf
    return results


def _fisher_window_rows(windows, counts, tot_count, width, genome, alternative):
    """The per-window statistics of enrich_in_window (enrich.py:320-338) for a frame of windows."""
    result = windows.copy()
    effective_genome_len = sum(len(x) for x in genome) - width + 1
    result["len"] = result.end - result.start - width + 1
    result["count"] = counts
    result["tot_count"] = tot_count
    result["expected"] = (result.tot_count * result.len) / effective_genome_len
    tot_neg = effective_genome_len - tot_count
    odds, ps = [], []
    for c, ln in zip(result["count"], result["len"]):
        o, p = stats.fisher_exact([[c, ln - c], [tot_count, tot_neg]], alternative=alternative)
        odds.append(o)
        ps.append(p)
    result["odds"] = odds
    result["p"] = ps
    return result


def _sites_of_matrix(sites, matrix_id):
    """(total sites of the PWM, its width) from a sites DataFrame or a DeviceSites."""
    from .scan import DeviceSites

    if isinstance(sites, DeviceSites):
        import torch

        m = sites.model
        sel = np.flatnonzero(np.asarray(m.Matrix_ids) == matrix_id)
        k = sites.res.keys[:sites.res.n_hits]
        motif = k & ((1 << sites.res.motif_bits) - 1)
        tot = int(torch.isin(motif, torch.from_numpy(sel).to(k.device)).sum().item()) if sites.res.n_hits else 0
        return tot, int(m.widths[sel[0]])
    matrix_sites = sites[sites.Matrix_id == matrix_id].reset_index(drop=True)
    return len(matrix_sites), int(matrix_sites.width[0])


def enrich_in_window(window, sites, genome, matrix_id=None, alternative="two-sided"):
    """Fisher's exact test for enrichment of a PWM's sites in one window relative to the whole genome (enrich.py:287-338;
    `window`: a row / one-row frame with id, start, end)."""
    from .scan import _count_in_window

    tot, width = _sites_of_matrix(sites, matrix_id)
    w = window.to_frame().T if isinstance(window, pd.Series) else window
    counts = [_count_in_window(row, sites, matrix_id) for _, row in w.iterrows()]
    res = _fisher_window_rows(w, counts, tot, width, genome, alternative)
    return res.iloc[0] if isinstance(window, pd.Series) else res


def enrich_in_sliding_windows(sites, genome, width, shift=None, matrix_id=None, alternative="two-sided"):
    """Enrichment of a PWM's sites in sliding windows across a genome (enrich.py:341-374): tiling windows, window counts
    (on the device when `sites` is a scan.DeviceSites), Fisher's exact test per window, overall BH-FDR (`padj`)."""
    from .scan import count_in_sliding_windows

    counted = count_in_sliding_windows(sites, genome, matrix_id, width, shift)
    tot, mwidth = _sites_of_matrix(sites, matrix_id)
    windows = _fisher_window_rows(counted[["id", "start", "end"]], counted["count"].to_numpy(), tot, mwidth, genome,
                                  alternative).reset_index(drop=True)
    windows["padj"] = _fdrcorrection(windows.p)
    return windows
