"""TEST INFRASTRUCTURE ONLY -- CPU restatement of the DataFrame layer of smeagol/scan.py on top of
oracle/scan_oracle.py.  It restates, group by group exactly as the reference iterates them:

  * grouping / row order        encode.py:185-218 (SeqGroups) and encode.py:137-160 (SeqEncoding)
  * sites schema                scan.py:40-74   (_locate_sites)
  * counts (sparse crosstab)    scan.py:114-147 (_count_sites)
  * binned counts               scan.py:77-111  (_bin_sites_by_score)
  * concat / combine / stats    scan.py:211-312 (_find_sites_in_groups); std pinned to ddof=1, which is what the
                                reference's golden tests/data/shuf_stats.csv holds (pandas < 2 behaviour of
                                ``.agg([len, np.mean, np.std])``)

Never imported by the product package.
"""
from collections import OrderedDict

import numpy as np
import pandas as pd

from . import scan_oracle as so


def make_groups(ids, names, group_by=None):
    """Return a list of lists of record indices, as SeqGroups forms them."""
    n = len(ids)
    if group_by is not None and n > 1:
        key = {"name": names, "id": ids}[group_by]
        d = OrderedDict()
        for i in range(n):
            d.setdefault(key[i], []).append(i)
        return list(d.values())
    return [[i] for i in range(n)]


def group_rows(idx, ids, names, sense, rcomp):
    """ids / names / senses arrays of one SeqEncoding (encode.py:137-160)."""
    gid = np.array([ids[i] for i in idx], dtype=object)
    gnm = np.array([names[i] for i in idx], dtype=object)
    comp = so.SENSE_COMPLEMENT[sense]
    if rcomp == "none":
        return gid, gnm, np.array([sense] * len(idx), dtype=object)
    if rcomp == "only":
        return gid, gnm, np.array([comp] * len(idx), dtype=object)
    return (np.tile(gid, 2), np.tile(gnm, 2),
            np.array([sense] * len(idx) + [comp] * len(idx), dtype=object))


def scan_sequences(seqs, ids, names, matrix_ids, weights, threshold, sense="+", rcomp="none", outputs=("sites",),
                   score=False, group_by=None, combine_seqs=False, sep_ids=False, seq_batch=0, order="ref",
                   adjudicate=True):
    # seq_batch only changes how the reference batches rows through Keras (models.py:148-173), never results
    matrix_ids = np.array(matrix_ids, dtype=object)
    widths, max_scores, _ = so.model_constants(weights)
    outputs = list(outputs)
    binned = "binned_counts" in outputs
    frac = float(np.min(threshold)) if binned else threshold
    per_group = {k: [] for k in ("sites", "counts", "binned_counts")}
    n_flipped = 0
    for idx in make_groups(ids, names, group_by):
        lens = {len(seqs[i]) for i in idx}
        if len(lens) != 1:
            raise ValueError("Cannot encode - sequences have unequal length!")  # encode.py:172-176
        gid, gnm, gsn = group_rows(idx, ids, names, sense, rcomp)
        res = so.scan_rows([seqs[i] for i in idx], weights, frac, rcomp=rcomp, order=order, adjudicate=adjudicate)
        n_flipped += res["n_flipped"]
        r, p, m = res["row"], res["pos"], res["motif"]
        if "sites" in outputs:
            df = pd.DataFrame({"id": gid[r], "name": gnm[r], "sense": gsn[r], "start": p,
                               "Matrix_id": matrix_ids[m]})
            df["width"] = widths[m]
            df["end"] = df["start"] + df["width"]
            if score or binned:
                df["score"] = res["score"]
                df["max_score"] = max_scores[m]
                df["frac_score"] = df["score"] / df["max_score"]
            per_group["sites"].append(df)
        if "counts" in outputs or "stats" in outputs:
            if len(r) == 0:
                cdf = pd.DataFrame({"Matrix_id": [], "width": [], "id": [], "name": [], "sense": [], "num": []})
            else:
                ct = pd.crosstab(index=[matrix_ids[m], widths[m]], columns=[gid[r], gnm[r], gsn[r]])
                cdf = ct.melt(ignore_index=False).reset_index()
                cdf.columns = ["Matrix_id", "width", "id", "name", "sense", "num"]
            per_group["counts"].append(cdf)
        if binned:
            fs = res["score"] / max_scores[m]
            cut = pd.cut(fs, np.concatenate([np.asarray(threshold), [1]]))
            ct = pd.crosstab(index=[matrix_ids[m], widths[m]], columns=[gid[r], gnm[r], gsn[r], cut])
            bdf = ct.melt(ignore_index=False).reset_index()
            bdf.columns = ["Matrix_id", "width", "id", "name", "sense", "bin", "num"]
            per_group["binned_counts"].append(bdf)
    out = {k: pd.concat(v).reset_index(drop=True) for k, v in per_group.items() if v}
    if combine_seqs:
        keys = ["Matrix_id", "width", "sense"] + (["id"] if sep_ids else [])
        if "counts" in out:
            out["counts"] = out["counts"].groupby(keys).num.sum().reset_index()
        if "binned_counts" in out:
            bkeys = ["Matrix_id", "width", "sense", "bin"] + (["id"] if sep_ids else [])
            out["binned_counts"] = out["binned_counts"].groupby(bkeys, observed=False).num.sum().reset_index()
        if "stats" in outputs:
            g = out["counts"].groupby(["Matrix_id", "width", "sense"]).num
            st = pd.DataFrame({"len": g.size(), "avg": g.mean(), "sd": g.std(ddof=1)}).reset_index()
            out["stats"] = st
    if "counts" not in outputs and "counts" in out:
        del out["counts"]
    out["_n_flipped"] = n_flipped
    return out
