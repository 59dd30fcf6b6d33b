"""TEST / BASELINE INFRASTRUCTURE ONLY -- CPU port of the reference's own scan pipeline, used as the
``cpu_baseline`` and ``--impl reference`` arm of bench.py (kind "port": TensorFlow is not installable here, so the
Keras Conv1D of smeagol/models.py:50-68 is evaluated with torch's CPU conv1d, multi-threaded oneDNN like TF's).

It keeps the STRUCTURE of the reference path, because that structure is what its run time is made of:
  1. per-base Python dict encoding of every record and of its reverse complement (encode.py:66, :63-64);
  2. one dense (N, L, C) fp32 conv per SeqEncoding group, in row batches of 32 (Keras predict default);
  3. np.where over the dense tensor against fp64 thresholds, end trim, score gather (models.py:108-116);
  4. pandas assembly of sites / counts per group, concat (scan.py:40-147, 255-266).
Never imported by the product package.
"""
import time

import numpy as np
import pandas as pd
import torch

from . import scan_oracle as so

_ENC = {c: i for i, c in enumerate(so.CODES)}


def _integer_encode(seq, rc=False):
    if rc:
        seq = so.reverse_complement_str(seq)
    return np.array([_ENC[b] for b in seq], dtype="float32")


class PortModel:
    """PWMModel of the reference restated on torch CPU (models.py:35-100)."""

    def __init__(self, matrix_ids, weights):
        self.Matrix_ids = np.array(matrix_ids)
        self.widths = np.array([w.shape[0] for w in weights])
        self.max_scores = np.array([np.max(w, axis=1).sum() for w in weights])
        self.channels = len(weights)
        self.max_width = int(max(self.widths))
        k = np.zeros((self.max_width, 4, self.channels), dtype=np.float32)  # left-aligned, tail zero-padded
        for m, w in enumerate(weights):
            k[: w.shape[0], :, m] = w
        self.kernel = torch.from_numpy(np.ascontiguousarray(k.transpose(2, 1, 0)))  # (C, 4, Wmax)
        self.emb = torch.from_numpy(so.EMB32)

    def predict(self, seqs, batch_size=32):
        seqs = np.atleast_2d(seqs)
        outs = []
        for i in range(0, seqs.shape[0], batch_size):
            x = torch.from_numpy(seqs[i:i + batch_size].astype(np.int64))
            x = torch.nn.functional.pad(x, (0, self.max_width - 1))            # ZeroPadding1D (code 0 = 'Z')
            oh = self.emb[x].permute(0, 2, 1).contiguous()                      # Embedding -> (N, 4, L+W-1)
            outs.append(torch.nn.functional.conv1d(oh, self.kernel).permute(0, 2, 1).contiguous().numpy())
        return np.concatenate(outs, axis=0)

    def predict_with_threshold(self, seqs, frac_threshold, score=False):
        thresholds = frac_threshold * self.max_scores
        predictions = self.predict(seqs)
        thresholded = np.where(predictions > thresholds)
        select = thresholded[1] + self.widths[thresholded[2]] <= seqs.shape[1]
        thresholded = tuple(x[select] for x in thresholded)
        return thresholded, (predictions[thresholded] if score else None)


def scan_sequences(seqs, ids, names, model, threshold, sense="+", rcomp="both", outputs=("sites", "counts"),
                   score=True, timings=None):
    """group_by=None path of scan.scan_sequences (one group per record).  `timings` (dict) accumulates the stage
    split encode / conv+where / pandas in seconds."""
    t = timings if timings is not None else {}
    sites, counts = [], []
    comp = so.SENSE_COMPLEMENT[sense]
    for s, i, n in zip(seqs, ids, names):
        t0 = time.perf_counter()
        if rcomp == "both":
            rows = np.vstack([_integer_encode(s, False), _integer_encode(s, True)])
            gid, gnm, gsn = np.array([i, i]), np.array([n, n]), np.array([sense, comp])
        elif rcomp == "none":
            rows = np.vstack([_integer_encode(s, False)])
            gid, gnm, gsn = np.array([i]), np.array([n]), np.array([sense])
        else:
            rows = np.vstack([_integer_encode(s, True)])
            gid, gnm, gsn = np.array([i]), np.array([n]), np.array([comp])
        t1 = time.perf_counter()
        (r, p, m), sc = model.predict_with_threshold(rows, threshold, score)
        t2 = time.perf_counter()
        if "sites" in outputs:
            df = pd.DataFrame({"id": gid[r], "name": gnm[r], "sense": gsn[r], "start": p, "Matrix_id": model.Matrix_ids[m]})
            df["width"] = model.widths[m]
            df["end"] = df["start"] + df["width"]
            if sc is not None:
                df["score"] = sc
                df["max_score"] = model.max_scores[m]
                df["frac_score"] = df["score"] / df["max_score"]
            sites.append(df)
        if "counts" in outputs and len(r):
            ct = pd.crosstab(index=[model.Matrix_ids[m], model.widths[m]], columns=[gid[r], gnm[r], gsn[r]])
            cdf = ct.melt(ignore_index=False).reset_index()
            cdf.columns = ["Matrix_id", "width", "id", "name", "sense", "num"]
            counts.append(cdf)
        t3 = time.perf_counter()
        t["encode"] = t.get("encode", 0.0) + (t1 - t0)
        t["conv_where"] = t.get("conv_where", 0.0) + (t2 - t1)
        t["pandas"] = t.get("pandas", 0.0) + (t3 - t2)
    out = {}
    if sites:
        out["sites"] = pd.concat(sites).reset_index(drop=True)
    if counts:
        out["counts"] = pd.concat(counts).reset_index(drop=True)
    return out
