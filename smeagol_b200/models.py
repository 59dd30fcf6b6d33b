"""PWMModel -- API mirror of smeagol/models.py with the TensorFlow Conv1D replaced by the fused sm_100a scan.

The reference builds a Keras graph ZeroPadding1D -> Embedding(17,4) -> Conv1D over the PWM set (models.py:50-68).
Here the same attributes are kept (Matrix_ids, widths, max_scores, channels, max_width) and the PWM tables live on
the GPU (device.DevicePWMSet).  ``predict`` returns the same dense (N, L, C) float32 array for small inputs and
``predict_with_threshold`` returns the same ((seq_idx, pos_idx, pwm_idx), scores) tuple, produced by the fused
threshold + compaction kernel instead of a dense tensor and np.where.
"""
import numpy as np

from . import device as dev
from .encode import integer_encode


class PWMModel:
    """Holds a PWM library for scanning.

    Parameters (as the reference, models.py:21-33): Matrix_ids, widths, max_scores, channels, max_width.
    adjudicate: sites whose fp32 score lies within 1e-5 (relative to sum_k max_b |T[k,b]|) of the threshold are
    re-scored in fp64 on the device and decided there (north_star); False reproduces the reference's pure
    `(double)score_f32 > thr` compare (models.py:108).
    """

    def __init__(self, pwm_df, adjudicate=True, tc_min_width=None):
        self.Matrix_ids = np.array(pwm_df.Matrix_id)
        self.weights = [np.asarray(w) for w in pwm_df.weights]
        self.widths = np.array([w.shape[0] for w in self.weights])
        # dtype of the supplied weights, like pwm_df.weights.apply(lambda x: np.max(x, axis=1).sum()) (models.py:42-44)
        self.max_scores = np.array([np.max(w, axis=1).sum() for w in self.weights])
        self.channels = len(self.weights)
        self.max_width = max(self.widths)
        self.adjudicate = adjudicate
        self.tc_min_width = tc_min_width  # hybrid engine split (device.DevicePWMSet), None = library default
        self._dev = None
        self.last_scan_stats = None

    # ------------------------------------------------------------------ device tables (lazy: needs a GPU)
    @property
    def device_set(self):
        if self._dev is None:
            self._dev = dev.DevicePWMSet(self.weights, tc_min_width=self.tc_min_width)
        return self._dev

    def thresholds(self, frac_threshold):
        """Per-PWM score thresholds, computed exactly as the reference does (models.py:140-141)."""
        assert (frac_threshold >= 0) & (frac_threshold <= 1)
        return np.asarray(frac_threshold * self.max_scores, dtype=np.float64)

    # ------------------------------------------------------------------ reference API
    def _batch_from_rows(self, seqs):
        if len(seqs) and type(seqs[0]) == str:
            return dev.SeqBatch(list(seqs), kind="ascii"), len(seqs[0])
        arr = np.atleast_2d(np.asarray(seqs))
        codes = arr.astype(np.uint8)
        return dev.SeqBatch([codes[i] for i in range(codes.shape[0])], kind="codes"), codes.shape[1]

    def predict(self, seqs):
        """Dense scores of integer-encoded rows: (N, L) -> (N, L, C) float32 (models.py:92-100), including the
        zero-padded overhanging windows.  Meant for small inputs and tests; the scan path never materialises it."""
        batch, L = self._batch_from_rows(seqs)
        return dev.predict_dense(self.device_set, batch, L).cpu().numpy()

    def predict_with_threshold(self, seqs, frac_threshold, score=False, seq_batch=0):
        """Scan rows and threshold (models.py:119-178).  Rows are scanned as given (forward only), exactly as the
        reference does after SeqEncoding has materialised reverse-complement rows.  seq_batch is accepted for API
        compatibility; batching never changes results (models.py:148-173) and the device kernel does not need it.

        Returns ((seq_idx, pos_idx, pwm_idx), scores) in row-major (row, pos, motif) order; scores None if not score.
        """
        thr = self.thresholds(frac_threshold)
        if len(seqs) and type(seqs[0]) == str:
            seqs = np.vstack([integer_encode(seq, rc=False) for seq in seqs])
        batch, L = self._batch_from_rows(seqs)
        n = batch.n_rows
        out_rows = np.stack([np.arange(n, dtype=np.int32), np.full(n, 0, dtype=np.int32)], axis=1)
        res = dev.scan(self.device_set, batch, thr, dev.STRANDS["none"], out_rows, n, want_hits=True, want_counts=False,
                       adjudicate=self.adjudicate)
        self.last_scan_stats = res.stats
        row, pos, motif, sc = res.hits_host()
        return (row, pos, motif), (sc if score else None)
