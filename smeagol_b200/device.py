"""Device-side objects of the scan path: packed sequence batches, PWM tables and the fused scan driver.

Everything here talks to libsmeagol_b200.so through the C-ABI (smeagol_b200._lib); torch is used only for device
memory, streams and (in dist.py) torch.distributed.  There is no CPU fallback.
"""
import ctypes
import os

import numpy as np
import torch

from . import _lib

ROW_DTYPE = np.dtype(_lib.ROW_DTYPE_FIELDS)
CHUNK_DTYPE = np.dtype([("row", "<i4"), ("start", "<i4")])
NEAR_DTYPE = np.dtype([("row", "<i4"), ("motif", "<i4"), ("strand", "<i4"), ("start", "<i4"), ("score", "<f4"),
                       ("reserved", "<i4")])
ROW_PAD_BASES = 96  # every row is followed by at least this many 'Z' bases (halo + one prefetched word)
NEAR_REL = 1e-5     # half-width of the near-threshold band, relative to sum_k max_b |T[k,b]|
STRANDS = {"none": 1, "only": 2, "both": 3}
DEFAULT_ENGINE = "regtab"
DEFAULT_TC_KIND = "tc5"
SEED_MIN_LEN = 11     # shortest seed of the seed engine (csrc/kmer.h); 12 is the longest the index supports
SEED_MAX_REST = 14    # a PWM is seeded when W <= seed length + this
KMER_MAX_WIDTH = 12   # widest class of the k-mer index engine (csrc/kmer.cu)
KMER_ALPHA = 1.5     # a width class joins the k-mer index when 4^W <= KMER_ALPHA * window starts (break-even of index build vs tcgen05 scan, profiles/r02_bench_widths.txt)
                      # k-mers once costs about as much per (k-mer, PWM) as any engine spends per (position, PWM)  # tensor-core prefilter of the 'auto' / tensor-core engines: 'hmma' (mma.sync) or 'tc5' (tcgen05 + TMEM)
DEFAULT_TC_MIN_WIDTH = 7  # measured per width (profiles/r01_bench_widths.txt): the register-table engine wins up to W = 6 (hit density >= 6e-3 at thr 0.7), the tensor-core prefilter from W = 7


def _require_cuda(device=None):
    if not torch.cuda.is_available():
        raise _lib.SmeagolB200Error("smeagol_b200 needs a CUDA device (B200, sm_100a); there is no CPU fallback")
    return torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)


def _ptr(t):
    return ctypes.c_void_p(t.data_ptr()) if t is not None else ctypes.c_void_p(0)


def _stream():
    return ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)


def _to_dev(arr, device):
    """numpy (any dtype / structured) -> device uint8 tensor holding the same bytes."""
    raw = np.ascontiguousarray(arr).view(np.uint8).reshape(-1)
    return torch.from_numpy(raw).to(device, non_blocking=False)


def _as_bytes(seq):
    if isinstance(seq, (bytes, bytearray, memoryview)):
        return np.frombuffer(seq, dtype=np.uint8)
    if isinstance(seq, np.ndarray):
        return np.ascontiguousarray(seq, dtype=np.uint8).reshape(-1)
    return np.frombuffer(str(seq).encode("latin-1"), dtype=np.uint8)


def flatten_rows(arrs, pinned=True):
    """Copy uint8 rows into one host buffer with 16-byte aligned row starts -> (tensor, src_off)."""
    lens = np.array([a.shape[0] for a in arrs], dtype=np.int64)
    padded = (lens + 15) // 16 * 16
    src_off = np.concatenate([[0], np.cumsum(padded)[:-1]]).astype(np.int64)
    total_src = int(padded.sum())
    host = torch.empty(max(total_src, 16), dtype=torch.uint8, pin_memory=pinned)
    hnp = host.numpy()
    for a, o in zip(arrs, src_off):
        hnp[o:o + a.shape[0]] = a
    return host, src_off


class SeqBatch:
    """Rows packed on the GPU at 2 bits/base + 1-bit ambiguity mask (+ uint8 codes only if ambiguity exists).

    Replaces encode.integer_encode / SeqEncoding's (N, L) float32 array (smeagol/encode.py:47-68, 109-160).
    seqs: list of str / bytes / uint8 arrays (ASCII, or integer codes 0..16 when kind='codes').
    segments: optional list of (g_off, g_len, n_own) per row for rows that are segments of a longer sequence
    (chunked or multi-GPU scans); n_avail is the row's length (owned part + right halo).
    """

    def __init__(self, seqs, kind="ascii", segments=None, device=None, pinned=True, flat=None):
        """flat=(host_uint8_tensor, src_off, lens): rows already laid out in one (pinned) host buffer, row r at
        src_off[r] (multiples of 16), so that no host-side gather is needed (used by the end-to-end benchmark)."""
        self.device = _require_cuda(device)
        lib = _lib.load()
        if flat is None:
            arrs = [_as_bytes(s) for s in seqs]
            if len(arrs) == 0:
                raise ValueError("SeqBatch needs at least one row")
            lens = np.array([a.shape[0] for a in arrs], dtype=np.int64)
            host, src_off = flatten_rows(arrs, pinned)
        else:
            host, src_off, lens = flat
            lens = np.asarray(lens, dtype=np.int64)
            src_off = np.asarray(src_off, dtype=np.int64)
            arrs = None
        n_rows = len(lens)
        total_src = host.numel()
        if lens.max() >= 2 ** 31 - 256:
            raise ValueError("rows longer than 2^31 bases must be split into segments")
        rows = np.zeros(n_rows, dtype=ROW_DTYPE)
        row_words = ((lens + ROW_PAD_BASES + 63) // 64) * 4
        rows["word_off"] = np.concatenate([[0], np.cumsum(row_words)[:-1]])
        rows["n_avail"] = lens
        if segments is None:
            rows["g_off"] = 0
            rows["g_len"] = lens
            rows["n_own"] = lens
        else:
            seg = np.asarray(segments, dtype=np.int64).reshape(n_rows, 3)
            rows["g_off"], rows["g_len"], rows["n_own"] = seg[:, 0], seg[:, 1], seg[:, 2]
        self.total_words = int(row_words.sum())
        self.h2d_bytes = int(total_src)
        d_src = host.to(self.device, non_blocking=True)
        self.rows_host = rows
        self.n_rows = n_rows
        self.lens = lens
        self.d_rows = _to_dev(rows, self.device)
        d_src_off = torch.from_numpy(src_off).to(self.device)
        self.packed = torch.empty(self.total_words, dtype=torch.int32, device=self.device)
        self.amask = torch.empty(self.total_words, dtype=torch.int16, device=self.device)
        codes = torch.empty(self.total_words * 16, dtype=torch.uint8, device=self.device)
        status = torch.tensor([-1, 0], dtype=torch.int64, device=self.device)
        _lib.check(lib.smg_pack(_ptr(d_src), 0 if kind == "ascii" else 1, _ptr(self.d_rows), _ptr(d_src_off), n_rows,
                                self.total_words, _ptr(self.packed), _ptr(self.amask), _ptr(codes), _ptr(status), _stream()),
                   "smg_pack")
        st = status.cpu().numpy()
        if st[0] >= 0:
            bad = int(host[int(st[0])])
            raise KeyError(chr(int(bad)) if kind == "ascii" else int(bad))
        self.has_ambiguity = bool(st[1])
        self.codes = codes if self.has_ambiguity else None
        self.total_own = int(rows["n_own"].clip(min=0).sum())
        self.max_g_len = int(rows["g_len"].max())

    def chunks(self, chunk_len):
        """Work list of (row, first owned start) pairs, chunk_len a multiple of 16."""
        n_own = self.rows_host["n_own"].astype(np.int64).clip(min=0)
        per_row = (n_own + chunk_len - 1) // chunk_len
        total = int(per_row.sum())
        ch = np.zeros(total, dtype=CHUNK_DTYPE)
        if total:
            ch["row"] = np.repeat(np.arange(self.n_rows, dtype=np.int32), per_row)
            first = np.repeat(np.cumsum(per_row) - per_row, per_row)
            ch["start"] = ((np.arange(total, dtype=np.int64) - first) * chunk_len).astype(np.int32)
        return ch


class DevicePWMSet:
    """Device tables of a PWM library (models.py:70-90 pads PWMs into one Conv1D kernel; here: natural fp32 tables,
    reverse-complement tables and lane-major register-table groups, built by smg_pwmset_create)."""

    def __init__(self, weights, device=None, tc_min_width=None, kmer_max_width=0):
        """kmer_max_width > 0: PWMs up to that width are left to the k-mer index engine (they are in no lane group or
        tensor-core table of this set; ScanPlan asks for such a sibling through with_kmer()).
        Non-finite weights (log2 of a zero probability) are accepted as in the reference, where such a PWM never yields
        a site (ScanPlan switches it off).  tc_min_width: PWMs at least this wide go to the tensor-core engine, narrower ones to the
        register-table engine (engine 'auto'); 0 = no split (either engine alone covers every PWM)."""
        self.device = _require_cuda(device)
        if tc_min_width is None:
            tc_min_width = int(os.environ.get("SMEAGOL_B200_TC_MIN_WIDTH", DEFAULT_TC_MIN_WIDTH))
        self.tc_min_width = int(tc_min_width)
        self.kmer_max_width = int(kmer_max_width) if self.tc_min_width > 0 else 0
        self._siblings = {}
        lib = _lib.load()
        w32 = [np.ascontiguousarray(np.asarray(w), dtype=np.float32) for w in weights]
        self._w32 = w32
        for w in w32:
            if w.ndim != 2 or w.shape[1] != 4:
                raise ValueError("PWM weights must have shape (W, 4)")
        self.n_motifs = len(w32)
        self.widths = np.array([w.shape[0] for w in w32], dtype=np.int32)
        flat = np.concatenate([w.reshape(-1) for w in w32]).astype(np.float32)
        if np.any(np.isfinite(flat) & (np.abs(flat) > 6e4)):
            self.tc_min_width = self.kmer_max_width = 0  # fp16 prefilter tables cannot hold such weights: FP32 engine only
        self.abs_scale = np.array([np.abs(w.astype(np.float64)).max(axis=1).sum() for w in w32])
        handle = ctypes.c_void_p()
        with torch.cuda.device(self.device):
            kmin = self.kmer_max_width + 1 if self.kmer_max_width > 0 else 0
            _lib.check(lib.smg_pwmset_create_split3(flat.ctypes.data_as(ctypes.c_void_p),
                                                    self.widths.ctypes.data_as(ctypes.c_void_p), self.n_motifs, kmin,
                                                    max(self.tc_min_width, kmin) if self.tc_min_width > 0 else 0,
                                                    ctypes.byref(handle)), "smg_pwmset_create_split3")
        self.handle = handle
        info = (ctypes.c_int64 * 8)()
        _lib.check(lib.smg_pwmset_info(handle, info), "smg_pwmset_info")
        self.max_width = int(info[1])
        self.n_groups = int(info[2])
        self.n_buckets = int(info[3])
        self.n_wide = int(info[4])
        self.sum_padded_width = int(info[5])

    def with_kmer(self, kmer_max_width):
        """Sibling set whose PWMs up to kmer_max_width are left to the k-mer index engine (built once, cached)."""
        if kmer_max_width == self.kmer_max_width:
            return self
        if kmer_max_width not in self._siblings:
            self._siblings[kmer_max_width] = DevicePWMSet(self._w32, device=self.device, tc_min_width=self.tc_min_width,
                                                          kmer_max_width=kmer_max_width)
        return self._siblings[kmer_max_width]

    def close(self):
        for sib in getattr(self, "_siblings", {}).values():
            sib.close()
        self._siblings = {}
        if getattr(self, "handle", None):
            _lib.load().smg_pwmset_destroy(self.handle)
            self.handle = None

    def __del__(self):  # pragma: no cover
        try:
            self.close()
        except Exception:
            pass


def _round_down_f32(x):
    f = x.astype(np.float32)
    up = f.astype(np.float64) > x
    f[up] = np.nextafter(f[up], np.float32(-np.inf))
    return f


def _round_up_f32(x):
    f = x.astype(np.float32)
    dn = f.astype(np.float64) < x
    f[dn] = np.nextafter(f[dn], np.float32(np.inf))
    return f


def threshold_bounds(thr64, abs_scale, adjudicate=True):
    """fp32 kernel bounds for `score > thr`: with adjudication a band of NEAR_REL * abs_scale around thr goes to the
    fp64 adjudicator; without it thr_lo == thr_hi == largest float <= thr reproduces `(double)score_f32 > thr`
    (models.py:108) exactly."""
    thr64 = np.asarray(thr64, dtype=np.float64)
    if adjudicate:
        band = NEAR_REL * np.asarray(abs_scale, dtype=np.float64)
        return _round_down_f32(thr64 - band), _round_up_f32(thr64 + band)
    lo = _round_down_f32(thr64)
    return lo, lo.copy()


def _bits(n):
    return max(1, int(n - 1).bit_length()) if n > 1 else 1


class ScanResult:
    """Sorted hits (device), count table (device) and bookkeeping of one fused scan."""

    def __init__(self, keys, scores, n_hits, counts, pos_bits, motif_bits, stats):
        self.keys, self.scores, self.n_hits, self.counts = keys, scores, n_hits, counts
        self.pos_bits, self.motif_bits, self.stats = pos_bits, motif_bits, stats

    def hits_host(self):
        """(out_row, pos, motif) int64 arrays + float32 scores, in (row, pos, motif) order."""
        if self.keys is None or self.n_hits == 0:
            z = np.zeros(0, dtype=np.int64)
            return z, z.copy(), z.copy(), np.zeros(0, dtype=np.float32)
        k = self.keys[:self.n_hits].cpu().numpy().view(np.uint64)
        sc = self.scores[:self.n_hits].cpu().numpy()
        motif = (k & np.uint64((1 << self.motif_bits) - 1)).astype(np.int64)
        pos = ((k >> np.uint64(self.motif_bits)) & np.uint64((1 << self.pos_bits) - 1)).astype(np.int64)
        row = (k >> np.uint64(self.motif_bits + self.pos_bits)).astype(np.int64)
        return row, pos, motif, sc

    def key_checksum(self):
        """Order-independent checksum of the hit set (sum and xor of the keys), computed on the device."""
        if self.keys is None or self.n_hits == 0:
            return 0, 0
        k = self.keys[:self.n_hits]
        x = k.clone()
        while x.numel() > 1:  # xor-reduce
            h = x.numel() // 2
            y = torch.bitwise_xor(x[:h], x[h:2 * h])
            x = torch.cat([y, x[2 * h:]]) if x.numel() % 2 else y
        return int(k.sum().item()), int(x.item())


def pick_chunk_len(total_own, n_groups, target_ctas=148 * 16 * 8, n_rows=1):
    """Owned window starts per CTA: enough CTAs to fill the GPU several times over, at most 4096 -- 2048 when the batch
    is many short rows (measured on config c2: 1,000 genomes of 5-15 kb, +2.7 %; the tail of every row is a partial
    chunk)."""
    if os.environ.get("SMEAGOL_B200_CHUNK_LEN"):  # tuning only
        return int(os.environ["SMEAGOL_B200_CHUNK_LEN"])
    c = (total_own * max(n_groups, 1)) // max(target_ctas, 1)
    c = ((c + 15) // 16) * 16
    cap = 2048 if total_own < 32768 * max(n_rows, 1) else 4096
    return int(min(cap, max(256, c)))


def choose_index_engines(widths, tc_min_width, total_own, forced_kmer=None, forced_seed=None, allow_seed=True):
    """Which per-call indexes a scan of `total_own` window starts uses (host logic, no GPU): returns (wk, kmer_w, seed_len).

    wk: widest class the k-mer index may take, 4^W <= KMER_ALPHA * window starts (the measured break-even of the index
    build against the tcgen05 scan, profiles/r02_bench_widths.txt); kmer_w: widest PRESENT width <= wk (0: no k-mer
    side).  seed_len: 0, or the seed length (11 / 12) when the seed engine is cheaper than the tensor-core prefilter for
    the PWMs on the tensor-core side.  Cost model from the c4 / c2 steps and the per-width matrix: the per-call seed
    index costs 0.7e-13 s per (12-mer, column) and 1.4e-13 s per (11-mer, column) -- the bound prunes better on longer
    k-mers --; its scan + exact pass 0.8e-11 s per position + 1.2e-14 s per (position,
    column) with 12-base seeds, twice that with 11 (candidates double per base the seed loses); the tcgen05 prefilter
    1.1 .. 1.7e-13 s per (position, column).  Seeds win on long scans and lose on short ones."""
    widths = np.asarray(widths)
    wk = 0
    for w in range(1, KMER_MAX_WIDTH + 1):
        if 4 ** w <= KMER_ALPHA * max(total_own, 1):
            wk = w
    if forced_kmer is not None:
        wk = max(0, min(KMER_MAX_WIDTH, int(forced_kmer)))
    present = np.unique(widths)
    present = present[present <= wk]
    kmer_w = int(present.max()) if present.size else 0
    seed_len = 0
    if allow_seed and wk >= SEED_MIN_LEN:
        tc_w = widths[(widths > max(kmer_w, tc_min_width - 1)) & (widths <= 32)]
        nv, L = 2 * int(tc_w.size), float(max(total_own, 1))
        best = None
        for S in range(min(wk, KMER_MAX_WIDTH), SEED_MIN_LEN - 1, -1):
            if nv == 0 or int(tc_w.min()) <= S or int(tc_w.max()) > S + SEED_MAX_REST:
                continue
            cost = nv * 4.0 ** S * 0.7e-13 * 2.0 ** (KMER_MAX_WIDTH - S) + L * (0.8e-11 + nv * 1.2e-14 * 2.0 ** (KMER_MAX_WIDTH - S))
            if best is None or cost < best[0]:
                best = (cost, S)
        if forced_seed is not None and nv:
            best = (0.0, int(forced_seed))
        if best is not None and best[0] < L * nv * 1.4e-13:
            seed_len = best[1]
    return wk, kmer_w, seed_len


class ScanPlan:
    """A prepared fused scan: all device buffers allocated once, kernels enqueued without host synchronisation.

    launch() enqueues (counter/count reset, smg_scan, smg_adjudicate) on the current stream; finish() reads the
    counters (one small D2H), regrows and relaunches on overflow (never truncates), and sorts the hits.
    """

    def __init__(self, pwmset, batch, thr64, strands, out_rows, n_out_rows, want_hits=True, want_counts=True,
                 adjudicate=True, chunk_len=None, hit_cap=None, near_cap=None, engine=None):
        """engine: 'regtab' (fused FP32 register-table kernel), 'hmma' (mma.sync tensor-core prefilter + exact second
        pass) or 'tc5' (tcgen05 / TMEM prefilter + the same exact second pass); identical results.  Default:
        $SMEAGOL_B200_ENGINE or DEFAULT_ENGINE; a PWM set built with a width split runs 'auto' (register-table kernel below
        the split, the prefilter $SMEAGOL_B200_TC_KIND / DEFAULT_TC_KIND above it)."""
        self.lib = _lib.load()
        self.engine = engine or os.environ.get("SMEAGOL_B200_ENGINE", DEFAULT_ENGINE)
        self.tc_kind = os.environ.get("SMEAGOL_B200_TC_KIND", DEFAULT_TC_KIND)
        if self.engine == "tc5":      # 'hmma' / 'tc5' name the prefilter explicitly; 'auto' takes the default kind
            self.engine, self.tc_kind = "hmma", "tc5"
        elif self.engine == "hmma":
            self.tc_kind = "hmma"
        if pwmset.tc_min_width > 0:
            self.engine = "auto"    # the PWM set was built split: both engines run, each on its side
        elif self.engine == "auto" or (self.engine == "hmma" and pwmset.n_wide > 0):
            self.engine = "regtab"  # PWMs wider than 32 only exist in the register-table / generic path
        # ---- k-mer index engine for the narrow width classes (engine 'auto' only)
        self.kmer_w = 0
        self.seed_len = 0
        if self.engine == "auto" and os.environ.get("SMEAGOL_B200_KMER", "1") != "0":
            wk, self.kmer_w, want_seed = choose_index_engines(
                pwmset.widths, pwmset.tc_min_width, batch.total_own, os.environ.get("SMEAGOL_B200_KMER_MAX_WIDTH"),
                os.environ.get("SMEAGOL_B200_SEED_LEN"),
                os.environ.get("SMEAGOL_B200_SEED", "1") != "0" and "SMEAGOL_B200_TC_KIND" not in os.environ)
            pwmset = pwmset.with_kmer(self.kmer_w)
            if want_seed:
                info = (ctypes.c_int64 * 4)()
                _lib.check(self.lib.smg_seed_prepare(pwmset.handle, want_seed, info), "smg_seed_prepare")
                if info[3] == 1 and info[0] > 0:
                    self.seed_len, self.seed_nv, self._seed_hl = want_seed, int(info[0]), int(info[1])
                    self.tc_kind = "seed"
        self.pwmset, self.batch = pwmset, batch
        dev = batch.device
        M = pwmset.n_motifs
        thr64 = np.ascontiguousarray(thr64, dtype=np.float64)
        assert thr64.shape == (M,)
        # A PWM with a non-finite weight (log2 of a zero probability) never scores a hit in the reference: its one-hot
        # contraction multiplies every weight of a row, 0 * -inf = NaN, and NaN > thr is False (models.py:52-68, :108).
        # The kernels add only the selected weight, so such PWMs are switched off through their thresholds.
        dead = ~np.isfinite(np.asarray(pwmset.abs_scale, dtype=np.float64)) | ~np.isfinite(thr64)
        scale = np.where(dead, 0.0, np.asarray(pwmset.abs_scale, dtype=np.float64))
        thr64 = np.where(dead, np.inf, thr64)
        lo, hi = threshold_bounds(thr64, scale, adjudicate)
        self.motif_bits = _bits(M)
        self.pos_bits = _bits(batch.max_g_len + 1)
        self.row_bits = _bits(n_out_rows)
        if self.motif_bits + self.pos_bits + self.row_bits > 63:
            raise ValueError("hit key does not fit in 63 bits (rows=%d, length=%d, motifs=%d)"
                             % (n_out_rows, batch.max_g_len, M))
        self.chunk_len = (pick_chunk_len(batch.total_own, pwmset.n_groups, n_rows=batch.n_rows) if chunk_len is None
                          else int(chunk_len))
        if chunk_len is None and self.engine in ("hmma", "auto") and self.tc_kind in ("tc5", "seed"):  # (seed: it may fall back)
            self.chunk_len = max(512, (self.chunk_len + 511) // 512 * 512)  # the tcgen05 prefilter works in tiles of 512 starts
        chunks = batch.chunks(self.chunk_len)
        self.n_chunks = len(chunks)
        self.d_chunks = _to_dev(chunks, dev) if len(chunks) else torch.zeros(8, dtype=torch.uint8, device=dev)
        self.d_lo = torch.from_numpy(lo).to(dev)
        self.d_hi = torch.from_numpy(hi).to(dev)
        self.d_thr = torch.from_numpy(thr64).to(dev)
        self.d_out_rows = torch.from_numpy(np.ascontiguousarray(out_rows, dtype=np.int32)).to(dev)
        self.strands = int(strands)
        self.adjudicate = adjudicate
        self.want_hits, self.want_counts = want_hits, want_counts
        cells = batch.total_own * M * bin(self.strands).count("1")
        self.hit_cap = int(min(max(1 << 16, cells * 0.004), 1 << 31)) if hit_cap is None else int(hit_cap)
        self.near_cap = (1 << 16) if near_cap is None else int(near_cap)
        self.counts = torch.zeros((batch.n_rows, M, 2), dtype=torch.int32, device=dev) if want_counts else None
        self.counters = torch.zeros(_lib.N_COUNTERS, dtype=torch.int64, device=dev)
        self.cand_cap = 0
        if self.engine in ("hmma", "auto"):
            # prefilter bound: below thr_lo by 2^-9 * sum_k max_b |T| (>= 4x the fp16 rounding error of the operands)
            pre = _round_down_f32(lo.astype(np.float64) - 2.0 ** -9 * scale - 1e-30)
            self.d_pre = torch.from_numpy(pre).to(dev)
            self.cand_pos_bits = _bits(int(batch.rows_host["n_avail"].max()) + 2)
            if self.cand_pos_bits + self.motif_bits + 1 + _bits(batch.n_rows) > 62:
                raise ValueError("candidate key does not fit in 62 bits")
            # tc5: every epilogue warp of every CTA keeps two blocks of 128 candidate slots reserved
            self.cand_cap = int(self.hit_cap * 1.25) + 4096 + ((1 << 20) if self.tc_kind == "tc5" else 0)
            if self.tc_kind == "seed":  # ~0.1 (12-mer seeds) .. 0.5 (10-mers) candidates per position and strand
                per = {12: 0.15, 11: 0.35}.get(self.seed_len, 0.7)
                self.cand_cap = int(batch.total_own * 2 * per) + (1 << 16)
        self._alloc()
        if self.kmer_w > 0:
            self._build_kmer_index()
        if self.seed_len > 0:
            self._build_seed_index()

    def _build_seed_index(self):
        """One-class k-mer index of the wide PWMs' seeds for this plan's thresholds (csrc/kmer.h: seed engine)."""
        dev, lib = self.batch.device, self.lib
        self.seed_head = torch.empty(max(self._seed_hl, 1), dtype=torch.int32, device=dev)
        cap = max(1 << 16, min(int(4.0 ** self.seed_len * 0.15), 1 << 27))
        ctr = torch.zeros(_lib.N_COUNTERS, dtype=torch.int64, device=dev)
        while True:
            self.seed_entries = torch.empty(cap * 16, dtype=torch.uint8, device=dev)
            ctr.zero_()
            _lib.check(lib.smg_seed_index_build(self.pwmset.handle, self.seed_len, self.strands, _ptr(self.d_lo), _ptr(self.seed_head),
                                                self._seed_hl, _ptr(self.seed_entries), cap, _ptr(ctr), _stream()), "smg_seed_index_build")
            need = int(ctr[_lib.CTR_KMER_ENTRIES].item())
            if need <= cap:
                break
            if need > (1 << 27):
                # a threshold so low that the seeds are not selective (> 2 GB of entries): the tensor-core prefilter takes over
                self.seed_len, self.tc_kind = 0, os.environ.get("SMEAGOL_B200_TC_KIND", DEFAULT_TC_KIND)
                self.seed_head = self.seed_entries = None
                need_cap = int(self.hit_cap * 1.25) + 4096 + (1 << 20)
                if need_cap > self.cand_cap:
                    self.cand_cap = need_cap
                    self.cand = torch.empty(self.cand_cap, dtype=torch.int64, device=dev)
                return
            cap = min(int(need * 1.1) + 1024, (1 << 31) - 2)
        self.seed_cap, self.seed_n_entries, self._seed_ctr = cap, need, ctr

    def _build_kmer_index(self):
        """Score every k-mer of the narrow width classes once for this plan's thresholds (csrc/kmer.cu)."""
        dev, lib = self.batch.device, self.lib
        hl = ctypes.c_int64(0)
        _lib.check(lib.smg_kmer_index_size(self.pwmset.handle, self.kmer_w, ctypes.byref(hl)), "smg_kmer_index_size")
        self.kmer_head = torch.empty(max(int(hl.value), 1), dtype=torch.int32, device=dev)
        w = self.pwmset.widths.astype(np.int64)
        est = int(sum(4.0 ** int(x) * 2 * 3e-3 for x in w[w <= self.kmer_w]))
        cap = max(1 << 16, min(est, 1 << 26))
        ctr = torch.zeros(_lib.N_COUNTERS, dtype=torch.int64, device=dev)
        while True:
            self.kmer_entries = torch.empty(cap * 16, dtype=torch.uint8, device=dev)
            ctr.zero_()
            _lib.check(lib.smg_kmer_index_build(self.pwmset.handle, self.kmer_w, self.strands, _ptr(self.d_lo), _ptr(self.d_hi),
                                                _ptr(self.d_thr), _ptr(self.kmer_head), int(hl.value), _ptr(self.kmer_entries), cap,
                                                _ptr(ctr), _stream()), "smg_kmer_index_build")
            need = int(ctr[_lib.CTR_KMER_ENTRIES].item())
            if need <= cap:
                break
            if need >= (1 << 31) - 1:
                raise _lib.SmeagolB200Error("k-mer index too large (threshold too low for the k-mer engine)")
            cap = min(int(need * 1.1) + 1024, (1 << 31) - 2)
        self.kmer_cap, self.kmer_n_entries = cap, need
        self._kmer_hl, self._kmer_ctr = int(hl.value), ctr

    def launch_kmer_build(self):
        """Enqueue the k-mer index build again (same thresholds, known capacity, no host sync): what a fresh call pays
        per (PWM set, thresholds) -- benchmarks put it inside the timed step."""
        if self.seed_len > 0:
            self._seed_ctr.zero_()
            _lib.check(self.lib.smg_seed_index_build(self.pwmset.handle, self.seed_len, self.strands, _ptr(self.d_lo),
                                                     _ptr(self.seed_head), self._seed_hl, _ptr(self.seed_entries), self.seed_cap,
                                                     _ptr(self._seed_ctr), _stream()), "smg_seed_index_build")
        if self.kmer_w <= 0:
            return
        self._kmer_ctr.zero_()
        _lib.check(self.lib.smg_kmer_index_build(self.pwmset.handle, self.kmer_w, self.strands, _ptr(self.d_lo), _ptr(self.d_hi),
                                                 _ptr(self.d_thr), _ptr(self.kmer_head), self._kmer_hl, _ptr(self.kmer_entries),
                                                 self.kmer_cap, _ptr(self._kmer_ctr), _stream()), "smg_kmer_index_build")

    def launch_kmer_side(self):
        """K-mer index engine on the width classes up to kmer_w (hits, counts and tallies like the other engines)."""
        if self.kmer_w <= 0:
            return
        b, lib = self.batch, self.lib
        _lib.check(lib.smg_scan_kmer(self.pwmset.handle, self.kmer_w, _ptr(b.packed), _ptr(b.amask), _ptr(b.codes), _ptr(b.d_rows),
                                     b.n_rows, _ptr(self.d_out_rows), _ptr(self.d_chunks), self.n_chunks, self.chunk_len,
                                     self.strands, _ptr(self.d_lo), _ptr(self.d_hi), _ptr(self.d_thr), self.pos_bits, self.motif_bits,
                                     _ptr(self.kmer_head), _ptr(self.kmer_entries), self.kmer_cap, _ptr(self.keys), _ptr(self.scores),
                                     self.hit_cap, _ptr(self.counts), _ptr(self.counters), _stream()), "smg_scan_kmer")

    def _alloc(self):
        dev = self.batch.device
        self.keys = torch.empty(self.hit_cap, dtype=torch.int64, device=dev) if self.want_hits else None
        self.scores = torch.empty(self.hit_cap, dtype=torch.float32, device=dev) if self.want_hits else None
        self.near = torch.empty(self.near_cap * _lib.NEAR_BYTES, dtype=torch.uint8, device=dev)
        self.cand = torch.empty(self.cand_cap, dtype=torch.int64, device=dev) if self.engine in ("hmma", "auto") else None

    @property
    def key_bits(self):
        return self.motif_bits + self.pos_bits + self.row_bits

    def launch(self):
        self.counters.zero_()
        if self.counts is not None:
            self.counts.zero_()
        self.launch_scan()
        if self.engine == "regtab" and self.adjudicate:
            self.launch_adjudicate()

    def launch_scan(self):
        """Enqueue the scan kernels only (the tensor-core engine's exact second pass includes the adjudication)."""
        if self.engine == "auto":  # k-mer index side, register-table side (+ its adjudication), then the tensor-core side
            self.launch_kmer_side()
            self.launch_regtab_side()
        if self.engine in ("hmma", "auto"):
            self.launch_tc_side()
            return
        self._launch_regtab()

    def launch_regtab_side(self):
        """Register-table engine on its side of the width split, with its fp64 adjudication (engine 'auto')."""
        if self.pwmset.n_groups == 0 and self.pwmset.n_wide == 0:
            return  # every narrow PWM is served by the k-mer index
        self._launch_regtab()
        if self.adjudicate:
            self.launch_adjudicate()

    def launch_tc_side(self):
        """Tensor-core prefilter + exact second pass on its side of the width split."""
        b, lib = self.batch, self.lib
        if self.tc_kind == "seed":
            _lib.check(lib.smg_scan_seed(self.pwmset.handle, self.seed_len, _ptr(b.packed), _ptr(b.amask), _ptr(b.codes), _ptr(b.d_rows),
                                         b.n_rows, _ptr(self.d_out_rows), _ptr(self.d_chunks), self.n_chunks, self.chunk_len,
                                         self.strands, _ptr(self.d_lo), _ptr(self.d_hi), _ptr(self.d_thr), self.pos_bits,
                                         self.motif_bits, self.cand_pos_bits, _ptr(self.seed_head), _ptr(self.seed_entries),
                                         _ptr(self.cand), self.cand_cap, _ptr(self.keys), _ptr(self.scores), self.hit_cap,
                                         _ptr(self.counts), _ptr(self.counters), _stream()), "smg_scan_seed")
            return
        fn = lib.smg_scan_tc5 if self.tc_kind == "tc5" else lib.smg_scan_tc
        _lib.check(fn(self.pwmset.handle, _ptr(b.packed), _ptr(b.amask), _ptr(b.codes), _ptr(b.d_rows), b.n_rows,
                                   _ptr(self.d_out_rows), _ptr(self.d_chunks), self.n_chunks, self.chunk_len, self.strands,
                                   _ptr(self.d_lo), _ptr(self.d_hi), _ptr(self.d_pre), _ptr(self.d_thr), self.pos_bits,
                                   self.motif_bits, self.cand_pos_bits, _ptr(self.cand), self.cand_cap, _ptr(self.keys),
                                   _ptr(self.scores), self.hit_cap, _ptr(self.counts), _ptr(self.counters), _stream()),
                   "smg_scan_tc5" if self.tc_kind == "tc5" else "smg_scan_tc")

    def _launch_regtab(self):
        b, lib = self.batch, self.lib
        _lib.check(lib.smg_scan(self.pwmset.handle, _ptr(b.packed), _ptr(b.amask), _ptr(b.codes), _ptr(b.d_rows), b.n_rows,
                                _ptr(self.d_out_rows), _ptr(self.d_chunks), self.n_chunks, self.chunk_len, self.strands,
                                _ptr(self.d_lo), _ptr(self.d_hi), self.pos_bits, self.motif_bits, _ptr(self.keys),
                                _ptr(self.scores), self.hit_cap, _ptr(self.near), self.near_cap, _ptr(self.counts),
                                _ptr(self.counters), _stream()), "smg_scan")

    def launch_adjudicate(self):
        b, lib = self.batch, self.lib
        _lib.check(lib.smg_adjudicate(self.pwmset.handle, _ptr(b.packed), _ptr(b.amask), _ptr(b.codes), _ptr(b.d_rows),
                                      _ptr(self.d_out_rows), _ptr(self.d_thr), _ptr(self.near), self.near_cap, self.pos_bits,
                                      self.motif_bits, _ptr(self.keys), _ptr(self.scores), self.hit_cap, _ptr(self.counts),
                                      _ptr(self.counters), _stream()), "smg_adjudicate")

    def finish(self, sort=True):
        while True:
            c = self.counters.cpu().numpy()
            n_app, n_near = int(c[_lib.CTR_HITS_APPENDED]), int(c[_lib.CTR_NEAR])  # CTR_NEAR: entries of the `near` buffer only
            grow = False
            if n_near > self.near_cap:
                self.near_cap = int(n_near * 1.25) + 1024
                grow = True
            if self.want_hits and n_app > self.hit_cap:
                self.hit_cap = int(n_app * 1.1) + 1024
                grow = True
            if self.engine in ("hmma", "auto") and int(c[_lib.CTR_CANDIDATES]) > self.cand_cap:
                self.cand_cap = int(int(c[_lib.CTR_CANDIDATES]) * 1.1) + 1024
                grow = True
            if not grow:
                break
            self._alloc()
            self.launch()
        n_hits = n_app if self.want_hits else 0
        stats = {"n_definite": int(c[_lib.CTR_DEFINITE]), "n_near": n_near + int(c[_lib.CTR_NEAR_TC]),
                 "n_adjudicated_accepted": int(c[_lib.CTR_ADJ_ACCEPTED]),
                 "n_hits_total": int(c[_lib.CTR_DEFINITE]) + int(c[_lib.CTR_ADJ_ACCEPTED]), "chunk_len": self.chunk_len,
                 "n_chunks": self.n_chunks, "hit_cap": self.hit_cap, "near_cap": self.near_cap, "engine": self.engine,
                 "tc_kind": self.tc_kind if self.engine in ("hmma", "auto") else None, "kmer_max_width": self.kmer_w,
                 "kmer_index_entries": getattr(self, "kmer_n_entries", 0),
                 "n_candidates": int(c[_lib.CTR_CANDIDATES])}
        keys, scores = self.keys, self.scores
        if self.want_hits and sort and n_hits > 0:
            keys, scores = sort_hits(keys, scores, n_hits, self.key_bits, motif_bits=self.motif_bits,
                                     total_positions=self.batch.total_own * bin(self.strands).count("1"))
        return ScanResult(keys, scores, n_hits, self.counts, self.pos_bits, self.motif_bits, stats)


def scan(pwmset, batch, thr64, strands, out_rows, n_out_rows, want_hits=True, want_counts=True, adjudicate=True,
         chunk_len=None, hit_cap=None, near_cap=None, sort=True, engine=None):
    """Fused two-strand scan of a SeqBatch against a DevicePWMSet.

    thr64: per-PWM thresholds (frac * max_scores, computed by the caller exactly as models.py:141 does).
    out_rows: int32 (n_rows, 2) output row of (row, strand) in the reference's row order; n_out_rows their count.
    Returns a ScanResult.  Hit-buffer overflow is detected and the scan retried with a larger buffer, never truncated.
    """
    plan = ScanPlan(pwmset, batch, thr64, strands, out_rows, n_out_rows, want_hits, want_counts, adjudicate, chunk_len,
                    hit_cap, near_cap, engine)
    plan.launch()
    return plan.finish(sort=sort)


def sort_hits(keys, scores, n, key_bits, keys_out=None, scores_out=None, temp=None, return_temp=False, motif_bits=None,
              n_out_rows=None, total_positions=None, overflow_flag=None):
    """Sort the first n (key, score) pairs into the reference's np.where order.  temp: reusable scratch (grown when too
    small; pass return_temp=True to get the possibly re-allocated buffer back as a third result).

    With motif_bits (and the scan's geometry) the BUCKETED sort is used (csrc/sort_buckets.cu: one scatter pass into
    (output row, position block) buckets + a shared-memory sort per bucket); if a bucket overflows (extreme local hit
    density) the call falls back to the radix sort -- one small D2H read of the overflow flag decides."""
    lib = _lib.load()
    dev = keys.device
    if keys_out is None:
        keys_out = torch.empty(n, dtype=torch.int64, device=dev)
        scores_out = torch.empty(n, dtype=torch.float32, device=dev)
    if motif_bits is not None and 0 < n < (1 << 32) and os.environ.get("SMEAGOL_B200_SORT", "bucket") == "bucket":
        # positions per bucket: ~300-400 hits on average (a warp orders up to 1,024 in shared memory), at most 2^26 buckets
        per_pos = n / max(float(total_positions or n), 1.0)
        b = int(np.clip(np.floor(np.log2(max(384.0 / max(per_pos, 1e-9), 1.0))), 0, min(9, 31 - motif_bits)))
        shift = motif_bits + b
        if key_bits - shift <= 26 and shift < key_bits:
            tb = ctypes.c_uint64(0)
            _lib.check(lib.smg_sort_hits_bucketed(_ptr(keys), _ptr(scores), _ptr(keys_out), _ptr(scores_out), n, key_bits, motif_bits, shift,
                                                  None, ctypes.byref(tb), None, _stream()), "smg_sort_hits_bucketed(query)")
            if temp is None or temp.numel() < int(tb.value):
                temp = torch.empty(max(int(tb.value), 16), dtype=torch.uint8, device=dev)
            if overflow_flag is None:
                overflow_flag = torch.zeros(1, dtype=torch.int32, device=dev)
            tb = ctypes.c_uint64(temp.numel())
            _lib.check(lib.smg_sort_hits_bucketed(_ptr(keys), _ptr(scores), _ptr(keys_out), _ptr(scores_out), n, key_bits, motif_bits, shift,
                                                  _ptr(temp), ctypes.byref(tb), _ptr(overflow_flag), _stream()),
                       "smg_sort_hits_bucketed")
            if int(overflow_flag.item()) == 0:
                return (keys_out, scores_out, temp) if return_temp else (keys_out, scores_out)
    tb = ctypes.c_uint64(0)
    _lib.check(lib.smg_sort_hits(_ptr(keys), _ptr(scores), _ptr(keys_out), _ptr(scores_out), n, key_bits, None,
                                 ctypes.byref(tb), _stream()), "smg_sort_hits(query)")
    if temp is None or temp.numel() < int(tb.value):
        temp = torch.empty(max(int(tb.value), 16), dtype=torch.uint8, device=dev)
    tb = ctypes.c_uint64(temp.numel())
    _lib.check(lib.smg_sort_hits(_ptr(keys), _ptr(scores), _ptr(keys_out), _ptr(scores_out), n, key_bits, _ptr(temp),
                                 ctypes.byref(tb), _stream()), "smg_sort_hits")
    if return_temp:
        return keys_out, scores_out, temp
    return keys_out, scores_out


def predict_dense(pwmset, batch, row_len):
    lib = _lib.load()
    out = torch.empty((batch.n_rows, row_len, pwmset.n_motifs), dtype=torch.float32, device=batch.device)
    _lib.check(lib.smg_predict_dense(pwmset.handle, _ptr(batch.packed), _ptr(batch.amask), _ptr(batch.codes), _ptr(batch.d_rows),
                                     batch.n_rows, row_len, _ptr(out), _stream()), "smg_predict_dense")
    return out
