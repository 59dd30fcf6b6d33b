/* smeagol_b200 -- C-ABI of the B200-native PWM scanning hot path.
 *
 * This is the drop-in boundary for SMEAGOL's scan path.  The reference (gruber-sciencelab/SMEAGOL v0.1.1) is pure
 * Python and has no FFI layer; its seam is PWMModel.predict_with_threshold (smeagol/models.py:119-178), fed by
 * encode.integer_encode (smeagol/encode.py:47-68) and consumed by scan._find_sites_seq (smeagol/scan.py:150-208).
 * Each entry point below cites the reference interface it replaces.  INTEGRATION.md shows the ctypes binding a
 * maintainer of the reference would add.
 *
 * Conventions
 *  - plain C types only; every `d_*` pointer is DEVICE memory owned by the caller (the library never frees it);
 *    the PWM-set handle owns its own device tables.
 *  - `stream` is a cudaStream_t passed as void*; all kernels are enqueued asynchronously on it.
 *  - return value: SMG_OK or an SMG_ERR_* code; smg_last_error() gives a message for the calling thread.
 *  - the library never falls back to the CPU.
 */
#ifndef SMEAGOL_B200_H
#define SMEAGOL_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SMG_OK 0
#define SMG_ERR_INVALID 1      /* bad argument */
#define SMG_ERR_CUDA 2         /* a CUDA runtime call failed */
#define SMG_ERR_UNSUPPORTED 3  /* e.g. no sm_100 device */

#define SMG_MAX_REGTAB_WIDTH 32 /* wider PWMs go through scan_wide_kernel (one thread per window start, tables from global memory) */
#define SMG_STRAND_FWD 1
#define SMG_STRAND_RC 2

typedef struct smg_pwmset* smg_pwmset_t;

/* One scanned row: a whole sequence, or a segment of a longer one (chunked / multi-GPU scans).  Bases are packed
 * 16 per 32-bit word starting at word_off; rows are padded with >= 96 'Z' bases.  A window start s (row-local)
 * is reported by this row iff 0 <= s < n_own and g_off + s + W <= g_len (models.py:110 trim rule). */
typedef struct {
    int64_t word_off; /* first word of the row in d_packed / d_amask; base offset in d_codes is 16*word_off */
    int64_t g_off;    /* coordinate of the row's first base in the full sequence (0 for whole sequences) */
    int64_t g_len;    /* length of the full sequence */
    int32_t n_avail;  /* bases present in the buffer (owned + right halo) */
    int32_t n_own;    /* window starts owned by this row */
} smg_row_t;

typedef struct {
    int32_t row;   /* index into the row array */
    int32_t start; /* first owned window start of the chunk (multiple of 16) */
} smg_chunk_t;

/* near-threshold candidate awaiting fp64 adjudication (north_star: sites within 1e-5 of the threshold) */
typedef struct {
    int32_t row;
    int32_t motif;
    int32_t strand; /* 0 = forward, 1 = reverse complement */
    int32_t start;  /* row-local window start on the forward string */
    float score;    /* fp32 kernel score */
    int32_t reserved;
} smg_near_t;

/* counters written by smg_scan / smg_adjudicate (uint64 each) */
#define SMG_CTR_HITS_APPENDED 0 /* entries reserved in the hit buffer (> hit_cap means overflow: grow and rescan) */
#define SMG_CTR_NEAR 1          /* near-threshold candidates (> near_cap means overflow) */
#define SMG_CTR_DEFINITE 2      /* hits decided by the fp32 kernel */
#define SMG_CTR_ADJ_ACCEPTED 3  /* near-threshold candidates accepted by the fp64 adjudicator */
#define SMG_CTR_LAUNCHES 4      /* kernel launches issued by the library on behalf of this counter block */
#define SMG_CTR_CANDIDATES 5    /* smg_scan_tc: candidates emitted by the tensor-core prefilter (> cand_cap: overflow) */
#define SMG_CTR_NEAR_TC 6       /* smg_scan_tc: near-threshold candidates adjudicated inside its exact pass (a tally: they
                                   never occupy d_near, so they do not count against near_cap) */
#define SMG_CTR_KMER_ENTRIES 7   /* smg_kmer_index_build / smg_seed_index_build: entry slots reserved, in blocks of 64 per warp (> entry_cap: overflow, rebuild with a larger buffer) */
#define SMG_N_COUNTERS 8

/* one (k-mer, PWM, strand) hit of the k-mer index: col = motif * 2 + strand | flags (bit 30: decided by the fp64
 * adjudicator, bit 29: rejected by it), score = the fp32 window score, next = next entry of the same k-mer or -1 */
typedef struct {
    uint32_t col;
    float score;
    int32_t next;
    int32_t reserved;
} smg_kmer_entry_t;
#define SMG_KMER_MAX_WIDTH 12

int smg_version(void);
const char* smg_last_error(void);
/* number of exported kernels launched by this process so far (bench.py's gpu_launches claim) */
uint64_t smg_launch_count(void);

/* PWMModel.__init__ / _pad_weights / set_model_weights (models.py:35-48, 70-90): weights are the fp32-cast PWMs
 * concatenated row-major (sum(widths) x 4, columns A,C,G,T), HOST memory.  Builds the device tables: natural
 * layout, reverse-complement tables, and the lane-major register-table groups of the fused scan kernel.
 * Non-finite weights are stored as given; to reproduce the reference (0 * -inf = NaN in its one-hot contraction, so
 * such a PWM never yields a site) pass +inf thresholds for those PWMs to the scan calls. */
int smg_pwmset_create(const float* h_weights, const int32_t* h_widths, int32_t n_motifs, smg_pwmset_t* out);
/* Hybrid engine: PWMs narrower than tc_min_width (0 = no split, else a width in 1..33) are scanned by smg_scan (register-table
 * kernel), the others by smg_scan_tc (tensor-core prefilter + exact second pass); a complete scan calls both on the
 * same hit / count / counter buffers.  With tc_min_width = 0 either entry point alone covers every PWM. */
int smg_pwmset_create_split(const float* h_weights, const int32_t* h_widths, int32_t n_motifs, int32_t tc_min_width,
                            smg_pwmset_t* out);
/* Three-engine split: PWMs narrower than regtab_min_width belong to the k-mer index engine alone (smg_scan_kmer with
 * kmer_max_width = regtab_min_width - 1), those in [regtab_min_width, tc_min_width) and those wider than
 * SMG_MAX_REGTAB_WIDTH to smg_scan, those in [tc_min_width, SMG_MAX_REGTAB_WIDTH] to smg_scan_tc / smg_scan_tc5.
 * regtab_min_width = 0 or 1: no k-mer engine (same as smg_pwmset_create_split). */
int smg_pwmset_create_split3(const float* h_weights, const int32_t* h_widths, int32_t n_motifs, int32_t regtab_min_width,
                             int32_t tc_min_width, smg_pwmset_t* out);
int smg_pwmset_destroy(smg_pwmset_t set);
/* info[0]=n_motifs info[1]=max_width info[2]=n_regtab_groups info[3]=n_buckets info[4]=n_wide_motifs
 * info[5]=sum over groups of padded width (dual groups count both strands) info[6]=128-column blocks and
 * info[7]=column groups of the tcgen05 prefilter */
int smg_pwmset_info(smg_pwmset_t set, int64_t info[8]);

/* encode.integer_encode (encode.py:47-68) on device.  d_src: concatenated ASCII bytes (input_kind 0) or integer
 * codes 0..16 as uint8 (input_kind 1); row r occupies d_src[src_off[r] .. src_off[r]+rows[r].n_avail).
 * Outputs: 2-bit packed bases, a 1-bit ambiguity mask (one uint16 per packed word), optionally the uint8 codes
 * (d_codes may be NULL when the caller knows there is no ambiguity code).  d_status[0] = index into d_src of the
 * first invalid character or -1 (the reference raises KeyError, encode.py:66); d_status[1] = 1 if any base is
 * not A/C/G/T/U.  d_status must be initialised to {-1, 0}. */
int smg_pack(const uint8_t* d_src, int input_kind, const smg_row_t* d_rows, const int64_t* d_src_off,
             int32_t n_rows, int64_t total_words, uint32_t* d_packed, uint16_t* d_amask, uint8_t* d_codes,
             int64_t* d_status, void* stream);

/* scan._bin_sites_by_score (scan.py:77-111) on the device ("next" row 8(f)-3): histogram of frac_score =
 * score / max_score (fp64) of a hit list (sorted or not) over right-closed bins (d_edges[b], d_edges[b+1]], b < n_bins
 * (pd.cut semantics; d_edges has n_bins + 1 ascending entries).  d_hist[n_out_rows][n_motifs][n_bins] must be zeroed by
 * the caller; hits outside (d_edges[0], d_edges[n_bins]] are not counted. */
int smg_bin_hits(const uint64_t* d_keys, const float* d_scores, uint64_t n_hits, int pos_bits, int motif_bits,
                 const double* d_max_scores, const double* d_edges, int32_t n_bins, int32_t n_motifs, int32_t* d_hist, void* stream);

/* matrices.pairwise_ncorrs (matrices.py:396-433) on the device ("next" row 8(f)-4): all pairwise normalised Pearson
 * correlations (max over the alignments of align_pms and over the four orientation combinations, exactly as ncorr and
 * the reference's reverse_complement compute them) of n position matrices.  d_mats: the matrices concatenated row-major
 * (sum of widths x 4 doubles), d_offs[m] = first row of matrix m; d_out: n x n doubles (symmetric, diagonal 1). */
int smg_pairwise_ncorrs(const double* d_mats, const int32_t* d_widths, const int64_t* d_offs, int32_t n, double* d_out, void* stream);

/* scan.count_in_sliding_windows (scan.py:381-423) on the device ("next" row 8(f)-4): for every hit whose PWM is selected
 * (d_motif_sel[motif] != 0) one count in each tiling window [k*shift, k*shift + width) of its record that contains the hit's
 * start.  d_row_rec[out_row] = record index of an output row (-1: ignore); record r owns the windows d_win_off[r] ..
 * d_win_off[r+1] of d_counts (zeroed by the caller).  The hit list may be unsorted. */
int smg_window_counts(const uint64_t* d_keys, uint64_t n_hits, int pos_bits, int motif_bits, const uint8_t* d_motif_sel,
                      const int32_t* d_row_rec, const int64_t* d_win_off, int64_t width, int64_t shift, int32_t* d_counts, void* stream);

/* utils.shuffle_records (utils.py:36-83) on the device ("next" row 8(f)-1): n_copies k-mer preserving shuffles of one
 * sequence, all in one launch.  d_tok: the sequence as integer codes 0..16 (len bytes).  k = 2 (dinucleotide, the
 * published algorithm of deeplift.dinuc_shuffle): d_succ = the successors tok[i+1] bucketed by tok[i] in original order
 * (len - 1 bytes), d_off[18] = bucket offsets.  k = 1: uniform permutation (d_succ, d_off unused).  Output: ASCII, copy c
 * at d_out + c * stride.  d_scratch (n_copies * len bytes) is only needed when len > 200 KiB (shared memory otherwise).
 * Seeded and deterministic, but not numpy's / Python's random stream. */
int smg_kmer_shuffle(const uint8_t* d_tok, int64_t len, const uint8_t* d_succ, const int64_t* d_off, int32_t n_copies, int32_t k,
                     uint64_t seed, uint8_t* d_scratch, uint8_t* d_out, int64_t stride, void* stream);

/* io.read_fasta (io.py:16-41, Biopython SeqIO.parse(handle, "fasta")) for the part that touches every byte, on the
 * device ("next" row 8(f)-2: FASTA text -> per-record ASCII, which smg_pack then encodes).  d_bytes: the file's bytes
 * (after gunzip), 16-byte aligned, n < 2^31.
 * smg_fasta_index: offsets of the header lines ('>' as the first byte of a line), UNORDERED, into d_hdr_start;
 *   *d_n_hdr = how many were found (> hdr_cap: call again with a larger buffer).  The caller sorts them.
 * smg_fasta_compact: with the SORTED header offsets: d_hdr_end[r] = offset of the '\n' ending header r (or n);
 *   d_seq_len[r] = bases of record r; d_out = the sequences of all records back to back in file order, with
 *   '\n', '\r', ' ' and '\t' removed (record r starts at the exclusive prefix sum of d_seq_len); *d_n_out = their
 *   total.  Bytes before the first header are ignored, as Biopython does.  upper != 0 maps a-z to A-Z on the fly
 *   (the reference raises KeyError on lower case; soft-masked genomes need this).  d_bytes is MODIFIED (header lines
 *   are blanked).  Temporary storage as in CUB: call with d_temp = NULL to get *temp_bytes. */
int smg_fasta_index(const uint8_t* d_bytes, int64_t n, int64_t* d_hdr_start, uint64_t hdr_cap, uint64_t* d_n_hdr, void* stream);
int smg_fasta_compact(uint8_t* d_bytes, int64_t n, const int64_t* d_hdr_start, int64_t n_hdr, int64_t* d_hdr_end,
                      uint64_t* d_seq_len, uint8_t* d_out, int64_t* d_n_out, int upper, void* d_temp, uint64_t* temp_bytes,
                      void* stream);

/* PWMModel.predict_with_threshold (models.py:102-178) fused with the per-(row, motif, strand) counting of
 * scan._count_sites (scan.py:114-147): both strands, threshold, end trim, stream compaction, counts.
 *  d_thr_lo / d_thr_hi [n_motifs]: fp32 bounds of the near-threshold band; score > thr_hi is a definite hit,
 *      thr_lo < score <= thr_hi is appended to d_near for smg_adjudicate.  With thr_lo == thr_hi == the largest
 *      float <= frac*max_score the kernel reproduces the reference's `(double)score > thr` compare exactly.
 *  d_out_rows [n_rows][2]: output row index of (row, strand) in the reference's row order (encode.py:137-160),
 *      used as the most significant key field; strands bit mask selects the strands scanned.
 *  hit key = out_row << (pos_bits+motif_bits) | pos << motif_bits | motif, pos in the coordinates of the scanned
 *      string (for the RC strand: g_len - start - W, as the reference reports them).
 *  d_hit_keys/d_hit_scores may be NULL (count-only); d_counts [n_rows][n_motifs][2] int32 may be NULL.
 *  d_counters: SMG_N_COUNTERS uint64, zeroed by the caller. */
int smg_scan(smg_pwmset_t set, const uint32_t* d_packed, const uint16_t* d_amask, const uint8_t* d_codes,
             const smg_row_t* d_rows, int32_t n_rows, const int32_t* d_out_rows, const smg_chunk_t* d_chunks,
             int32_t n_chunks, int32_t chunk_len, int strands, const float* d_thr_lo, const float* d_thr_hi,
             int pos_bits, int motif_bits, uint64_t* d_hit_keys, float* d_hit_scores, uint64_t hit_cap,
             smg_near_t* d_near, uint64_t near_cap, int32_t* d_counts, uint64_t* d_counters, void* stream);

/* Same contract and results as smg_scan + smg_adjudicate (hits, counts and counters are bit-identical), computed in two
 * passes: (1) a tensor-core prefilter (mma.sync, fp16 weights x soft one-hot generated in registers, threshold on the
 * accumulator fragment) emits every (row, start, motif, strand) whose approximate score exceeds d_thr_pre[motif]
 * (the caller passes thr_lo minus a margin >= the fp16 error bound) into d_cand; (2) every candidate is re-scored in
 * fp32 in the fixed summation order, the near-threshold band is adjudicated in fp64 against d_thr64, hits and counts
 * are written.  candidate = row << (cand_pos_bits+motif_bits+1) | start << (motif_bits+1) | motif << 1 | strand.
 * PWMs wider than SMG_MAX_REGTAB_WIDTH are not supported here (use smg_scan). */
int smg_scan_tc(smg_pwmset_t set, const uint32_t* d_packed, const uint16_t* d_amask, const uint8_t* d_codes,
                const smg_row_t* d_rows, int32_t n_rows, const int32_t* d_out_rows, const smg_chunk_t* d_chunks,
                int32_t n_chunks, int32_t chunk_len, int strands, const float* d_thr_lo, const float* d_thr_hi,
                const float* d_thr_pre, const double* d_thr64, int pos_bits, int motif_bits, int cand_pos_bits,
                uint64_t* d_cand, uint64_t cand_cap, uint64_t* d_hit_keys, float* d_hit_scores, uint64_t hit_cap,
                int32_t* d_counts, uint64_t* d_counters, void* stream);

/* Same contract, arguments and results as smg_scan_tc; the prefilter runs on the 5th-generation tensor cores instead:
 * tcgen05.mma (kind::f16, fp32 accumulators in TMEM) with the fp16 tables of 128 (PWM, strand) columns resident in shared
 * memory as the A operand and the soft one-hot of the sequence as a Toeplitz B operand (one 8-byte entry per base, the
 * shared-memory descriptor overlaps its rows, no im2col matrix), 16 epilogue warps reading the accumulators back with
 * tcgen05.ld and thresholding them (smeagol_b200/csrc/scan_tc5.cuh).  chunk_len should be a multiple of 512. */
int smg_scan_tc5(smg_pwmset_t set, const uint32_t* d_packed, const uint16_t* d_amask, const uint8_t* d_codes,
                 const smg_row_t* d_rows, int32_t n_rows, const int32_t* d_out_rows, const smg_chunk_t* d_chunks,
                 int32_t n_chunks, int32_t chunk_len, int strands, const float* d_thr_lo, const float* d_thr_hi,
                 const float* d_thr_pre, const double* d_thr64, int pos_bits, int motif_bits, int cand_pos_bits,
                 uint64_t* d_cand, uint64_t cand_cap, uint64_t* d_hit_keys, float* d_hit_scores, uint64_t hit_cap,
                 int32_t* d_counts, uint64_t* d_counters, void* stream);

/* K-mer index engine for PWMs of width <= kmer_max_width <= SMG_KMER_MAX_WIDTH (smeagol_b200/csrc/kmer.cu): the same
 * contract and results as smg_scan + smg_adjudicate for those PWMs (hits, fp32 scores, counts, counters), computed by
 * scoring every k-mer of every width class once per call (smg_kmer_index_build: fp32 in the fixed order, near-threshold
 * band adjudicated in fp64) and then looking window contents up (smg_scan_kmer); windows holding IUPAC codes are scored
 * directly.  The index depends on the thresholds and the strand mask; it is valid for any sequences.
 *  smg_kmer_index_size: *head_len = int32 slots of d_head for this PWM set and kmer_max_width.
 *  smg_kmer_index_build: d_head (head_len int32) and d_entries (entry_cap) are written; d_counters[SMG_CTR_KMER_ENTRIES]
 *      (zeroed by the caller) = entries needed (> entry_cap: rebuild with a larger buffer). */
int smg_kmer_index_size(smg_pwmset_t set, int kmer_max_width, int64_t* head_len);
int smg_kmer_index_build(smg_pwmset_t set, int kmer_max_width, int strands, const float* d_thr_lo, const float* d_thr_hi,
                         const double* d_thr64, int32_t* d_head, int64_t head_len, smg_kmer_entry_t* d_entries,
                         uint64_t entry_cap, uint64_t* d_counters, void* stream);
int smg_scan_kmer(smg_pwmset_t set, int kmer_max_width, const uint32_t* d_packed, const uint16_t* d_amask, const uint8_t* d_codes,
                  const smg_row_t* d_rows, int32_t n_rows, const int32_t* d_out_rows, const smg_chunk_t* d_chunks,
                  int32_t n_chunks, int32_t chunk_len, int strands, const float* d_thr_lo, const float* d_thr_hi,
                  const double* d_thr64, int pos_bits, int motif_bits, const int32_t* d_head, const smg_kmer_entry_t* d_entries,
                  uint64_t entry_cap, uint64_t* d_hit_keys, float* d_hit_scores, uint64_t hit_cap, int32_t* d_counts,
                  uint64_t* d_counters, void* stream);

/* fp64 adjudication of the near-threshold candidates: score64 = sum of the fp32 operands in double, hit iff
 * score64 > d_thr64[motif] (thr = frac * max_scores computed on the host as models.py:141 does). */
int smg_adjudicate(smg_pwmset_t set, const uint32_t* d_packed, const uint16_t* d_amask, const uint8_t* d_codes,
                   const smg_row_t* d_rows, const int32_t* d_out_rows, const double* d_thr64,
                   const smg_near_t* d_near, uint64_t near_cap, int pos_bits, int motif_bits,
                   uint64_t* d_hit_keys, float* d_hit_scores, uint64_t hit_cap, int32_t* d_counts,
                   uint64_t* d_counters, void* stream);

/* np.where row-major (row, pos, motif) order (models.py:108): radix sort of the hit keys with their scores.
 * Call with d_temp == NULL to query *temp_bytes. */
int smg_sort_hits(const uint64_t* d_keys_in, const float* d_scores_in, uint64_t* d_keys_out, float* d_scores_out,
                  uint64_t n, int key_bits, void* d_temp, uint64_t* temp_bytes, void* stream);

/* The same order, bucketed (smeagol_b200/csrc/sort_buckets.cu): one scatter pass moves every hit into its bucket
 * key >> bucket_shift (an output row x a block of positions; bucket_shift = motif_bits + log2(positions per bucket) <= 31,
 * at most 2^9 positions per bucket and 2^28 buckets), then every bucket is ordered in shared memory.  48 B of DRAM traffic per hit instead of the radix
 * sort's 128.  *d_overflow (device uint32): reserved, always written as 0 (a bucket may hold any number of hits).  n < 2^32.
 * Call with d_temp == NULL to query *temp_bytes. */
int smg_sort_hits_bucketed(const uint64_t* d_keys_in, const float* d_scores_in, uint64_t* d_keys_out, float* d_scores_out,
                           uint64_t n, int key_bits, int motif_bits, int bucket_shift, void* d_temp, uint64_t* temp_bytes,
                           uint32_t* d_overflow, void* stream);

/* PWMModel.predict (models.py:92-100): dense (n_rows, L, n_motifs) fp32 scores of equal-length rows, including
 * the zero-padded overhanging windows.  Small inputs / tests only. */
int smg_predict_dense(smg_pwmset_t set, const uint32_t* d_packed, const uint16_t* d_amask, const uint8_t* d_codes,
                      const smg_row_t* d_rows, int32_t n_rows, int32_t row_len, float* d_out, void* stream);

/* variant._variant_effect_on_site (variant.py:7-47) for explicit (site, variant) pairs, fp64:
 * out_ref[i] = sum_k T[k, ref_k], out_alt[i] = same with the base at rel_pos replaced by alt_code. */
int smg_variant_sites(smg_pwmset_t set, const uint32_t* d_packed, const uint16_t* d_amask, const uint8_t* d_codes,
                      const smg_row_t* d_rows, const int32_t* d_site_row, const int32_t* d_site_start,
                      const int32_t* d_site_motif, const int32_t* d_rel_pos, const uint8_t* d_alt_code, int64_t n,
                      double* d_out_ref, double* d_out_alt, void* stream);

/* BASELINE config 5 (batched restatement of variant.py:7-47 + scan.py:8-37): for every SNV (row, pos, alt) and every
 * PWM, the maximum reference and alternate score over all windows covering pos on both strands (fp32).  Outputs
 * [n_snv][n_motifs].  Every score is the fixed ascending-k fp32 sum (bit-identical to the CPU restatement). */
int smg_variant_windows(smg_pwmset_t set, const uint32_t* d_packed, const uint16_t* d_amask, const uint8_t* d_codes,
                        const smg_row_t* d_rows, const int32_t* d_snv_row, const int32_t* d_snv_pos,
                        const uint8_t* d_snv_alt, int64_t n_snv, float* d_ref_max, float* d_alt_max, void* stream);

/* Same, with the choice of how alternate scores are formed and of the output layout.  motif_major = 0: outputs
 * [n_snv][n_motifs] as above; motif_major = 1: outputs [n_motifs][n_snv] -- the layout the kernels write as whole
 * 128-byte lines (the SNV-major layout costs ~8x the DRAM traffic: partial sectors that the L2 reads back).
 * exact = 1: alternate scores as smg_variant_windows.  exact = 0: the
 * alternate score of a window is (reference score - T[k][ref base]) + T[k][alt base] (k = offset of the SNV in the
 * window) -- the reference maxima are unchanged, alternate scores may differ from the fresh sum by a few ulp of the
 * largest partial sum (<= 1e-5 * sum_k max_b |T[k][b]|); falls back to exact = 1 when a table holds a non-finite entry. */
int smg_variant_windows_mode(smg_pwmset_t set, const uint32_t* d_packed, const uint16_t* d_amask, const uint8_t* d_codes,
                             const smg_row_t* d_rows, const int32_t* d_snv_row, const int32_t* d_snv_pos,
                             const uint8_t* d_snv_alt, int64_t n_snv, float* d_ref_max, float* d_alt_max, int32_t exact,
                             int32_t motif_major, void* stream);

/* SEED ENGINE for the PWMs of the tensor-core side (same contract as smg_scan_tc5: exact second pass, identical results).
 * A window of a PWM of width W can only be a hit if its most selective seed_len consecutive positions score more than
 * threshold - (maximum of the remaining positions); every (PWM, strand) becomes a forward-only virtual PWM of width
 * seed_len in a one-class k-mer index built per call, and the scan turns every list entry into a candidate window that
 * finalize re-scores.  smg_seed_prepare: info[0] = virtual PWMs, info[1] = int32 slots of the index buffer (d_head),
 * info[2] = largest seed offset, info[3] = 1 if every PWM of the tensor-core side is seedable (W <= seed_len + 14). */
int smg_seed_prepare(smg_pwmset_t set, int seed_len, int64_t info[4]);
int smg_seed_index_build(smg_pwmset_t set, int seed_len, int strands, const float* d_thr_lo, int32_t* d_head, int64_t head_len,
                         smg_kmer_entry_t* d_entries, uint64_t entry_cap, uint64_t* d_counters, void* stream);
int smg_scan_seed(smg_pwmset_t set, int seed_len, const uint32_t* d_packed, const uint16_t* d_amask, const uint8_t* d_codes,
                  const smg_row_t* d_rows, int32_t n_rows, const int32_t* d_out_rows, const smg_chunk_t* d_chunks,
                  int32_t n_chunks, int32_t chunk_len, int strands, const float* d_thr_lo, const float* d_thr_hi,
                  const double* d_thr64, int pos_bits, int motif_bits, int cand_pos_bits, const int32_t* d_head,
                  const smg_kmer_entry_t* d_entries, uint64_t* d_cand, uint64_t cand_cap, uint64_t* d_hit_keys,
                  float* d_hit_scores, uint64_t hit_cap, int32_t* d_counts, uint64_t* d_counters, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* SMEAGOL_B200_H */
