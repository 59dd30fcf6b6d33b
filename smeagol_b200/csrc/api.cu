// C-ABI entry points of libsmeagol_b200.so (see include/smeagol_b200.h) and the small non-hot kernels:
// 2-bit packer, dense predict, fp64 adjudicator, wide-motif fallback, variant scoring, hit sort.
#include <algorithm>
#include <atomic>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>
#include <string>
#include <vector>

#include <cub/device/device_radix_sort.cuh>

#include "common.cuh"
#include "launchers.h"
#include "variant_regtab.cuh"
#include "scan_hmma.cuh"
#include "hmma_launchers.h"
#include <cuda_fp16.h>
#include "variant_launchers.h"
#include "scan_tc5_params.cuh"
#include "kmer.h"
#include "tc5_launchers.h"

// ------------------------------------------------------------------------------------------ host-side state
static thread_local std::string g_err;
static std::atomic<uint64_t> g_launches{0};

static int fail(int code, const std::string& msg)
{
    g_err = msg;
    return code;
}
#define SMG_CUDA(call)                                                                         \
    do {                                                                                       \
        cudaError_t e_ = (call);                                                               \
        if (e_ != cudaSuccess)                                                                 \
            return fail(SMG_ERR_CUDA, std::string(#call) + ": " + cudaGetErrorString(e_));     \
    } while (0)
#define SMG_LAUNCH_CHECK(name)                                                                 \
    do {                                                                                       \
        ++g_launches;                                                                          \
        cudaError_t e_ = cudaGetLastError();                                                   \
        if (e_ != cudaSuccess) return fail(SMG_ERR_CUDA, std::string(name) + ": " + cudaGetErrorString(e_)); \
    } while (0)

struct Group {
    int wb;  // padded width
    int nm;  // PWMs per lane (1 or 2)
    int tc;  // 1: every PWM of the group is on the tensor-core side of the split (decided by TRUE widths, not wb)
    int lane_motif[SMG_MAX_PWMS_PER_LANE][32];
    int64_t tab_off;
};
struct Bucket {
    int wb;
    int nm;
    int tc;
    int first_group;
    int n_groups;
};
struct smg_pwmset {
    bool all_finite = true;  // no +-inf / NaN table entry (the delta form of the SNV kernel needs that)
    int n_motifs = 0;
    int max_w = 0;
    std::vector<int> widths;
    std::vector<int64_t> offs;
    std::vector<Group> groups;
    std::vector<Bucket> buckets;
    std::vector<int> wide;
    std::vector<int> dual_motifs, nondual_motifs;
    // tensor-core prefilter: (PWM, strand) columns grouped by k16 steps KS = ceil(W / 4)
    struct HmmaBucket { int ks, nb, first_colgroup, n_colgroups; };
    std::vector<HmmaBucket> hbuckets;
    int64_t n_hcols = 0;
    int split = 0;  // > 0: smg_scan covers PWMs narrower than this, smg_scan_tc the others (hybrid engine)
    int kmin = 0;   // > 1: PWMs narrower than this belong to the k-mer index engine alone (in no lane group / fragment table)
    int32_t* d_cls_motifs = nullptr;  // motif ids sorted by width (k-mer index classes)
    uint32_t* d_afrag = nullptr;
    int64_t* d_cg_afrag_off = nullptr;
    int32_t* d_col_id = nullptr;
    float* d_thr_pre_col = nullptr;
    int32_t* d_dual_list = nullptr;
    int32_t* d_nondual_list = nullptr;
    // tcgen05 prefilter (scan_tc5.cuh): 128-column blocks of (PWM, strand) columns grouped into column groups whose fp16
    // images fit in one CTA's shared memory
    int t5_n_blocks = 0, t5_n_groups = 0, t5_max_image = 0;
    uint8_t* d_t5_img = nullptr;
    int64_t* d_t5_cg_img_off = nullptr;
    int32_t* d_t5_cg_first_blk = nullptr;
    int32_t* d_t5_blk_ks = nullptr;
    int32_t* d_t5_blk_off = nullptr;
    int32_t* d_t5_col_id = nullptr;
    float* d_t5_thr_pre_col = nullptr;
    int64_t sum_padded = 0;
    // seed engine (kmer.h): per seed length S, the virtual forward-only PWMs of the tensor-core side's (PWM, strand) columns
    struct SeedSet {
        int S = 0, nv = 0, o_max = 0;
        bool eligible = false;  // every PWM of the tensor-core side satisfies W <= S + kSeedMaxRest
        std::vector<int> widths;  // nv x S (layout of the one-class index)
        float* d_vtab = nullptr; int64_t* d_voff = nullptr; int32_t* d_vwidth = nullptr; int32_t* d_vcol = nullptr;
        int32_t* d_voffset = nullptr; float* d_vrest = nullptr; float* d_vthr = nullptr; double* d_vthr64 = nullptr;
        int32_t* d_cls = nullptr;
    };
    std::map<int, SeedSet*> seeds;
    std::vector<float> h_nat;  // host copy of the natural tables (seed selection)
    float* d_nat = nullptr;
    int32_t* d_width = nullptr;
    int64_t* d_off = nullptr;
    float* d_gtab = nullptr;
    int64_t* d_gtab_off = nullptr;
    int32_t* d_lane_motif = nullptr;
    int32_t* d_wide = nullptr;
};

#ifndef SMG_DUAL_MAX_WIDTH
#define SMG_DUAL_MAX_WIDTH 32
#endif
// PWMs up to this width keep both strands in one lane (float2 pairs, half-swapped table); wider ones use one lane
// per (PWM, strand) column
static const int kDualMaxWidth = SMG_DUAL_MAX_WIDTH;
#ifndef SMG_PAIR_MAX_WIDTH
#define SMG_PAIR_MAX_WIDTH 0  // measured: two PWMs per lane lose more occupancy than they save in dispatch (off)
#endif
static const int kPairMaxWidth = SMG_PAIR_MAX_WIDTH;  // lanes hold two PWMs up to this (padded) width

// side streams used to run the per-width-bucket launches of one smg_scan call concurrently (per device)
struct StreamPool {
    static const int kSide = 7;
    cudaStream_t side[kSide];
    cudaEvent_t fork;
    cudaEvent_t join[kSide];
    bool ok = false;
};
static StreamPool g_pools[64];
static StreamPool* get_pool()
{
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev >= 64) return nullptr;
    StreamPool& sp = g_pools[dev];
    if (!sp.ok) {
        for (int i = 0; i < StreamPool::kSide; ++i) {
            if (cudaStreamCreateWithFlags(&sp.side[i], cudaStreamNonBlocking) != cudaSuccess) return nullptr;
            if (cudaEventCreateWithFlags(&sp.join[i], cudaEventDisableTiming) != cudaSuccess) return nullptr;
        }
        if (cudaEventCreateWithFlags(&sp.fork, cudaEventDisableTiming) != cudaSuccess) return nullptr;
        sp.ok = true;
    }
    return &sp;
}  // PWMs up to this width keep both strands in one lane (float2 pairs)

// ------------------------------------------------------------------------------------------ kernels
static __constant__ uint8_t c_lut[256];

__global__ void __launch_bounds__(256) pack_kernel(const uint8_t* __restrict__ src, const int kind,
                                                   const smg_row_t* __restrict__ rows, const int64_t* __restrict__ src_off,
                                                   const int n_rows, const int64_t total_words, uint32_t* __restrict__ packed,
                                                   uint16_t* __restrict__ amask, uint8_t* __restrict__ codes,
                                                   unsigned long long* __restrict__ status)
{
    // the 256-entry LUT is read with per-lane (divergent) indices: shared memory, not the constant cache
    __shared__ uint8_t s_lut[256];
    s_lut[threadIdx.x] = c_lut[threadIdx.x];
    __syncthreads();
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t wi = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; wi < total_words; wi += stride) {
        int lo = 0, hi = n_rows - 1;  // largest r with rows[r].word_off <= wi
        while (lo < hi) {
            const int mid = (lo + hi + 1) >> 1;
            if (rows[mid].word_off <= wi) lo = mid; else hi = mid - 1;
        }
        const smg_row_t row = rows[lo];
        const int64_t q0 = (wi - row.word_off) * 16;
        const int64_t so = src_off[lo];
        uint8_t raw[16];
        const int64_t rem = (int64_t)row.n_avail - q0;
        if (rem >= 16 && ((so + q0) & 15) == 0) {
            const uint4 v = *reinterpret_cast<const uint4*>(src + so + q0);
            memcpy(raw, &v, 16);
        } else {
#pragma unroll
            for (int j = 0; j < 16; ++j) raw[j] = (j < rem) ? src[so + q0 + j] : 0;
        }
        uint32_t pw = 0, aw = 0;
        uint8_t cd[16];
        bool any_amb = false;
#pragma unroll
        for (int j = 0; j < 16; ++j) {
            int code = 0;
            if (j < rem) {
                code = kind ? (int)raw[j] : (int)s_lut[raw[j]];
                if (code > 16) {  // KeyError in the reference (encode.py:66)
                    atomicMin(&status[0], (unsigned long long)(so + q0 + j));
                    code = 0;
                }
                if (code < 1 || code > 5) any_amb = true;
            }
            const bool plain = code >= 1 && code <= 5;
            const uint32_t two = plain ? (code == 5 ? 3u : (uint32_t)(code - 1)) : 0u;
            pw |= two << (2 * j);
            aw |= (plain ? 0u : 1u) << j;
            cd[j] = (uint8_t)code;
        }
        packed[wi] = pw;
        amask[wi] = (uint16_t)aw;
        if (codes && aw) {  // codes are only ever read where the mask is set: skip the 1 B/base write for plain words
            uint4 v;
            memcpy(&v, cd, 16);
            *reinterpret_cast<uint4*>(codes + wi * 16) = v;
        }
        if (any_amb) status[1] = 1ull;
    }
}

// Fallback for PWMs wider than SMG_MAX_REGTAB_WIDTH: one thread per window start, tables from global memory.
__global__ void __launch_bounds__(128) scan_wide_kernel(const ScanParams p, const int32_t* __restrict__ wide, const int n_wide)
{
    const smg_chunk_t ch = p.chunks[blockIdx.x];
    const smg_row_t row = p.rows[ch.row];
    const int c0 = ch.start;
    const int c1 = min(c0 + p.chunk_len, row.n_own);
    const int key_shift = p.pos_bits + p.motif_bits;
    for (int s = c0 + threadIdx.x; s < c1; s += blockDim.x) {
        for (int i = 0; i < n_wide; ++i) {
            const int motif = wide[i];
            const int wm = p.motif_width[motif];
            const long long smax64 = row.g_len - row.g_off - (long long)wm;
            if ((long long)s > smax64) continue;
            for (int st = 0; st < 2; ++st) {
                if (!((p.strands >> st) & 1)) continue;
                const float v = window_score32(p, row, motif, st, s);
                if (!(v > p.thr_lo[motif])) continue;
                if (v > p.thr_hi[motif]) {
                    atomicAdd(&p.counters[SMG_CTR_DEFINITE], 1ull);
                    if (p.counts) atomicAdd(p.counts + ((size_t)ch.row * p.n_motifs + motif) * 2 + st, 1);
                    if (p.hit_keys) {
                        const unsigned long long idx = atomicAdd(&p.counters[SMG_CTR_HITS_APPENDED], 1ull);
                        if (idx < p.hit_cap) {
                            const long long pos = st ? (smax64 - (long long)s) : (row.g_off + (long long)s);
                            p.hit_keys[idx] = ((unsigned long long)(unsigned)p.out_rows[ch.row * 2 + st] << key_shift) |
                                              ((unsigned long long)pos << p.motif_bits) | (unsigned long long)(unsigned)motif;
                            p.hit_scores[idx] = v;
                        }
                    }
                } else {
                    const unsigned long long idx = atomicAdd(&p.counters[SMG_CTR_NEAR], 1ull);
                    if (idx < p.near_cap) {
                        smg_near_t n;
                        n.row = ch.row; n.motif = motif; n.strand = st; n.start = s; n.score = v; n.reserved = 0;
                        p.near[idx] = n;
                    }
                }
            }
        }
    }
}

__global__ void __launch_bounds__(128) adjudicate_kernel(const ScanParams p, const double* __restrict__ thr64)
{
    const unsigned long long n_near = min(p.counters[SMG_CTR_NEAR], p.near_cap);
    const int key_shift = p.pos_bits + p.motif_bits;
    for (unsigned long long i = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x; i < n_near;
         i += (unsigned long long)gridDim.x * blockDim.x) {
        const smg_near_t n = p.near[i];
        const smg_row_t row = p.rows[n.row];
        const int wm = p.motif_width[n.motif];
        const double s64 = window_score64(p.packed, p.amask, p.codes, row, p.nat_tab + p.motif_off[n.motif] * 4, wm, n.strand,
                                          n.start, -1, 0);
        if (s64 > thr64[n.motif]) {
            atomicAdd(&p.counters[SMG_CTR_ADJ_ACCEPTED], 1ull);
            if (p.counts) atomicAdd(p.counts + ((size_t)n.row * p.n_motifs + n.motif) * 2 + n.strand, 1);
            if (p.hit_keys) {
                const unsigned long long idx = atomicAdd(&p.counters[SMG_CTR_HITS_APPENDED], 1ull);
                if (idx < p.hit_cap) {
                    const long long smax64 = row.g_len - row.g_off - (long long)wm;
                    const long long pos = n.strand ? (smax64 - (long long)n.start) : (row.g_off + (long long)n.start);
                    p.hit_keys[idx] = ((unsigned long long)(unsigned)p.out_rows[n.row * 2 + n.strand] << key_shift) |
                                      ((unsigned long long)pos << p.motif_bits) | (unsigned long long)(unsigned)n.motif;
                    p.hit_scores[idx] = n.score;
                }
            }
        }
    }
}

__global__ void __launch_bounds__(256) dense_kernel(const ScanParams p, const int n_rows, const int row_len, float* __restrict__ out)
{
    const int64_t total = (int64_t)n_rows * row_len * p.n_motifs;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        const int m = (int)(i % p.n_motifs);
        const int64_t rp = i / p.n_motifs;
        const int pos = (int)(rp % row_len);
        const int r = (int)(rp / row_len);
        out[i] = window_score32(p, p.rows[r], m, 0, pos);
    }
}

__global__ void __launch_bounds__(128) variant_sites_kernel(const ScanParams p, const int32_t* __restrict__ site_row,
                                                            const int32_t* __restrict__ site_start, const int32_t* __restrict__ site_motif,
                                                            const int32_t* __restrict__ rel_pos, const uint8_t* __restrict__ alt_code,
                                                            const int64_t n, double* __restrict__ out_ref, double* __restrict__ out_alt)
{
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        const smg_row_t row = p.rows[site_row[i]];
        const int m = site_motif[i];
        const int wm = p.motif_width[m];
        const float* tab = p.nat_tab + p.motif_off[m] * 4;
        const int64_t s = site_start[i];
        out_ref[i] = window_score64(p.packed, p.amask, p.codes, row, tab, wm, 0, s, -1, 0);
        out_alt[i] = window_score64(p.packed, p.amask, p.codes, row, tab, wm, 0, s, s + rel_pos[i], alt_code[i]);
    }
}

// Batched SNV scoring (BASELINE config 5).  One warp per (32-motif block, SNV tile); lane = motif; tables are read
// from global memory (L1-resident: 32 motifs x W x 4 floats), the 2W-1 bases around the SNV are warp-uniform.
__global__ void __launch_bounds__(128) variant_windows_kernel(const ScanParams p, const int32_t* __restrict__ snv_row,
                                                              const int32_t* __restrict__ snv_pos, const uint8_t* __restrict__ snv_alt,
                                                              const int64_t n_snv, float* __restrict__ ref_max, float* __restrict__ alt_max,
                                                              const int32_t* __restrict__ motif_list, const int n_list,
                                                              const unsigned int* __restrict__ slow_list,
                                                              const unsigned int* __restrict__ slow_count,
                                                              const int64_t snv_stride, const int64_t motif_stride)
{
    const int lane = threadIdx.x & 31;
    const int warp_in_block = threadIdx.x >> 5;
    const int li = blockIdx.y * 32 + lane;
    const bool active = li < n_list;
    const int m = active ? motif_list[li] : 0;
    const int wm = active ? p.motif_width[m] : 0;
    const float* tab = active ? p.nat_tab + p.motif_off[m] * 4 : nullptr;
    const float ninf = __int_as_float(0xff800000);
    // slow_list != nullptr: only the SNVs the register-table kernels could not take (usually none: the grid exits at once)
    const int64_t n_todo = slow_list ? (int64_t)*slow_count : n_snv;
    for (int64_t vi = (int64_t)blockIdx.x * 4 + warp_in_block; vi < n_todo; vi += (int64_t)gridDim.x * 4) {
        const int64_t v = slow_list ? (int64_t)slow_list[vi] : vi;
        const smg_row_t row = p.rows[snv_row[v]];
        const int64_t pos = snv_pos[v];
        const int alt = snv_alt[v] <= 16 ? snv_alt[v] : 0;
        float rmax = ninf, amax = ninf;
        if (active) {
            const int64_t s_lo = max((int64_t)0, pos - wm + 1);
            const int64_t s_hi = min(pos, row.g_len - row.g_off - (int64_t)wm);
            for (int64_t s = s_lo; s <= s_hi; ++s) {
                for (int st = 0; st < 2; ++st) {
                    float ra = 0.f, aa = 0.f;
                    for (int k = 0; k < wm; ++k) {
                        const int c = smg_fetch_code(p.packed, p.amask, p.codes, row, s + k);
                        const int ca = (s + k == pos) ? alt : c;
                        const float* tr = tab + (st ? (wm - 1 - k) : k) * 4;
                        const float t0 = st ? tr[3] : tr[0], t1 = st ? tr[2] : tr[1], t2 = st ? tr[1] : tr[2], t3 = st ? tr[0] : tr[3];
                        const float tr_ref = smg_mix(c_emb[c][0], c_emb[c][1], c_emb[c][2], c_emb[c][3], t0, t1, t2, t3);
                        const float tr_alt = smg_mix(c_emb[ca][0], c_emb[ca][1], c_emb[ca][2], c_emb[ca][3], t0, t1, t2, t3);
                        ra = (k == 0) ? tr_ref : __fadd_rn(ra, tr_ref);
                        aa = (k == 0) ? tr_alt : __fadd_rn(aa, tr_alt);
                    }
                    rmax = fmaxf(rmax, ra);
                    amax = fmaxf(amax, aa);
                }
            }
            ref_max[v * snv_stride + m * motif_stride] = rmax;
            alt_max[v * snv_stride + m * motif_stride] = amax;
        }
    }
}

// per-column prefilter thresholds of the tensor-core path (depend on the call's thresholds and strand mask)
__global__ void hmma_thr_kernel(const int32_t* __restrict__ col_id, const float* __restrict__ thr_pre, const int strands,
                                const int64_t n, float* __restrict__ out)
{
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const int cid = col_id[i];
    out[i] = (cid >= 0 && ((strands >> (cid & 1)) & 1)) ? thr_pre[cid >> 1] : __int_as_float(0x7f800000);
}

// Exact second pass of the tensor-core path: every candidate (row, start, motif, strand) emitted by the fp16 prefilter is
#ifndef SMG_FINALIZE_MIN_BLOCKS
#define SMG_FINALIZE_MIN_BLOCKS 4
#endif
// re-scored in fp32 in the register-table kernel's summation order, the near-threshold band is adjudicated in fp64,
// and hits / counts are produced exactly as smg_scan + smg_adjudicate would.
__global__ void __launch_bounds__(256, SMG_FINALIZE_MIN_BLOCKS) finalize_candidates_kernel(const ScanParams p, const unsigned long long* __restrict__ cand,
                                                                  const unsigned long long cand_cap, const int cand_shift,
                                                                  const double* __restrict__ thr64)
{
    const unsigned long long n_cand = min(p.counters[SMG_CTR_CANDIDATES], cand_cap);
    const int key_shift = p.pos_bits + p.motif_bits;
    const unsigned lane = threadIdx.x & 31u;
    // whole blocks stay convergent: the hit list is extended with ONE atomic per block and iteration (the append counter
    // is a single address; one atomic per warp serialised in the L2 atomic unit)
    const unsigned long long n_round = (n_cand + 255ull) & ~255ull;
    __shared__ unsigned s_cnt[8];
    __shared__ unsigned long long s_base;
    const unsigned warp = threadIdx.x >> 5;
    unsigned n_def = 0, n_near = 0, n_acc = 0;  // per-thread tallies, one atomic per warp at the end
    for (unsigned long long i = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x; i < n_round;
         i += (unsigned long long)gridDim.x * blockDim.x) {
        bool hit = false;
        float s32 = 0.f;
        unsigned long long key = 0;
        if (i < n_cand) {
            const unsigned long long c = cand[i];
            if (c != SMG_CAND_INVALID) {  // padding of a block reserved by the tcgen05 prefilter's consumer warps
            const int st = (int)(c & 1ull);
            const int motif = (int)((c >> 1) & ((1ull << p.motif_bits) - 1ull));
            const long long s = (long long)((c >> (p.motif_bits + 1)) & ((1ull << (cand_shift - p.motif_bits - 1)) - 1ull));
            const int r = (int)(c >> cand_shift);
            const smg_row_t row = p.rows[r];
            const int wm = p.motif_width[motif];
            const long long smax64 = row.g_len - row.g_off - (long long)wm;
            if (s <= smax64 && ((p.strands >> st) & 1)) {
                if (!window_score32_acgt(p, row, motif, st, s, &s32)) s32 = window_score32(p, row, motif, st, s);
                if (s32 > p.thr_hi[motif]) {
                    hit = true;
                    ++n_def;
                } else if (s32 > p.thr_lo[motif]) {
                    ++n_near;
                    const double s64 = window_score64(p.packed, p.amask, p.codes, row, p.nat_tab + p.motif_off[motif] * 4, wm, st,
                                                      s, -1, 0);
                    if (s64 > thr64[motif]) {
                        hit = true;
                        ++n_acc;
                    }
                }
                if (hit) {
                    if (p.counts) atomicAdd(p.counts + ((size_t)r * p.n_motifs + motif) * 2 + st, 1);
                    const long long pos = st ? (smax64 - s) : (row.g_off + s);
                    key = ((unsigned long long)(unsigned)p.out_rows[r * 2 + st] << key_shift) |
                          ((unsigned long long)pos << p.motif_bits) | (unsigned long long)(unsigned)motif;
                }
            }
            }
        }
        if (p.hit_keys) {
            const unsigned m = __ballot_sync(SMG_FULL_MASK, hit);
            if (lane == 0) s_cnt[warp] = (unsigned)__popc(m);
            __syncthreads();
            if (threadIdx.x == 0) {
                unsigned tot = 0;
#pragma unroll
                for (int w = 0; w < 8; ++w) { const unsigned c = s_cnt[w]; s_cnt[w] = tot; tot += c; }
                s_base = tot ? atomicAdd(&p.counters[SMG_CTR_HITS_APPENDED], (unsigned long long)tot) : 0ull;
            }
            __syncthreads();
            if (hit) {
                const unsigned long long slot = s_base + s_cnt[warp] + __popc(m & ((1u << lane) - 1u));
                if (slot < p.hit_cap) {
                    p.hit_keys[slot] = key;
                    p.hit_scores[slot] = s32;
                }
            }
            __syncthreads();  // s_cnt / s_base are rewritten in the next iteration
        }
    }
    n_def = __reduce_add_sync(SMG_FULL_MASK, n_def);
    n_near = __reduce_add_sync(SMG_FULL_MASK, n_near);
    n_acc = __reduce_add_sync(SMG_FULL_MASK, n_acc);
    if (lane == 0) {
        if (n_def) atomicAdd(&p.counters[SMG_CTR_DEFINITE], (unsigned long long)n_def);
        if (n_near) atomicAdd(&p.counters[SMG_CTR_NEAR_TC], (unsigned long long)n_near);
        if (n_acc) atomicAdd(&p.counters[SMG_CTR_ADJ_ACCEPTED], (unsigned long long)n_acc);
    }
}

// scan._bin_sites_by_score (scan.py:77-111) on the device: frac = score / max_score in fp64, pd.cut semantics
// (right-closed intervals (e[b], e[b+1]]; values outside (e[0], e[n_bins]] fall in no bin), one histogram cell per
// (output row, motif, bin).
__global__ void __launch_bounds__(256) bin_hits_kernel(const unsigned long long* __restrict__ keys, const float* __restrict__ scores,
                                                       const unsigned long long n_hits, const int pos_bits, const int motif_bits,
                                                       const double* __restrict__ max_scores, const double* __restrict__ edges,
                                                       const int n_bins, int* __restrict__ hist, const int n_motifs)
{
    for (unsigned long long i = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x; i < n_hits;
         i += (unsigned long long)gridDim.x * blockDim.x) {
        const unsigned long long k = keys[i];
        const int motif = (int)(k & ((1ull << motif_bits) - 1ull));
        const long long row = (long long)(k >> (pos_bits + motif_bits));
        const double frac = (double)scores[i] / max_scores[motif];
        if (!(frac > edges[0]) || !(frac <= edges[n_bins])) continue;
        int lo = 0, hi = n_bins - 1;  // largest b with edges[b] < frac
        while (lo < hi) {
            const int mid = (lo + hi + 1) >> 1;
            if (edges[mid] < frac) lo = mid; else hi = mid - 1;
        }
        atomicAdd(hist + ((size_t)row * n_motifs + motif) * n_bins + lo, 1);
    }
}

// ------------------------------------------------------------------------------------------ host helpers
static int init_lut()
{
    static bool done[64] = {false};
    int dev = 0;
    SMG_CUDA(cudaGetDevice(&dev));
    if (dev < 64 && done[dev]) return SMG_OK;
    cudaDeviceProp prop;
    SMG_CUDA(cudaGetDeviceProperties(&prop, dev));
    if (prop.major != 10) return fail(SMG_ERR_UNSUPPORTED, "smeagol_b200 is built for sm_100a (B200) only");
    uint8_t lut[256];
    memset(lut, 255, sizeof(lut));
    const char* alphabet = "ZACGTUNWSMKRYBDHV";  // encode.py:11-41, uppercase only (lowercase raises KeyError there)
    for (int i = 0; i < 17; ++i) lut[(unsigned char)alphabet[i]] = (uint8_t)i;
    SMG_CUDA(cudaMemcpyToSymbol(c_lut, lut, 256));
    if (dev < 64) done[dev] = true;
    return SMG_OK;
}

template <typename T>
static int upload(T** dptr, const std::vector<T>& h)
{
    const size_t bytes = std::max<size_t>(h.size(), 1) * sizeof(T);
    SMG_CUDA(cudaMalloc((void**)dptr, bytes));
    if (!h.empty()) SMG_CUDA(cudaMemcpy(*dptr, h.data(), h.size() * sizeof(T), cudaMemcpyHostToDevice));
    return SMG_OK;
}

// used by the other translation units of the library (fasta.cu)
int smg_fail_from(int code, const std::string& msg) { return fail(code, msg); }
void smg_count_launch() { ++g_launches; }

// ------------------------------------------------------------------------------------------ C ABI
extern "C" {

int smg_version(void) { return 100; }
const char* smg_last_error(void) { return g_err.c_str(); }
uint64_t smg_launch_count(void) { return g_launches.load(); }

int smg_pwmset_create(const float* h_weights, const int32_t* h_widths, int32_t n_motifs, smg_pwmset_t* out)
{
    return smg_pwmset_create_split(h_weights, h_widths, n_motifs, 0, out);
}

int smg_pwmset_create_split(const float* h_weights, const int32_t* h_widths, int32_t n_motifs, int32_t tc_min_width,
                            smg_pwmset_t* out)
{
    return smg_pwmset_create_split3(h_weights, h_widths, n_motifs, 0, tc_min_width, out);
}

int smg_pwmset_create_split3(const float* h_weights, const int32_t* h_widths, int32_t n_motifs, int32_t regtab_min_width,
                             int32_t tc_min_width, smg_pwmset_t* out)
{
    if (!h_weights || !h_widths || n_motifs <= 0 || !out) return fail(SMG_ERR_INVALID, "smg_pwmset_create: bad argument");
    if (regtab_min_width < 0 || regtab_min_width > SMG_KMER_MAX_WIDTH + 1)
        return fail(SMG_ERR_INVALID, "smg_pwmset_create_split3: regtab_min_width must be in 0..13");
    if (regtab_min_width > 1 && tc_min_width > 0 && tc_min_width < regtab_min_width)
        return fail(SMG_ERR_INVALID, "smg_pwmset_create_split3: tc_min_width must not be below regtab_min_width");
    if (tc_min_width < 0 || tc_min_width > SMG_MAX_REGTAB_WIDTH + 1)
        return fail(SMG_ERR_INVALID, "smg_pwmset_create_split: tc_min_width must be 0 (no split) or a width in 1..33");
    int rc = init_lut();
    if (rc != SMG_OK) return rc;
    smg_pwmset* s = new smg_pwmset();
    s->n_motifs = n_motifs;
    s->split = tc_min_width;
    s->kmin = regtab_min_width;
    s->widths.assign(h_widths, h_widths + n_motifs);
    s->offs.resize(n_motifs);
    int64_t tot = 0;
    for (int m = 0; m < n_motifs; ++m) {
        if (h_widths[m] <= 0) { delete s; return fail(SMG_ERR_INVALID, "smg_pwmset_create: PWM width must be >= 1"); }
        s->offs[m] = tot;
        tot += h_widths[m];
        s->max_w = std::max(s->max_w, (int)h_widths[m]);
    }
    s->all_finite = true;
    for (int64_t i = 0; i < tot * 4; ++i) s->all_finite = s->all_finite && std::isfinite(h_weights[i]);
    // Non-finite weights (log2 of a zero probability) are stored as given.  The reference's one-hot contraction turns
    // every window of such a PWM into NaN (never a hit); the kernels add the selected weight only, so the caller
    // reproduces that by passing +inf thresholds for these PWMs (smeagol_b200/device.py does).

    // ---- lane groups: PWMs sorted by width; a lane holds NM = 2 PWMs while 2 * width <= 2 * kPairMaxWidth (more adds
    // per uniform branch), else 1; 32 * NM PWMs per group; remainders are carried up to the next width (zero padding)
    {
        std::vector<int> pool;
        auto emit_group = [&](int wb, int nm, size_t n_take) {
            Group g;
            g.wb = std::max(wb, 2);
            g.nm = nm;
            g.tc = 0;
            g.tab_off = 0;
            for (int m = 0; m < SMG_MAX_PWMS_PER_LANE; ++m)
                for (int l = 0; l < 32; ++l) g.lane_motif[m][l] = -1;
            for (size_t i = 0; i < n_take; ++i) g.lane_motif[i / 32][i % 32] = pool[i];
            // groups never mix the two sides (the pool is flushed at the split), so the first PWM decides; the padded
            // width must not: width-1 PWMs are padded to wb = 2 and still belong to the register-table side of split 2
            if (s->split > 0 && n_take > 0 && h_widths[pool[0]] >= s->split) g.tc = 1;
            pool.erase(pool.begin(), pool.begin() + n_take);
            s->groups.push_back(g);
        };
        int last_w = 0;
        for (int w = 1; w <= kDualMaxWidth; ++w) {
            bool any = false;
            for (int m = 0; m < n_motifs; ++m)
                if (h_widths[m] == w && w >= s->kmin) { pool.push_back(m); any = true; }
            if (any) last_w = w;
            const int nm = (std::max(w, 2) <= kPairMaxWidth) ? 2 : 1;
            while (pool.size() >= (size_t)32 * nm) emit_group(w, nm, (size_t)32 * nm);
            // hybrid engine: lane groups never mix PWMs from both sides of the split
            if (s->split > 0 && w == s->split - 1)
                while (!pool.empty()) emit_group(last_w, 1, std::min(pool.size(), (size_t)32));
        }
        while (!pool.empty()) {
            const int nm = (std::max(last_w, 2) <= kPairMaxWidth && pool.size() > 32) ? 2 : 1;
            emit_group(last_w, nm, std::min(pool.size(), (size_t)32 * nm));
        }
    }
    for (int m = 0; m < n_motifs; ++m)
        if (h_widths[m] > SMG_MAX_REGTAB_WIDTH) s->wide.push_back(m);
    // buckets of equal (wb, nm), groups reordered so that a bucket is contiguous
    std::stable_sort(s->groups.begin(), s->groups.end(), [](const Group& a, const Group& b) {
        if (a.nm != b.nm) return a.nm > b.nm;
        if (a.tc != b.tc) return a.tc < b.tc;
        return a.wb < b.wb;
    });
    std::vector<float> gtab;
    std::vector<int64_t> gtab_off;
    std::vector<int32_t> lane_motif;
    for (size_t gi = 0; gi < s->groups.size(); ++gi) {
        Group& g = s->groups[gi];
        if (s->buckets.empty() || s->buckets.back().wb != g.wb || s->buckets.back().nm != g.nm || s->buckets.back().tc != g.tc)
            s->buckets.push_back({g.wb, g.nm, g.tc, (int)gi, 0});
        s->buckets.back().n_groups++;
        g.tab_off = (int64_t)gtab.size();
        gtab_off.push_back(g.tab_off);
        const size_t slot_floats = (size_t)g.wb * 4 * 32 * 2;
        gtab.resize(gtab.size() + slot_floats * g.nm, 0.f);
        s->sum_padded += (int64_t)g.wb * 2 * g.nm;
        for (int sm = 0; sm < SMG_MAX_PWMS_PER_LANE; ++sm) {
            for (int l = 0; l < 32; ++l) {
                const int m = g.lane_motif[sm][l];
                lane_motif.push_back(m);
                if (m < 0 || sm >= g.nm) continue;
                float* gt = gtab.data() + g.tab_off + slot_floats * sm;
                const int wm = h_widths[m];
                const float* T = h_weights + s->offs[m] * 4;
                // the forward table is zero-padded at the tail, the reverse-complement table
                // Trc[k][b] = T[W-1-k][3-b] at the HEAD, so that rc[k][b] == fwdpad[WB-1-k][3-b] for every k (the
                // kernel keeps only half of the {fwd, rc} pairs and reads the others half-swapped)
                for (int k = 0; k < g.wb; ++k)
                    for (int b = 0; b < 4; ++b) {
                        const size_t e = (size_t)(k * 4 + b) * 32 + l;
                        const int kk = g.wb - 1 - k;
                        gt[e * 2 + 0] = k < wm ? T[k * 4 + b] : 0.f;
                        gt[e * 2 + 1] = kk < wm ? T[kk * 4 + (3 - b)] : 0.f;
                    }
            }
        }
    }
    // ---- tensor-core prefilter tables: fp16 A fragments in mma.m16n8k16 register order
    std::vector<uint32_t> afrag;
    std::vector<int64_t> cg_off;
    std::vector<int32_t> col_id;
    {
        // 16-column blocks per warp for each K depth; SMEAGOL_B200_HMMA_NB="n1,n2,...,n8" overrides (tuning only)
        int nb_tab[9] = {0, 4, 3, 4, 4, 4, 4, 3, 3};  // measured (tools/sweep_nb.sh); K = 32 is hit-dense: smaller tiles, 6 warps
        if (const char* e = getenv("SMEAGOL_B200_HMMA_NB")) {
            int k = 1;
            for (const char* q = e; *q && k <= 8; ++k) {
                nb_tab[k] = atoi(q);
                while (*q && *q != ',') ++q;
                if (*q == ',') ++q;
            }
        }
        auto nb_for = [&](int ks) { return nb_tab[ks]; };
        auto h16 = [](float v) { return (uint32_t)__half_as_ushort(__float2half_rn(v)); };
        for (int ks = 1; ks <= SMG_MAX_REGTAB_WIDTH / 4; ++ks) {
            std::vector<int> cols;  // motif * 2 + strand
            for (int m = 0; m < n_motifs; ++m)
                if ((h_widths[m] + 3) / 4 == ks && (s->split == 0 || h_widths[m] >= s->split) && h_widths[m] >= s->kmin) {  // tensor-core side only
                    cols.push_back(m * 2);
                    cols.push_back(m * 2 + 1);
                }
            if (cols.empty()) continue;
            const int nb = nb_for(ks), per = nb * 16;
            const int ncg = (int)((cols.size() + per - 1) / per);
            // column groups of different buckets have different sizes: col arrays are indexed by colgroup * per within
            // the bucket, so every bucket gets its own base such that (first_colgroup * per) is its first column
            // (and beyond every column-group index already used by a bucket with a different group size)
            const int first_cg = std::max((int)((col_id.size() + per - 1) / per), (int)cg_off.size());
            col_id.resize((size_t)first_cg * per, -1);
            s->hbuckets.push_back({ks, nb, first_cg, ncg});
            for (int cg = 0; cg < ncg; ++cg) {
                while ((int64_t)cg_off.size() < first_cg + cg) cg_off.push_back(0);
                cg_off.push_back((int64_t)afrag.size());
                const size_t base = afrag.size();
                afrag.resize(base + (size_t)nb * ks * 4 * 32, 0u);
                for (int c = 0; c < per; ++c) {
                    const size_t ci = (size_t)cg * per + c;
                    const int cid = ci < cols.size() ? cols[ci] : -1;
                    col_id.push_back(cid);
                    if (cid < 0) continue;
                    const int m = cid >> 1, st = cid & 1, wm = h_widths[m];
                    const float* T = h_weights + s->offs[m] * 4;
                    const int blk = c / 16, rrow = c % 16;
                    for (int kstep = 0; kstep < ks; ++kstep)
                        for (int kin = 0; kin < 16; ++kin) {
                            const int k = kstep * 4 + kin / 4, b = kin % 4;  // PWM row, base
                            float v = 0.f;
                            if (k < wm) v = st ? T[(wm - 1 - k) * 4 + (3 - b)] : T[k * 4 + b];
                            // fragment register r and lane holding (row rrow, k index kin)
                            const int r = (rrow >= 8 ? 1 : 0) + (kin >= 8 ? 2 : 0);
                            const int lane = (rrow % 8) * 4 + (kin % 8) / 2;
                            uint32_t& reg = afrag[base + ((size_t)(blk * ks + kstep) * 4 + r) * 32 + lane];
                            reg |= h16(v) << (16 * (kin & 1));
                        }
                }
            }
        }
        s->n_hcols = (int64_t)col_id.size();
    }
    // ---- tcgen05 prefilter tables: columns sorted by K depth, 128 per block; a block's image is its fp16 weights as the
    // K-major A operand of tcgen05.mma without swizzle: element (column m, k) at (k / 8) * 2048 + m * 16 + (k % 8) * 2,
    // k = 4 * PWM row + base; blocks are dealt into column groups of at most smg_tc5_max_image_bytes()
    std::vector<uint8_t> t5_img;
    std::vector<int64_t> t5_cg_img_off;
    std::vector<int32_t> t5_cg_first_blk, t5_blk_ks, t5_blk_off, t5_col_id;
    {
        std::vector<int> cols;  // motif * 2 + strand, ascending K depth
        for (int ks = 1; ks <= SMG_MAX_REGTAB_WIDTH / 4; ++ks)
            for (int m = 0; m < n_motifs; ++m)
                if ((h_widths[m] + 3) / 4 == ks && (s->split == 0 || h_widths[m] >= s->split) && h_widths[m] >= s->kmin) {
                    cols.push_back(m * 2);
                    cols.push_back(m * 2 + 1);
                }
        const int n_blocks = (int)((cols.size() + 127) / 128);
        std::vector<int> bks(n_blocks, 1);
        int units = 0;
        for (int b = 0; b < n_blocks; ++b) {
            const size_t last = std::min(cols.size(), (size_t)(b + 1) * 128) - 1;
            bks[b] = (h_widths[cols[last] >> 1] + 3) / 4;  // sorted by K depth: the last column is the deepest
            units += bks[b];
        }
        const int cap_units = std::min(smg_tc5_max_image_bytes() / smg::kTc5UnitBytes, 1 << 20);
        // Column groups: block i goes to group i % G (dealt round-robin, so every group gets the same mix of K depths and
        // hit densities and the CTAs of all groups finish together); G = the smallest count for which every group fits
        auto deal_ok = [&](int G) {
            for (int g = 0; g < G; ++g) {
                int used = 0, nb = 0;
                for (int q = g; q < n_blocks; q += G) { used += bks[q]; ++nb; }
                if (used > cap_units || nb > smg_tc5_max_blocks()) return false;
            }
            return true;
        };
        int n_groups = 0;
        if (n_blocks) {
            n_groups = std::max(1, (units + cap_units - 1) / cap_units);
            while (n_groups < n_blocks && !deal_ok(n_groups)) ++n_groups;
        }
        auto h16 = [](float v) {
            if (v == v && v > 60000.f && v < 3.0e38f) v = 60000.f;   // finite weights stay finite in fp16
            if (v == v && v < -60000.f && v > -3.0e38f) v = -60000.f;
            return __half_as_ushort(__float2half_rn(v));
        };
        int emitted = 0;
        for (int g = 0; g < n_groups; ++g) {
            t5_cg_first_blk.push_back(emitted);
            t5_cg_img_off.push_back((int64_t)t5_img.size());
            int used = 0;
            for (int b = g; b < n_blocks; b += n_groups) {
                const int ks = bks[b];
                t5_blk_ks.push_back(ks);
                t5_blk_off.push_back(used * smg::kTc5UnitBytes);
                const size_t base = t5_img.size();
                t5_img.resize(base + (size_t)ks * smg::kTc5UnitBytes, 0);
                for (int mrow = 0; mrow < 128; ++mrow) {
                    const size_t ci = (size_t)b * 128 + mrow;
                    const int cid = ci < cols.size() ? cols[ci] : -1;
                    t5_col_id.push_back(cid);
                    if (cid < 0) continue;
                    const int m = cid >> 1, st = cid & 1, wm = h_widths[m];
                    const float* T = h_weights + s->offs[m] * 4;
                    for (int k = 0; k < 4 * wm; ++k) {
                        const int row = k >> 2, bb = k & 3;
                        const float v = st ? T[(wm - 1 - row) * 4 + (3 - bb)] : T[row * 4 + bb];
                        const uint16_t hv = h16(v);
                        memcpy(&t5_img[base + (size_t)(k / 8) * 2048 + (size_t)mrow * 16 + (size_t)(k % 8) * 2], &hv, 2);
                    }
                }
                used += ks;
                ++emitted;
            }
            s->t5_max_image = std::max(s->t5_max_image, used * smg::kTc5UnitBytes);
        }
        if (n_blocks && (emitted != n_blocks || !deal_ok(n_groups))) {
            delete s;
            return fail(SMG_ERR_INVALID, "smg_pwmset_create: tcgen05 column groups do not cover the library");
        }
        t5_cg_first_blk.push_back(n_blocks);
        s->t5_n_blocks = n_blocks;
        s->t5_n_groups = n_groups;
    }
    std::vector<int32_t> cls_motifs;
    for (int w = 1; w <= SMG_KMER_MAX_WIDTH; ++w)
        for (int m = 0; m < n_motifs; ++m)
            if (h_widths[m] == w) cls_motifs.push_back(m);
    std::vector<float> nat(h_weights, h_weights + tot * 4);
    std::vector<int32_t> wv(h_widths, h_widths + n_motifs);
    std::vector<int32_t> widev(s->wide.begin(), s->wide.end());
    for (int m = 0; m < n_motifs; ++m) (h_widths[m] <= kDualMaxWidth ? s->dual_motifs : s->nondual_motifs).push_back(m);
    std::vector<int32_t> dualv(s->dual_motifs.begin(), s->dual_motifs.end()), nondualv(s->nondual_motifs.begin(), s->nondual_motifs.end());
    s->h_nat = nat;
    if ((rc = upload(&s->d_nat, nat)) || (rc = upload(&s->d_width, wv)) || (rc = upload(&s->d_off, s->offs)) ||
        (rc = upload(&s->d_gtab, gtab)) || (rc = upload(&s->d_gtab_off, gtab_off)) ||
        (rc = upload(&s->d_lane_motif, lane_motif)) ||
        (rc = upload(&s->d_afrag, afrag)) || (rc = upload(&s->d_cg_afrag_off, cg_off)) || (rc = upload(&s->d_col_id, col_id)) ||
        (rc = upload(&s->d_thr_pre_col, std::vector<float>((size_t)std::max<int64_t>(s->n_hcols, 1), 0.f))) ||
        (rc = upload(&s->d_t5_img, t5_img)) || (rc = upload(&s->d_t5_cg_img_off, t5_cg_img_off)) ||
        (rc = upload(&s->d_t5_cg_first_blk, t5_cg_first_blk)) || (rc = upload(&s->d_t5_blk_ks, t5_blk_ks)) ||
        (rc = upload(&s->d_t5_blk_off, t5_blk_off)) || (rc = upload(&s->d_t5_col_id, t5_col_id)) ||
        (rc = upload(&s->d_t5_thr_pre_col, std::vector<float>((size_t)std::max(s->t5_n_blocks * 128, 1), 0.f))) ||
        (rc = upload(&s->d_cls_motifs, cls_motifs)) ||
        (rc = upload(&s->d_wide, widev)) || (rc = upload(&s->d_dual_list, dualv)) || (rc = upload(&s->d_nondual_list, nondualv))) {
        smg_pwmset_destroy(s);
        return rc;
    }
    *out = s;
    return SMG_OK;
}

int smg_pwmset_destroy(smg_pwmset_t s)
{
    if (!s) return SMG_OK;
    cudaFree(s->d_nat); cudaFree(s->d_width); cudaFree(s->d_off); cudaFree(s->d_gtab); cudaFree(s->d_gtab_off);
    cudaFree(s->d_cls_motifs);
    for (auto& kv : s->seeds) {
        smg_pwmset::SeedSet* q = kv.second;
        cudaFree(q->d_vtab); cudaFree(q->d_voff); cudaFree(q->d_vwidth); cudaFree(q->d_vcol); cudaFree(q->d_voffset);
        cudaFree(q->d_vrest); cudaFree(q->d_vthr); cudaFree(q->d_vthr64); cudaFree(q->d_cls);
        delete q;
    }
    cudaFree(s->d_t5_img); cudaFree(s->d_t5_cg_img_off); cudaFree(s->d_t5_cg_first_blk); cudaFree(s->d_t5_blk_ks);
    cudaFree(s->d_t5_blk_off); cudaFree(s->d_t5_col_id); cudaFree(s->d_t5_thr_pre_col);
    cudaFree(s->d_lane_motif); cudaFree(s->d_afrag); cudaFree(s->d_cg_afrag_off); cudaFree(s->d_col_id); cudaFree(s->d_thr_pre_col); cudaFree(s->d_wide); cudaFree(s->d_dual_list); cudaFree(s->d_nondual_list);
    delete s;
    return SMG_OK;
}

int smg_pwmset_info(smg_pwmset_t s, int64_t info[8])
{
    if (!s || !info) return fail(SMG_ERR_INVALID, "smg_pwmset_info: bad argument");
    memset(info, 0, 8 * sizeof(int64_t));
    info[0] = s->n_motifs; info[1] = s->max_w; info[2] = (int64_t)s->groups.size(); info[3] = (int64_t)s->buckets.size();
    info[4] = (int64_t)s->wide.size(); info[5] = s->sum_padded; info[6] = s->t5_n_blocks; info[7] = s->t5_n_groups;
    return SMG_OK;
}

int smg_pack(const uint8_t* d_src, int input_kind, const smg_row_t* d_rows, const int64_t* d_src_off, int32_t n_rows,
             int64_t total_words, uint32_t* d_packed, uint16_t* d_amask, uint8_t* d_codes, int64_t* d_status, void* stream)
{
    if (!d_src || !d_rows || !d_src_off || n_rows <= 0 || total_words <= 0 || !d_packed || !d_amask || !d_status)
        return fail(SMG_ERR_INVALID, "smg_pack: bad argument");
    int rc = init_lut();
    if (rc != SMG_OK) return rc;
    const int64_t blocks = std::min<int64_t>((total_words + 255) / 256, 148 * 64);
    pack_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(d_src, input_kind, d_rows, d_src_off, n_rows, total_words,
                                                                    d_packed, d_amask, d_codes,
                                                                    reinterpret_cast<unsigned long long*>(d_status));
    SMG_LAUNCH_CHECK("pack_kernel");
    return SMG_OK;
}

static void fill_params(ScanParams& p, smg_pwmset_t s, const uint32_t* d_packed, const uint16_t* d_amask, const uint8_t* d_codes,
                        const smg_row_t* d_rows, const int32_t* d_out_rows)
{
    memset(&p, 0, sizeof(p));
    p.packed = d_packed; p.amask = d_amask; p.codes = d_codes; p.rows = d_rows; p.out_rows = d_out_rows;
    p.n_motifs = s->n_motifs; p.motif_width = s->d_width; p.motif_off = s->d_off; p.nat_tab = s->d_nat;
    p.group_tab = s->d_gtab; p.group_tab_off = s->d_gtab_off; p.lane_motif = s->d_lane_motif;
}

int smg_scan(smg_pwmset_t s, const uint32_t* d_packed, const uint16_t* d_amask, const uint8_t* d_codes, const smg_row_t* d_rows,
             int32_t n_rows, const int32_t* d_out_rows, const smg_chunk_t* d_chunks, int32_t n_chunks, int32_t chunk_len,
             int strands, const float* d_thr_lo, const float* d_thr_hi, int pos_bits, int motif_bits, uint64_t* d_hit_keys,
             float* d_hit_scores, uint64_t hit_cap, smg_near_t* d_near, uint64_t near_cap, int32_t* d_counts,
             uint64_t* d_counters, void* stream)
{
    if (!s || !d_packed || !d_amask || !d_rows || n_rows <= 0 || !d_out_rows || !d_chunks || n_chunks < 0 || !d_thr_lo ||
        !d_thr_hi || !d_counters || !d_near)
        return fail(SMG_ERR_INVALID, "smg_scan: bad argument");
    if (chunk_len <= 0 || (chunk_len & 15)) return fail(SMG_ERR_INVALID, "smg_scan: chunk_len must be a positive multiple of 16");
    if ((strands & 3) == 0 || (strands & ~3)) return fail(SMG_ERR_INVALID, "smg_scan: strands must be 1, 2 or 3");
    if (pos_bits <= 0 || motif_bits <= 0 || pos_bits + motif_bits > 63) return fail(SMG_ERR_INVALID, "smg_scan: bad key layout");
    if ((1ll << motif_bits) < s->n_motifs) return fail(SMG_ERR_INVALID, "smg_scan: motif_bits too small");
    if ((d_hit_keys == nullptr) != (d_hit_scores == nullptr)) return fail(SMG_ERR_INVALID, "smg_scan: hit buffers must both be set or NULL");
    if (n_chunks == 0) return SMG_OK;
    ScanParams p;
    fill_params(p, s, d_packed, d_amask, d_codes, d_rows, d_out_rows);
    p.chunks = d_chunks; p.chunk_len = chunk_len; p.strands = strands; p.thr_lo = d_thr_lo; p.thr_hi = d_thr_hi;
    p.pos_bits = pos_bits; p.motif_bits = motif_bits;
    p.hit_keys = reinterpret_cast<unsigned long long*>(d_hit_keys); p.hit_scores = d_hit_scores; p.hit_cap = hit_cap;
    p.near = d_near; p.near_cap = near_cap; p.counts = d_counts;
    p.counters = reinterpret_cast<unsigned long long*>(d_counters);
    cudaStream_t st = (cudaStream_t)stream;
    // Width buckets are independent: fork them onto side streams so that their tail waves overlap, join on `st`.
    StreamPool* pool = get_pool();
    if (!pool) return fail(SMG_ERR_CUDA, "smg_scan: cannot create the side streams");
    SMG_CUDA(cudaEventRecord(pool->fork, st));
    int bi = 0;
    for (const Bucket& b : s->buckets) {
        if (b.tc) continue;  // the tensor-core engine owns these (smg_scan_tc)
        smg_launch_fn fn = smg_get_launcher(b.wb, b.nm);
        if (!fn) return fail(SMG_ERR_INVALID, "smg_scan: no kernel for padded width " + std::to_string(b.wb));
        cudaStream_t bs = (bi == 0) ? st : pool->side[(bi - 1) % StreamPool::kSide];
        if (bi > 0 && bi <= StreamPool::kSide) SMG_CUDA(cudaStreamWaitEvent(bs, pool->fork, 0));
        for (int g0 = 0; g0 < b.n_groups; g0 += 65535) {  // gridDim.y limit
            ScanParams q = p;
            q.first_group = b.first_group + g0;
            fn(q, n_chunks, std::min(65535, b.n_groups - g0), bs);
            SMG_LAUNCH_CHECK("scan_regtab_kernel");
        }
        ++bi;
    }
    const int used = std::min(bi - 1, (int)StreamPool::kSide);
    for (int i = 0; i < used; ++i) {
        SMG_CUDA(cudaEventRecord(pool->join[i], pool->side[i]));
        SMG_CUDA(cudaStreamWaitEvent(st, pool->join[i], 0));
    }
    if (!s->wide.empty()) {
        scan_wide_kernel<<<(unsigned)n_chunks, 128, 0, st>>>(p, s->d_wide, (int)s->wide.size());
        SMG_LAUNCH_CHECK("scan_wide_kernel");
    }
    return SMG_OK;
}

int smg_scan_tc(smg_pwmset_t s, const uint32_t* d_packed, const uint16_t* d_amask, const uint8_t* d_codes, const smg_row_t* d_rows,
                int32_t n_rows, const int32_t* d_out_rows, const smg_chunk_t* d_chunks, int32_t n_chunks, int32_t chunk_len,
                int strands, const float* d_thr_lo, const float* d_thr_hi, const float* d_thr_pre, const double* d_thr64,
                int pos_bits, int motif_bits, int cand_pos_bits, uint64_t* d_cand, uint64_t cand_cap, uint64_t* d_hit_keys,
                float* d_hit_scores, uint64_t hit_cap, int32_t* d_counts, uint64_t* d_counters, void* stream)
{
    if (!s || !d_packed || !d_amask || !d_rows || n_rows <= 0 || !d_out_rows || !d_chunks || n_chunks < 0 || !d_thr_lo ||
        !d_thr_hi || !d_thr_pre || !d_thr64 || !d_counters || !d_cand || cand_cap == 0)
        return fail(SMG_ERR_INVALID, "smg_scan_tc: bad argument");
    if (chunk_len <= 0 || (chunk_len & 15)) return fail(SMG_ERR_INVALID, "smg_scan_tc: chunk_len must be a positive multiple of 16");
    if ((strands & 3) == 0 || (strands & ~3)) return fail(SMG_ERR_INVALID, "smg_scan_tc: strands must be 1, 2 or 3");
    if (pos_bits <= 0 || motif_bits <= 0 || pos_bits + motif_bits > 63) return fail(SMG_ERR_INVALID, "smg_scan_tc: bad key layout");
    if ((1ll << motif_bits) < s->n_motifs) return fail(SMG_ERR_INVALID, "smg_scan_tc: motif_bits too small");
    if (cand_pos_bits <= 0 || cand_pos_bits + motif_bits + 1 > 62) return fail(SMG_ERR_INVALID, "smg_scan_tc: bad candidate layout");
    if ((d_hit_keys == nullptr) != (d_hit_scores == nullptr)) return fail(SMG_ERR_INVALID, "smg_scan_tc: hit buffers must both be set or NULL");
    if (n_chunks == 0) return SMG_OK;
    {   // nothing on the tensor-core side of the split (e.g. no PWM that wide in the set): no launches at all
        const bool any = !s->hbuckets.empty();  // with a split the fragment tables hold the tensor-core side only
        if (!any && (s->wide.empty() || s->split > 0)) return SMG_OK;
    }
    cudaStream_t st = (cudaStream_t)stream;
    ScanParams p;
    fill_params(p, s, d_packed, d_amask, d_codes, d_rows, d_out_rows);
    p.chunks = d_chunks; p.chunk_len = chunk_len; p.strands = strands; p.thr_lo = d_thr_lo; p.thr_hi = d_thr_hi;
    p.pos_bits = pos_bits; p.motif_bits = motif_bits;
    p.hit_keys = reinterpret_cast<unsigned long long*>(d_hit_keys); p.hit_scores = d_hit_scores; p.hit_cap = hit_cap;
    p.counts = d_counts; p.counters = reinterpret_cast<unsigned long long*>(d_counters);
    if (s->n_hcols > 0) {
        hmma_thr_kernel<<<(unsigned)((s->n_hcols + 255) / 256), 256, 0, st>>>(s->d_col_id, d_thr_pre, strands, s->n_hcols, s->d_thr_pre_col);
        SMG_LAUNCH_CHECK("hmma_thr_kernel");
    }
    smg::HmmaParams hp;
    hp.sp = p; hp.afrag = s->d_afrag; hp.cg_afrag_off = s->d_cg_afrag_off; hp.thr_pre_col = s->d_thr_pre_col; hp.col_id = s->d_col_id;
    hp.cand_shift = cand_pos_bits + motif_bits + 1;
    hp.cand = reinterpret_cast<unsigned long long*>(d_cand); hp.cand_cap = cand_cap;
    StreamPool* pool = get_pool();
    if (!pool) return fail(SMG_ERR_CUDA, "smg_scan_tc: cannot create the side streams");
    SMG_CUDA(cudaEventRecord(pool->fork, st));
    int bi = 0;
    for (const smg_pwmset::HmmaBucket& b : s->hbuckets) {
        smg_hmma_launch_fn fn = smg_get_hmma_launcher(b.ks, b.nb);
        if (!fn) return fail(SMG_ERR_INVALID, "smg_scan_tc: no kernel for KS " + std::to_string(b.ks));
        cudaStream_t bs = (bi == 0) ? st : pool->side[(bi - 1) % StreamPool::kSide];
        if (bi > 0 && bi <= StreamPool::kSide) SMG_CUDA(cudaStreamWaitEvent(bs, pool->fork, 0));
        hp.first_colgroup = b.first_colgroup;
        fn(hp, n_chunks, b.n_colgroups, bs);
        SMG_LAUNCH_CHECK("scan_hmma_kernel");
        ++bi;
    }
    const int used = std::min(std::max(bi - 1, 0), (int)StreamPool::kSide);
    for (int i = 0; i < used; ++i) {
        SMG_CUDA(cudaEventRecord(pool->join[i], pool->side[i]));
        SMG_CUDA(cudaStreamWaitEvent(st, pool->join[i], 0));
    }
    // PWMs wider than the register / fragment kernels: emitted as candidates by a generic kernel? No -- scored directly.
    if (!s->wide.empty() && s->split == 0) return fail(SMG_ERR_UNSUPPORTED, "smg_scan_tc: PWMs wider than 32 need smg_scan");
    finalize_candidates_kernel<<<148 * SMG_FINALIZE_MIN_BLOCKS * 2, 256, 0, st>>>(p, reinterpret_cast<const unsigned long long*>(d_cand), cand_cap,
                                                       hp.cand_shift, d_thr64);
    SMG_LAUNCH_CHECK("finalize_candidates_kernel");
    return SMG_OK;
}

int smg_scan_tc5(smg_pwmset_t s, const uint32_t* d_packed, const uint16_t* d_amask, const uint8_t* d_codes, const smg_row_t* d_rows,
                 int32_t n_rows, const int32_t* d_out_rows, const smg_chunk_t* d_chunks, int32_t n_chunks, int32_t chunk_len,
                 int strands, const float* d_thr_lo, const float* d_thr_hi, const float* d_thr_pre, const double* d_thr64,
                 int pos_bits, int motif_bits, int cand_pos_bits, uint64_t* d_cand, uint64_t cand_cap, uint64_t* d_hit_keys,
                 float* d_hit_scores, uint64_t hit_cap, int32_t* d_counts, uint64_t* d_counters, void* stream)
{
    if (!s || !d_packed || !d_amask || !d_rows || n_rows <= 0 || !d_out_rows || !d_chunks || n_chunks < 0 || !d_thr_lo ||
        !d_thr_hi || !d_thr_pre || !d_thr64 || !d_counters || !d_cand || cand_cap == 0)
        return fail(SMG_ERR_INVALID, "smg_scan_tc5: bad argument");
    if (chunk_len <= 0 || (chunk_len & 15)) return fail(SMG_ERR_INVALID, "smg_scan_tc5: chunk_len must be a positive multiple of 16");
    if ((strands & 3) == 0 || (strands & ~3)) return fail(SMG_ERR_INVALID, "smg_scan_tc5: strands must be 1, 2 or 3");
    if (pos_bits <= 0 || motif_bits <= 0 || pos_bits + motif_bits > 63) return fail(SMG_ERR_INVALID, "smg_scan_tc5: bad key layout");
    if ((1ll << motif_bits) < s->n_motifs) return fail(SMG_ERR_INVALID, "smg_scan_tc5: motif_bits too small");
    if (cand_pos_bits <= 0 || cand_pos_bits + motif_bits + 1 > 62) return fail(SMG_ERR_INVALID, "smg_scan_tc5: bad candidate layout");
    if ((d_hit_keys == nullptr) != (d_hit_scores == nullptr)) return fail(SMG_ERR_INVALID, "smg_scan_tc5: hit buffers must both be set or NULL");
    if (!s->wide.empty() && s->split == 0) return fail(SMG_ERR_UNSUPPORTED, "smg_scan_tc5: PWMs wider than 32 need smg_scan");
    if (n_chunks == 0 || s->t5_n_blocks == 0) return SMG_OK;
    cudaStream_t st = (cudaStream_t)stream;
    ScanParams p;
    fill_params(p, s, d_packed, d_amask, d_codes, d_rows, d_out_rows);
    p.chunks = d_chunks; p.chunk_len = chunk_len; p.strands = strands; p.thr_lo = d_thr_lo; p.thr_hi = d_thr_hi;
    p.pos_bits = pos_bits; p.motif_bits = motif_bits;
    p.hit_keys = reinterpret_cast<unsigned long long*>(d_hit_keys); p.hit_scores = d_hit_scores; p.hit_cap = hit_cap;
    p.counts = d_counts; p.counters = reinterpret_cast<unsigned long long*>(d_counters);
    const int64_t n_cols = (int64_t)s->t5_n_blocks * 128;
    hmma_thr_kernel<<<(unsigned)((n_cols + 255) / 256), 256, 0, st>>>(s->d_t5_col_id, d_thr_pre, strands, n_cols, s->d_t5_thr_pre_col);
    SMG_LAUNCH_CHECK("hmma_thr_kernel");
    smg::Tc5Params hp;
    hp.sp = p; hp.wimg = s->d_t5_img; hp.cg_img_off = s->d_t5_cg_img_off; hp.cg_first_blk = s->d_t5_cg_first_blk;
    hp.blk_ks = s->d_t5_blk_ks; hp.blk_off = s->d_t5_blk_off; hp.thr_pre_col = s->d_t5_thr_pre_col; hp.col_id = s->d_t5_col_id;
    hp.n_chunks = n_chunks; hp.cand_shift = cand_pos_bits + motif_bits + 1;
    hp.cand = reinterpret_cast<unsigned long long*>(d_cand); hp.cand_cap = cand_cap;
    int n_sm = 148;
    {
        int dev = 0;
        SMG_CUDA(cudaGetDevice(&dev));
        SMG_CUDA(cudaDeviceGetAttribute(&n_sm, cudaDevAttrMultiProcessorCount, dev));
    }
    const int grid_x = std::max(1, std::min(n_chunks, n_sm / s->t5_n_groups));
    cudaError_t e = smg_launch_tc5(hp, grid_x, s->t5_n_groups, s->t5_max_image, st);
    if (e != cudaSuccess) return fail(SMG_ERR_CUDA, std::string("scan_tc5_kernel: ") + cudaGetErrorString(e));
    ++g_launches;
    finalize_candidates_kernel<<<148 * SMG_FINALIZE_MIN_BLOCKS * 2, 256, 0, st>>>(p, reinterpret_cast<const unsigned long long*>(d_cand), cand_cap,
                                                       hp.cand_shift, d_thr64);
    SMG_LAUNCH_CHECK("finalize_candidates_kernel");
    return SMG_OK;
}

// ------------------------------------------------------------------------------------------ seed engine (kmer.h)
static smg_pwmset::SeedSet* seed_set(smg_pwmset_t s, const int S)
{
    auto it = s->seeds.find(S);
    if (it != s->seeds.end()) return it->second;
    smg_pwmset::SeedSet* q = new smg_pwmset::SeedSet();
    q->S = S;
    q->eligible = s->split > 0;
    std::vector<float> vtab;
    std::vector<int64_t> voff;
    std::vector<int32_t> vwidth, vcol, voffset, cls;
    std::vector<float> vrest;
    for (int m = 0; m < s->n_motifs; ++m) {
        const int W = s->widths[m];
        if (!(s->split > 0 && W >= s->split && W >= s->kmin && W <= SMG_MAX_REGTAB_WIDTH)) continue;  // the tensor-core side
        if (W <= S || W > S + smg::kSeedMaxRest) { q->eligible = false; continue; }
        const float* T = s->h_nat.data() + s->offs[m] * 4;
        for (int st = 0; st < 2; ++st) {
            // strand table on the forward string: T[k][b] or Trc[k][b] = T[W-1-k][3-b]
            auto at = [&](int k, int b) { return st ? T[(W - 1 - k) * 4 + (3 - b)] : T[k * 4 + b]; };
            std::vector<double> mx(W), mu(W), var(W), amx(W);
            double sum_mx = 0.0, sum_abs = 0.0, thr_proxy = 0.0;
            for (int k = 0; k < W; ++k) {
                double a = -1e300, m1 = 0.0, m2 = 0.0, ab = 0.0;
                for (int b = 0; b < 4; ++b) {
                    const double v = std::isfinite(at(k, b)) ? (double)at(k, b) : -60000.0;
                    a = std::max(a, v); m1 += v; m2 += v * v; ab = std::max(ab, std::fabs(v));
                }
                thr_proxy += a;
                mx[k] = std::max(a, 0.0);  // 'Z' / padding contribute 0: the rest is bounded by max(0, max_b T)
                mu[k] = m1 / 4; var[k] = std::max(m2 / 4 - mu[k] * mu[k], 1e-12); amx[k] = ab;
                sum_mx += mx[k]; sum_abs += ab;
            }
            // the most selective S consecutive positions: largest z-score of the seed threshold (thr ~ 0.7 x max score)
            int best_o = 0;
            double best_z = -1e300;
            for (int o = 0; o + S <= W; ++o) {
                double smx = 0.0, smu = 0.0, sv = 0.0;
                for (int k = o; k < o + S; ++k) { smx += mx[k]; smu += mu[k]; sv += var[k]; }
                const double z = (0.7 * thr_proxy - (sum_mx - smx) - smu) / std::sqrt(sv);
                if (z > best_z) { best_z = z; best_o = o; }
            }
            double rest = 0.0;
            for (int k = 0; k < W; ++k)
                if (k < best_o || k >= best_o + S) rest += mx[k];
            voff.push_back((int64_t)vcol.size() * S);
            cls.push_back((int32_t)vcol.size());
            vcol.push_back(m * 2 + st);
            vwidth.push_back(S);
            voffset.push_back(best_o);
            // margin: fp32 rounding of both sums (<< 2^-16 x the score scale) + a constant
            vrest.push_back((float)(rest + 1e-3 + sum_abs * (1.0 / 65536.0)) * 1.000001f);
            for (int k = best_o; k < best_o + S; ++k)
                for (int b = 0; b < 4; ++b) vtab.push_back(at(k, b));
            q->o_max = std::max(q->o_max, best_o);
        }
    }
    q->nv = (int)vcol.size();
    q->widths.assign(q->nv, S);
    if (q->nv == 0) q->eligible = false;
    std::vector<float> vthr(std::max(q->nv, 1), 0.f);
    std::vector<double> vthr64(std::max(q->nv, 1), 0.0);
    if (upload(&q->d_vtab, vtab) || upload(&q->d_voff, voff) || upload(&q->d_vwidth, vwidth) || upload(&q->d_vcol, vcol) ||
        upload(&q->d_voffset, voffset) || upload(&q->d_vrest, vrest) || upload(&q->d_vthr, vthr) || upload(&q->d_vthr64, vthr64) ||
        upload(&q->d_cls, cls)) {
        delete q;
        return nullptr;
    }
    s->seeds[S] = q;
    return q;
}

// seed thresholds of a call: thr_lo of the real PWM minus the bound of the rest of the window; +inf switches a strand off
__global__ void seed_thr_kernel(const int32_t* __restrict__ vcol, const float* __restrict__ vrest, const float* __restrict__ thr_lo,
                                const int strands, const int nv, float* __restrict__ vthr)
{
    const int v = blockIdx.x * blockDim.x + threadIdx.x;
    if (v >= nv) return;
    const int col = vcol[v];
    vthr[v] = ((strands >> (col & 1)) & 1) ? thr_lo[col >> 1] - vrest[v] : __int_as_float(0x7f800000);
}

int smg_seed_prepare(smg_pwmset_t s, int seed_len, int64_t info[4])
{
    if (!s || !info || seed_len < 4 || seed_len > SMG_KMER_MAX_WIDTH) return fail(SMG_ERR_INVALID, "smg_seed_prepare: bad argument");
    smg_pwmset::SeedSet* q = seed_set(s, seed_len);
    if (!q) return fail(SMG_ERR_CUDA, "smg_seed_prepare: out of memory");
    info[0] = q->nv;
    info[1] = q->nv ? smg_kmer_head_len(q->widths, seed_len) : 0;
    info[2] = q->o_max;
    info[3] = q->eligible ? 1 : 0;
    return SMG_OK;
}

static int seed_kmer_params(smg::KmerParams& kp, smg_pwmset::SeedSet* q, int32_t* d_head, smg_kmer_entry_t* d_entries,
                            uint64_t entry_cap, uint64_t* d_counters)
{
    memset(&kp, 0, sizeof(kp));
    kp.sp.n_motifs = q->nv; kp.sp.motif_width = q->d_vwidth; kp.sp.motif_off = q->d_voff; kp.sp.nat_tab = q->d_vtab;
    kp.sp.strands = 1; kp.sp.thr_lo = q->d_vthr; kp.sp.thr_hi = q->d_vthr;  // forward-only virtual PWMs, no near band
    kp.sp.counters = reinterpret_cast<unsigned long long*>(d_counters);
    if (smg_kmer_fill_params(kp, q->widths, q->d_cls, q->S)) return 1;
    kp.head = d_head; kp.entries = d_entries; kp.entry_cap = entry_cap; kp.thr64 = q->d_vthr64;
    return 0;
}

int smg_seed_index_build(smg_pwmset_t s, int seed_len, int strands, const float* d_thr_lo, int32_t* d_head, int64_t head_len,
                         smg_kmer_entry_t* d_entries, uint64_t entry_cap, uint64_t* d_counters, void* stream)
{
    if (!s || !d_thr_lo || !d_head || !d_entries || entry_cap == 0 || entry_cap > 0x7fffffffull || !d_counters)
        return fail(SMG_ERR_INVALID, "smg_seed_index_build: bad argument");
    if ((strands & 3) == 0 || (strands & ~3)) return fail(SMG_ERR_INVALID, "smg_seed_index_build: strands must be 1, 2 or 3");
    smg_pwmset::SeedSet* q = seed_set(s, seed_len);
    if (!q || q->nv == 0) return fail(SMG_ERR_INVALID, "smg_seed_index_build: no seeded PWMs (smg_seed_prepare)");
    if (head_len != smg_kmer_head_len(q->widths, seed_len)) return fail(SMG_ERR_INVALID, "smg_seed_index_build: head_len mismatch");
    cudaStream_t st = (cudaStream_t)stream;
    seed_thr_kernel<<<(q->nv + 255) / 256, 256, 0, st>>>(q->d_vcol, q->d_vrest, d_thr_lo, strands, q->nv, q->d_vthr);
    SMG_LAUNCH_CHECK("seed_thr_kernel");
    smg::KmerParams kp;
    if (seed_kmer_params(kp, q, d_head, d_entries, entry_cap, d_counters)) return fail(SMG_ERR_INVALID, "smg_seed_index_build: bad seed length");
    int64_t n_heads = 0;
    smg_kmer_head_len(q->widths, seed_len, &n_heads);
    SMG_CUDA(cudaMemsetAsync(d_head, 0xff, (size_t)n_heads * sizeof(int32_t), st));
    SMG_CUDA(cudaMemsetAsync(d_head + n_heads, 0, (size_t)(head_len - n_heads) * sizeof(int32_t), st));
    uint64_t nl = 0;
    cudaError_t e = smg_kmer_launch_build(kp, st, &nl);
    g_launches += nl;
    if (e != cudaSuccess) return fail(SMG_ERR_CUDA, std::string("kmer_build_kernel (seeds): ") + cudaGetErrorString(e));
    return SMG_OK;
}

int smg_scan_seed(smg_pwmset_t s, int seed_len, const uint32_t* d_packed, const uint16_t* d_amask, const uint8_t* d_codes,
                  const smg_row_t* d_rows, int32_t n_rows, const int32_t* d_out_rows, const smg_chunk_t* d_chunks, int32_t n_chunks,
                  int32_t chunk_len, int strands, const float* d_thr_lo, const float* d_thr_hi, const double* d_thr64, int pos_bits,
                  int motif_bits, int cand_pos_bits, const int32_t* d_head, const smg_kmer_entry_t* d_entries, uint64_t* d_cand,
                  uint64_t cand_cap, uint64_t* d_hit_keys, float* d_hit_scores, uint64_t hit_cap, int32_t* d_counts,
                  uint64_t* d_counters, void* stream)
{
    if (!s || !d_packed || !d_amask || !d_rows || n_rows <= 0 || !d_out_rows || !d_chunks || n_chunks < 0 || !d_thr_lo ||
        !d_thr_hi || !d_thr64 || !d_counters || !d_cand || cand_cap == 0 || !d_head || !d_entries)
        return fail(SMG_ERR_INVALID, "smg_scan_seed: bad argument");
    if (chunk_len <= 0 || (chunk_len & 15)) return fail(SMG_ERR_INVALID, "smg_scan_seed: chunk_len must be a positive multiple of 16");
    if ((strands & 3) == 0 || (strands & ~3)) return fail(SMG_ERR_INVALID, "smg_scan_seed: strands must be 1, 2 or 3");
    if (pos_bits <= 0 || motif_bits <= 0 || pos_bits + motif_bits > 63) return fail(SMG_ERR_INVALID, "smg_scan_seed: bad key layout");
    if ((1ll << motif_bits) < s->n_motifs) return fail(SMG_ERR_INVALID, "smg_scan_seed: motif_bits too small");
    if (cand_pos_bits <= 0 || cand_pos_bits + motif_bits + 1 > 62) return fail(SMG_ERR_INVALID, "smg_scan_seed: bad candidate layout");
    if ((d_hit_keys == nullptr) != (d_hit_scores == nullptr)) return fail(SMG_ERR_INVALID, "smg_scan_seed: hit buffers must both be set or NULL");
    smg_pwmset::SeedSet* q = seed_set(s, seed_len);
    if (!q || q->nv == 0 || !q->eligible) return fail(SMG_ERR_INVALID, "smg_scan_seed: the PWM set is not eligible for seeding at this length");
    if (n_chunks == 0) return SMG_OK;
    cudaStream_t st = (cudaStream_t)stream;
    ScanParams p;
    fill_params(p, s, d_packed, d_amask, d_codes, d_rows, d_out_rows);
    p.chunks = d_chunks; p.chunk_len = chunk_len; p.strands = strands; p.thr_lo = d_thr_lo; p.thr_hi = d_thr_hi;
    p.pos_bits = pos_bits; p.motif_bits = motif_bits;
    p.hit_keys = reinterpret_cast<unsigned long long*>(d_hit_keys); p.hit_scores = d_hit_scores; p.hit_cap = hit_cap;
    p.counts = d_counts; p.counters = reinterpret_cast<unsigned long long*>(d_counters);
    smg::KmerParams kp;
    if (seed_kmer_params(kp, q, const_cast<int32_t*>(d_head), const_cast<smg_kmer_entry_t*>(d_entries), 1, d_counters))
        return fail(SMG_ERR_INVALID, "smg_scan_seed: bad seed length");
    smg::SeedParams sp;
    memset(&sp, 0, sizeof(sp));
    sp.sp = p; sp.head = d_head; sp.head_off = kp.head_off[q->S]; sp.bm_off = kp.bm_off[q->S]; sp.entries = d_entries;
    sp.vcol = q->d_vcol; sp.voffset = q->d_voffset; sp.vtab = q->d_vtab; sp.vthr = q->d_vthr;
    sp.nv = q->nv; sp.S = q->S; sp.o_max = q->o_max; sp.n_chunks = n_chunks;
    sp.cand_shift = cand_pos_bits + motif_bits + 1;
    sp.cand = reinterpret_cast<unsigned long long*>(d_cand); sp.cand_cap = cand_cap;
    cudaError_t e = smg_seed_launch_scan(sp, st);
    if (e != cudaSuccess) return fail(SMG_ERR_CUDA, std::string("seed_scan_kernel: ") + cudaGetErrorString(e));
    ++g_launches;
    finalize_candidates_kernel<<<148 * SMG_FINALIZE_MIN_BLOCKS * 2, 256, 0, st>>>(p, reinterpret_cast<const unsigned long long*>(d_cand), cand_cap,
                                                       sp.cand_shift, d_thr64);
    SMG_LAUNCH_CHECK("finalize_candidates_kernel");
    return SMG_OK;
}

int smg_kmer_index_size(smg_pwmset_t s, int kmer_max_width, int64_t* head_len)
{
    if (!s || !head_len || kmer_max_width < 1 || kmer_max_width > SMG_KMER_MAX_WIDTH)
        return fail(SMG_ERR_INVALID, "smg_kmer_index_size: bad argument");
    *head_len = smg_kmer_head_len(s->widths, kmer_max_width);
    return SMG_OK;
}

int smg_kmer_index_build(smg_pwmset_t s, int kmer_max_width, int strands, const float* d_thr_lo, const float* d_thr_hi,
                         const double* d_thr64, int32_t* d_head, int64_t head_len, smg_kmer_entry_t* d_entries,
                         uint64_t entry_cap, uint64_t* d_counters, void* stream)
{
    if (!s || kmer_max_width < 1 || kmer_max_width > SMG_KMER_MAX_WIDTH || !d_thr_lo || !d_thr_hi || !d_thr64 || !d_head ||
        !d_entries || entry_cap == 0 || entry_cap > 0x7fffffffull || !d_counters)
        return fail(SMG_ERR_INVALID, "smg_kmer_index_build: bad argument");
    if ((strands & 3) == 0 || (strands & ~3)) return fail(SMG_ERR_INVALID, "smg_kmer_index_build: strands must be 1, 2 or 3");
    if (head_len != smg_kmer_head_len(s->widths, kmer_max_width)) return fail(SMG_ERR_INVALID, "smg_kmer_index_build: head_len mismatch");
    if (head_len == 0) return SMG_OK;
    cudaStream_t st = (cudaStream_t)stream;
    smg::KmerParams kp;
    memset(&kp, 0, sizeof(kp));
    fill_params(kp.sp, s, nullptr, nullptr, nullptr, nullptr, nullptr);
    kp.sp.strands = strands; kp.sp.thr_lo = d_thr_lo; kp.sp.thr_hi = d_thr_hi;
    kp.sp.counters = reinterpret_cast<unsigned long long*>(d_counters);
    if (smg_kmer_fill_params(kp, s->widths, s->d_cls_motifs, kmer_max_width)) return fail(SMG_ERR_INVALID, "smg_kmer_index_build: bad width");
    kp.head = d_head; kp.entries = d_entries; kp.entry_cap = entry_cap; kp.thr64 = d_thr64;
    int64_t n_heads = 0;
    smg_kmer_head_len(s->widths, kmer_max_width, &n_heads);
    SMG_CUDA(cudaMemsetAsync(d_head, 0xff, (size_t)n_heads * sizeof(int32_t), st));                         // empty lists
    SMG_CUDA(cudaMemsetAsync(d_head + n_heads, 0, (size_t)(head_len - n_heads) * sizeof(int32_t), st));   // bitmaps
    uint64_t nl = 0;
    cudaError_t e = smg_kmer_launch_build(kp, st, &nl);
    g_launches += nl;
    if (e != cudaSuccess) return fail(SMG_ERR_CUDA, std::string("kmer_build_kernel: ") + cudaGetErrorString(e));
    return SMG_OK;
}

int smg_scan_kmer(smg_pwmset_t s, int kmer_max_width, const uint32_t* d_packed, const uint16_t* d_amask, const uint8_t* d_codes,
                  const smg_row_t* d_rows, int32_t n_rows, const int32_t* d_out_rows, const smg_chunk_t* d_chunks,
                  int32_t n_chunks, int32_t chunk_len, int strands, const float* d_thr_lo, const float* d_thr_hi,
                  const double* d_thr64, int pos_bits, int motif_bits, const int32_t* d_head, const smg_kmer_entry_t* d_entries,
                  uint64_t entry_cap, uint64_t* d_hit_keys, float* d_hit_scores, uint64_t hit_cap, int32_t* d_counts,
                  uint64_t* d_counters, void* stream)
{
    if (!s || kmer_max_width < 1 || kmer_max_width > SMG_KMER_MAX_WIDTH || !d_packed || !d_amask || !d_rows || n_rows <= 0 ||
        !d_out_rows || !d_chunks || n_chunks < 0 || !d_thr_lo || !d_thr_hi || !d_thr64 || !d_head || !d_entries || !d_counters)
        return fail(SMG_ERR_INVALID, "smg_scan_kmer: bad argument");
    if (chunk_len <= 0 || (chunk_len & 15)) return fail(SMG_ERR_INVALID, "smg_scan_kmer: chunk_len must be a positive multiple of 16");
    if ((strands & 3) == 0 || (strands & ~3)) return fail(SMG_ERR_INVALID, "smg_scan_kmer: strands must be 1, 2 or 3");
    if (pos_bits <= 0 || motif_bits <= 0 || pos_bits + motif_bits > 63) return fail(SMG_ERR_INVALID, "smg_scan_kmer: bad key layout");
    if ((1ll << motif_bits) < s->n_motifs) return fail(SMG_ERR_INVALID, "smg_scan_kmer: motif_bits too small");
    if ((d_hit_keys == nullptr) != (d_hit_scores == nullptr)) return fail(SMG_ERR_INVALID, "smg_scan_kmer: hit buffers must both be set or NULL");
    if (n_chunks == 0 || smg_kmer_head_len(s->widths, kmer_max_width) == 0) return SMG_OK;
    smg::KmerParams kp;
    memset(&kp, 0, sizeof(kp));
    fill_params(kp.sp, s, d_packed, d_amask, d_codes, d_rows, d_out_rows);
    ScanParams& p = kp.sp;
    p.chunks = d_chunks; p.chunk_len = chunk_len; p.strands = strands; p.thr_lo = d_thr_lo; p.thr_hi = d_thr_hi;
    p.pos_bits = pos_bits; p.motif_bits = motif_bits;
    p.hit_keys = reinterpret_cast<unsigned long long*>(d_hit_keys); p.hit_scores = d_hit_scores; p.hit_cap = hit_cap;
    p.counts = d_counts; p.counters = reinterpret_cast<unsigned long long*>(d_counters);
    if (smg_kmer_fill_params(kp, s->widths, s->d_cls_motifs, kmer_max_width)) return fail(SMG_ERR_INVALID, "smg_scan_kmer: bad width");
    kp.head = const_cast<int32_t*>(d_head); kp.entries = const_cast<smg_kmer_entry_t*>(d_entries); kp.entry_cap = entry_cap;
    kp.thr64 = d_thr64;
    cudaError_t e = smg_kmer_launch_scan(kp, n_chunks, (cudaStream_t)stream);
    if (e != cudaSuccess) return fail(SMG_ERR_CUDA, std::string("scan_kmer_kernel: ") + cudaGetErrorString(e));
    ++g_launches;
    return SMG_OK;
}

int smg_adjudicate(smg_pwmset_t s, const uint32_t* d_packed, const uint16_t* d_amask, const uint8_t* d_codes,
                   const smg_row_t* d_rows, const int32_t* d_out_rows, const double* d_thr64, const smg_near_t* d_near,
                   uint64_t near_cap, int pos_bits, int motif_bits, uint64_t* d_hit_keys, float* d_hit_scores, uint64_t hit_cap,
                   int32_t* d_counts, uint64_t* d_counters, void* stream)
{
    if (!s || !d_packed || !d_amask || !d_rows || !d_out_rows || !d_thr64 || !d_near || !d_counters)
        return fail(SMG_ERR_INVALID, "smg_adjudicate: bad argument");
    ScanParams p;
    fill_params(p, s, d_packed, d_amask, d_codes, d_rows, d_out_rows);
    p.pos_bits = pos_bits; p.motif_bits = motif_bits;
    p.hit_keys = reinterpret_cast<unsigned long long*>(d_hit_keys); p.hit_scores = d_hit_scores; p.hit_cap = hit_cap;
    p.near = const_cast<smg_near_t*>(d_near); p.near_cap = near_cap; p.counts = d_counts;
    p.counters = reinterpret_cast<unsigned long long*>(d_counters);
    const uint64_t want = std::max<uint64_t>(1, std::min<uint64_t>(near_cap, 1u << 20));
    adjudicate_kernel<<<(unsigned)((want + 127) / 128), 128, 0, (cudaStream_t)stream>>>(p, d_thr64);
    SMG_LAUNCH_CHECK("adjudicate_kernel");
    return SMG_OK;
}

int smg_sort_hits(const uint64_t* d_keys_in, const float* d_scores_in, uint64_t* d_keys_out, float* d_scores_out, uint64_t n,
                  int key_bits, void* d_temp, uint64_t* temp_bytes, void* stream)
{
    if (!temp_bytes || key_bits <= 0 || key_bits > 64) return fail(SMG_ERR_INVALID, "smg_sort_hits: bad argument");
    size_t tb = d_temp ? (size_t)*temp_bytes : 0;
    cudaError_t e = cub::DeviceRadixSort::SortPairs(d_temp, tb, reinterpret_cast<const unsigned long long*>(d_keys_in),
                                                    reinterpret_cast<unsigned long long*>(d_keys_out), d_scores_in, d_scores_out,
                                                    (size_t)n, 0, key_bits, (cudaStream_t)stream);
    if (e != cudaSuccess) return fail(SMG_ERR_CUDA, std::string("cub::DeviceRadixSort: ") + cudaGetErrorString(e));
    if (!d_temp) *temp_bytes = (uint64_t)tb; else ++g_launches;
    return SMG_OK;
}

int smg_bin_hits(const uint64_t* d_keys, const float* d_scores, uint64_t n_hits, int pos_bits, int motif_bits,
                 const double* d_max_scores, const double* d_edges, int32_t n_bins, int32_t n_motifs, int32_t* d_hist, void* stream)
{
    if (!d_max_scores || !d_edges || n_bins <= 0 || n_motifs <= 0 || !d_hist || pos_bits <= 0 || motif_bits <= 0 ||
        pos_bits + motif_bits > 63)
        return fail(SMG_ERR_INVALID, "smg_bin_hits: bad argument");
    if (n_hits == 0) return SMG_OK;
    if (!d_keys || !d_scores) return fail(SMG_ERR_INVALID, "smg_bin_hits: hit list missing");
    const unsigned blocks = (unsigned)std::min<uint64_t>((n_hits + 255) / 256, 148 * 16);
    bin_hits_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>(reinterpret_cast<const unsigned long long*>(d_keys), d_scores, n_hits,
                                                              pos_bits, motif_bits, d_max_scores, d_edges, n_bins, d_hist, n_motifs);
    SMG_LAUNCH_CHECK("bin_hits_kernel");
    return SMG_OK;
}

int smg_predict_dense(smg_pwmset_t s, const uint32_t* d_packed, const uint16_t* d_amask, const uint8_t* d_codes,
                      const smg_row_t* d_rows, int32_t n_rows, int32_t row_len, float* d_out, void* stream)
{
    if (!s || !d_packed || !d_amask || !d_rows || n_rows <= 0 || row_len <= 0 || !d_out)
        return fail(SMG_ERR_INVALID, "smg_predict_dense: bad argument");
    ScanParams p;
    fill_params(p, s, d_packed, d_amask, d_codes, d_rows, nullptr);
    const int64_t total = (int64_t)n_rows * row_len * s->n_motifs;
    const int64_t blocks = std::min<int64_t>((total + 255) / 256, 148 * 64);
    dense_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(p, n_rows, row_len, d_out);
    SMG_LAUNCH_CHECK("dense_kernel");
    return SMG_OK;
}

int smg_variant_sites(smg_pwmset_t s, const uint32_t* d_packed, const uint16_t* d_amask, const uint8_t* d_codes,
                      const smg_row_t* d_rows, const int32_t* d_site_row, const int32_t* d_site_start, const int32_t* d_site_motif,
                      const int32_t* d_rel_pos, const uint8_t* d_alt_code, int64_t n, double* d_out_ref, double* d_out_alt,
                      void* stream)
{
    if (!s || !d_packed || !d_amask || !d_rows || !d_site_row || !d_site_start || !d_site_motif || !d_rel_pos || !d_alt_code ||
        n < 0 || !d_out_ref || !d_out_alt)
        return fail(SMG_ERR_INVALID, "smg_variant_sites: bad argument");
    if (n == 0) return SMG_OK;
    ScanParams p;
    fill_params(p, s, d_packed, d_amask, d_codes, d_rows, nullptr);
    const int64_t blocks = std::min<int64_t>((n + 127) / 128, 148 * 32);
    variant_sites_kernel<<<(unsigned)blocks, 128, 0, (cudaStream_t)stream>>>(p, d_site_row, d_site_start, d_site_motif, d_rel_pos,
                                                                            d_alt_code, n, d_out_ref, d_out_alt);
    SMG_LAUNCH_CHECK("variant_sites_kernel");
    return SMG_OK;
}

int smg_variant_windows(smg_pwmset_t s, const uint32_t* d_packed, const uint16_t* d_amask, const uint8_t* d_codes,
                        const smg_row_t* d_rows, const int32_t* d_snv_row, const int32_t* d_snv_pos, const uint8_t* d_snv_alt,
                        int64_t n_snv, float* d_ref_max, float* d_alt_max, void* stream)
{
    return smg_variant_windows_mode(s, d_packed, d_amask, d_codes, d_rows, d_snv_row, d_snv_pos, d_snv_alt, n_snv, d_ref_max,
                                    d_alt_max, 1, 0, stream);
}

int smg_variant_windows_mode(smg_pwmset_t s, const uint32_t* d_packed, const uint16_t* d_amask, const uint8_t* d_codes,
                             const smg_row_t* d_rows, const int32_t* d_snv_row, const int32_t* d_snv_pos,
                             const uint8_t* d_snv_alt, int64_t n_snv, float* d_ref_max, float* d_alt_max, int32_t exact,
                             int32_t motif_major, void* stream)
{
    if (!s || !d_packed || !d_amask || !d_rows || !d_snv_row || !d_snv_pos || !d_snv_alt || n_snv < 0 || !d_ref_max || !d_alt_max)
        return fail(SMG_ERR_INVALID, "smg_variant_windows: bad argument");
    if (n_snv == 0) return SMG_OK;
    if (s->kmin > 1) return fail(SMG_ERR_INVALID, "smg_variant_windows: needs a PWM set built without a k-mer split");
    const int64_t snv_stride = motif_major ? 1 : (int64_t)s->n_motifs, motif_stride = motif_major ? n_snv : 1;
    ScanParams p;
    fill_params(p, s, d_packed, d_amask, d_codes, d_rows, nullptr);
    cudaStream_t st = (cudaStream_t)stream;
    const unsigned nb = (unsigned)std::min<int64_t>((n_snv + 3) / 4, 148 * 64);
    // PWMs up to kDualMaxWidth: register-table kernel (both strands per lane); SNVs it cannot take (ambiguity codes
    // in the region, ambiguous alt) are flagged and redone by the generic kernel, as are the wider PWMs.
    unsigned int* d_flags = nullptr;
    if (!s->dual_motifs.empty()) {
        // [n_snv] flags | [n_snv] list of flagged SNVs | count
        SMG_CUDA(cudaMallocAsync((void**)&d_flags, (size_t)(2 * n_snv + 1) * sizeof(unsigned int), st));
        SMG_CUDA(cudaMemsetAsync(d_flags, 0, (size_t)(2 * n_snv + 1) * sizeof(unsigned int), st));
        smg::VariantParams vp;
        vp.sp = p; vp.snv_row = d_snv_row; vp.snv_pos = d_snv_pos; vp.snv_alt = d_snv_alt; vp.n_snv = n_snv;
        vp.ref_max = d_ref_max; vp.alt_max = d_alt_max; vp.slow_flags = d_flags;
        vp.slow_list = d_flags + n_snv; vp.slow_count = d_flags + 2 * n_snv;
        vp.exact = (exact || !s->all_finite) ? 1 : 0;
        vp.snv_stride = snv_stride; vp.motif_stride = motif_stride;
        const int blocks = (int)std::min<int64_t>((n_snv + 31) / 32, 148 * 16 * 4);  // a CTA (one warp) takes 32 SNVs at a time
        for (const Bucket& b : s->buckets) {
            smg_variant_launch_fn fn = smg_get_variant_launcher(b.wb);
            if (!fn) return fail(SMG_ERR_INVALID, "smg_variant_windows: no kernel for padded width " + std::to_string(b.wb));
            vp.sp.first_group = b.first_group;
            fn(vp, blocks, b.n_groups, b.nm, st);
            SMG_LAUNCH_CHECK("variant_regtab_kernel");
        }
        dim3 gridf(nb, (unsigned)((s->dual_motifs.size() + 31) / 32), 1);
        variant_windows_kernel<<<gridf, 128, 0, st>>>(p, d_snv_row, d_snv_pos, d_snv_alt, n_snv, d_ref_max, d_alt_max,
                                                       s->d_dual_list, (int)s->dual_motifs.size(), d_flags + n_snv, d_flags + 2 * n_snv,
                                                       snv_stride, motif_stride);
        SMG_LAUNCH_CHECK("variant_windows_kernel(flagged)");
        SMG_CUDA(cudaFreeAsync(d_flags, st));
    }
    if (!s->nondual_motifs.empty()) {
        dim3 grid(nb, (unsigned)((s->nondual_motifs.size() + 31) / 32), 1);
        variant_windows_kernel<<<grid, 128, 0, st>>>(p, d_snv_row, d_snv_pos, d_snv_alt, n_snv, d_ref_max, d_alt_max,
                                                      s->d_nondual_list, (int)s->nondual_motifs.size(), nullptr, nullptr, snv_stride,
                                                      motif_stride);
        SMG_LAUNCH_CHECK("variant_windows_kernel");
    }
    return SMG_OK;
}
}  // extern "C"
