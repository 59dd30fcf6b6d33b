// Shared device-side definitions of the smeagol_b200 kernels (sm_100a only).
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

#include "../../include/smeagol_b200.h"

#define SMG_FULL_MASK 0xffffffffu
#define SMG_MAX_PWMS_PER_LANE 2  // a lane of the register-table kernels holds 1 or 2 PWMs (both strands each)

// 17 x 4 soft one-hot table of the reference alphabet Z,A,C,G,T,U,N,W,S,M,K,R,Y,B,D,H,V
// (smeagol/encode.py:11-29), fp32 as the reference's Embedding layer holds it (models.py:89).
#define SMG_T3 0.3333333432674407958984375f
static __constant__ float c_emb[17][4] = {
    {0.f, 0.f, 0.f, 0.f},             // Z
    {1.f, 0.f, 0.f, 0.f},             // A
    {0.f, 1.f, 0.f, 0.f},             // C
    {0.f, 0.f, 1.f, 0.f},             // G
    {0.f, 0.f, 0.f, 1.f},             // T
    {0.f, 0.f, 0.f, 1.f},             // U
    {0.25f, 0.25f, 0.25f, 0.25f},     // N
    {0.5f, 0.f, 0.f, 0.5f},           // W
    {0.f, 0.5f, 0.5f, 0.f},           // S
    {0.5f, 0.5f, 0.f, 0.f},           // M
    {0.f, 0.f, 0.5f, 0.5f},           // K
    {0.5f, 0.f, 0.5f, 0.f},           // R
    {0.f, 0.5f, 0.f, 0.5f},           // Y
    {0.f, SMG_T3, SMG_T3, SMG_T3},    // B
    {SMG_T3, 0.f, SMG_T3, SMG_T3},    // D
    {SMG_T3, SMG_T3, 0.f, SMG_T3},    // H
    {SMG_T3, SMG_T3, SMG_T3, 0.f},    // V
};

// Integer code (0..16) of base `q` (row-local) of a row.  Bases at or beyond n_avail are 'Z' (the reference's
// zero end-padding, models.py:54).
__device__ __forceinline__ int smg_fetch_code(const uint32_t* __restrict__ packed, const uint16_t* __restrict__ amask,
                                              const uint8_t* __restrict__ codes, const smg_row_t& row, int64_t q)
{
    if (q < 0 || q >= row.n_avail) return 0;
    const int64_t w = row.word_off + (q >> 4);
    const int j = (int)(q & 15);
    if ((amask[w] >> j) & 1) return codes ? (int)codes[row.word_off * 16 + q] : 0;
    return 1 + (int)((packed[w] >> (2 * j)) & 3u);
}

// fp32 contribution of integer code `c` at one PWM row (t0..t3 = A,C,G,T weights), in the fixed order
// ((e0*t0 + e1*t1) + e2*t2) + e3*t3 with every product and sum rounded (no FMA contraction).
__device__ __forceinline__ float smg_mix(const float e0, const float e1, const float e2, const float e3, const float t0,
                                         const float t1, const float t2, const float t3)
{
    return __fadd_rn(__fadd_rn(__fadd_rn(__fmul_rn(e0, t0), __fmul_rn(e1, t1)), __fmul_rn(e2, t2)), __fmul_rn(e3, t3));
}

__device__ __forceinline__ double smg_mix64(const int c, const float t0, const float t1, const float t2, const float t3)
{
    const double p0 = __dmul_rn((double)c_emb[c][0], (double)t0);
    const double p1 = __dmul_rn((double)c_emb[c][1], (double)t1);
    const double p2 = __dmul_rn((double)c_emb[c][2], (double)t2);
    const double p3 = __dmul_rn((double)c_emb[c][3], (double)t3);
    return __dadd_rn(__dadd_rn(__dadd_rn(p0, p1), p2), p3);
}

// Everything the fused scan kernels need.  Passed by value (kernel parameter space).
struct ScanParams {
    const uint32_t* packed;
    const uint16_t* amask;
    const uint8_t* codes;
    const smg_row_t* rows;
    const int32_t* out_rows;  // [n_rows][2]
    const smg_chunk_t* chunks;
    int32_t chunk_len;
    int32_t strands;
    const float* thr_lo;
    const float* thr_hi;
    int32_t pos_bits;
    int32_t motif_bits;
    unsigned long long* hit_keys;
    float* hit_scores;
    unsigned long long hit_cap;
    smg_near_t* near;
    unsigned long long near_cap;
    int32_t* counts;  // [n_rows][n_motifs][2]
    unsigned long long* counters;
    int32_t n_motifs;
    // PWM-set data
    const int32_t* motif_width;   // [n_motifs]
    const int64_t* motif_off;     // [n_motifs] row offset into nat_tab
    const float* nat_tab;         // [sum W][4]
    const float* group_tab;       // lane-major register tables of all groups
    const int64_t* group_tab_off; // [n_groups] offset (in floats) of each group's table
    const int32_t* lane_motif;    // [n_groups][SMG_MAX_PWMS_PER_LANE][32], -1 = empty slot
    int32_t first_group;          // first group of the bucket being launched
};

// ---------------------------------------------------------------------------------- exact window scores (shared)
// score of one window on the forward string, strand st, fp32 ascending-k order (same order as the register kernel)
__device__ __forceinline__ float window_score32(const ScanParams& p, const smg_row_t& row, const int motif, const int st, const int64_t s)
{
    const int wm = p.motif_width[motif];
    const float* tab = p.nat_tab + p.motif_off[motif] * 4;
    float acc = 0.f;
    for (int k = 0; k < wm; ++k) {
        const int c = smg_fetch_code(p.packed, p.amask, p.codes, row, s + k);
        const float* tr = tab + (st ? (wm - 1 - k) : k) * 4;
        const float t0 = st ? tr[3] : tr[0], t1 = st ? tr[2] : tr[1], t2 = st ? tr[1] : tr[2], t3 = st ? tr[0] : tr[3];
        const float term = smg_mix(c_emb[c][0], c_emb[c][1], c_emb[c][2], c_emb[c][3], t0, t1, t2, t3);
        acc = (k == 0) ? term : __fadd_rn(acc, term);
    }
    return acc;
}

// Same value as window_score32 (bit for bit) for a window of plain A/C/G/T bases that lies inside the row, W <= 32:
// the bases come from three packed words held in registers, one table load + one add per PWM row.  Returns false when
// the window needs the general path (ambiguity codes, 'Z', end padding, wider PWM).
__device__ __forceinline__ bool window_score32_acgt(const ScanParams& p, const smg_row_t& row, const int motif, const int st,
                                                    const int64_t s, float* out)
{
    const int wm = p.motif_width[motif];
    if (wm > 32 || s < 0 || s + wm > row.n_avail) return false;
    const int64_t w0 = row.word_off + (s >> 4);
    if ((p.amask[w0] | p.amask[w0 + 1] | p.amask[w0 + 2]) != 0) return false;  // rows are padded: w0 + 2 is allocated
    const int sh = 2 * (int)(s & 15);
    const unsigned long long lo = (unsigned long long)p.packed[w0] | ((unsigned long long)p.packed[w0 + 1] << 32);
    const unsigned long long hi = (unsigned long long)p.packed[w0 + 2];
    unsigned long long win = sh ? ((lo >> sh) | (hi << (64 - sh))) : lo;
    const float* tab = p.nat_tab + p.motif_off[motif] * 4;
    const int last = 4 * wm - 1;
    // forward: T[k][b]; reverse complement on the forward string: T[W-1-k][3-b] = flat index (4W-1) - (4k+b).
    // Ascending k, one rounded add per row: the summation order of the register-table kernel's ACGT path.
    float acc = 0.f;
    for (int k = 0; k < wm; ++k) {
        const int j = 4 * k + (int)(win & 3ull);
        win >>= 2;
        const float term = tab[st ? (last - j) : j];
        acc = (k == 0) ? term : __fadd_rn(acc, term);
    }
    *out = acc;
    return true;
}

__device__ __forceinline__ double window_score64(const uint32_t* packed, const uint16_t* amask, const uint8_t* codes, const smg_row_t& row,
                                 const float* tab, const int wm, const int st, const int64_t s, const int64_t sub_pos,
                                 const int sub_code)
{
    double acc = 0.0;
    for (int k = 0; k < wm; ++k) {
        int c = smg_fetch_code(packed, amask, codes, row, s + k);
        if (s + k == sub_pos) c = sub_code;
        const float* tr = tab + (st ? (wm - 1 - k) : k) * 4;
        const float t0 = st ? tr[3] : tr[0], t1 = st ? tr[2] : tr[1], t2 = st ? tr[1] : tr[2], t3 = st ? tr[0] : tr[3];
        const double term = smg_mix64(c, t0, t1, t2, t3);
        acc = (k == 0) ? term : __dadd_rn(acc, term);
    }
    return acc;
}

