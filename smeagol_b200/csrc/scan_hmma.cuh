// Tensor-core PREFILTER for the PWM scan (the "im2col" variant of north_star, restated so that it stays exact).
//
// The window score is a contraction of a one-hot operand with the log-odds table:
//     score[col][p] = sum_{k < W, b < 4} T_col[k][b] * X[p + k][b]            (models.py:50-68: Embedding + Conv1D)
// i.e. a GEMM with M = (PWM, strand) columns, N = positions, K = 4 * W.  Here the WEIGHTS are the A operand and stay in
// registers as fp16 fragments, the soft one-hot of the sequence is the B operand and is generated on the fly in
// registers from the 2-bit packed bases (no shared-memory staging, no im2col buffer), and the accumulator fragment is
// compared against the threshold IN REGISTERS -- which is why the legacy warp-level mma.sync is the right instruction:
// a tcgen05 version has to read every accumulator back from TMEM (64 B/clk/SM = 16 cells/clk/SM), below what mma.sync
// sustains for this K (951 MAC/clk/SM measured => 20-30 cells/clk/SM for W <= 12).
//
// fp16 weights cannot give the reference's fp32 scores, so this kernel is only a FILTER: it emits every cell whose
// approximate score exceeds (threshold - margin), margin = 2^-9 * sum_k max_b |T[k][b]| (>= 4x the fp16 rounding error
// of the operands), as a 64-bit candidate (row, start, motif, strand): per-lane queues of 32-bit tags in shared memory
// (predicated store + pointer bump in the loop), drained with ballot compaction.  finalize_candidates_kernel (api.cu) then
// re-scores each candidate in fp32 in the SAME fixed order as the register-table kernel, adjudicates the near-threshold
// band in fp64, and produces hits / counts that are bit-identical to the register-table path.
#pragma once
#include "common.cuh"

namespace smg {

struct HmmaParams {
    ScanParams sp;
    const uint32_t* afrag;        // fp16x2 A fragments, per column group: [NB][KS][4][32]
    const int64_t* cg_afrag_off;  // [n_colgroups] offset (in uint32) of each column group's fragments
    const float* thr_pre_col;     // [n_colgroups * NB * 16] prefilter threshold per column (+inf for empty columns)
    const int32_t* col_id;        // [n_colgroups * NB * 16] motif * 2 + strand, -1 for empty columns
    int32_t first_colgroup;
    int32_t cand_shift;           // candidate = row << cand_shift | start << (motif_bits + 1) | motif << 1 | strand
    unsigned long long* cand;
    unsigned long long cand_cap;
};

__device__ __forceinline__ void mma_f16(float (&c)[4], const uint32_t (&a)[4], const uint32_t b0, const uint32_t b1)
{
    asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                 : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
                 : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}

// first k16 slice of a tile: accumulate onto zero (the zero operand is RZ, no accumulator initialisation needed)
__device__ __forceinline__ void mma_f16_zero(float (&c)[4], const uint32_t (&a)[4], const uint32_t b0, const uint32_t b1)
{
    asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%10,%10,%10,%10};"
                 : "=f"(c[0]), "=f"(c[1]), "=f"(c[2]), "=f"(c[3])
                 : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1), "f"(0.f));
}

// first k16 slice of a tile with an explicit C operand: rows lane/4 (c0, c1) and lane/4 + 8 (c2, c3) of the block
__device__ __forceinline__ void mma_f16_init(float (&c)[4], const uint32_t (&a)[4], const uint32_t b0, const uint32_t b1,
                                             const float ca, const float cb)
{
    asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%10,%10,%11,%11};"
                 : "=f"(c[0]), "=f"(c[1]), "=f"(c[2]), "=f"(c[3])
                 : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1), "f"(ca), "f"(cb));
}

__device__ __forceinline__ float vmax4(const float (&c)[4])
{
    float d;
    asm("max.f32 %0, %1, %2, %3;" : "=f"(d) : "f"(c[0]), "f"(c[1]), "f"(c[2]));
    return fmaxf(d, c[3]);
}

// 0x3c00 (fp16 1.0) shifted left by s; PTX shl clamps shift amounts above 31, i.e. s >= 32 gives 0
__device__ __forceinline__ uint32_t shl_one_f16(const uint32_t s)
{
    uint32_t d;
    asm("shl.b32 %0, 0x3c00, %1;" : "=r"(d) : "r"(s));
    return d;
}

// fp16 pairs {e[c][bb], e[c][bb+1]} of the soft one-hot table (encode.py:11-29) for bb = 0 and bb = 2
static __constant__ uint32_t c_emb16[17][2] = {
    {0x00000000u, 0x00000000u},  // Z
    {0x00003c00u, 0x00000000u},  // A
    {0x3c000000u, 0x00000000u},  // C
    {0x00000000u, 0x00003c00u},  // G
    {0x00000000u, 0x3c000000u},  // T
    {0x00000000u, 0x3c000000u},  // U
    {0x34003400u, 0x34003400u},  // N  (0.25)
    {0x00003800u, 0x38000000u},  // W  (A, T)
    {0x38000000u, 0x00003800u},  // S  (C, G)
    {0x38003800u, 0x00000000u},  // M  (A, C)
    {0x00000000u, 0x38003800u},  // K  (G, T)
    {0x00003800u, 0x00003800u},  // R  (A, G)
    {0x38000000u, 0x38000000u},  // Y  (C, T)
    {0x35550000u, 0x35553555u},  // B  (C, G, T) 1/3 = 0x3555
    {0x00003555u, 0x35553555u},  // D  (A, G, T)
    {0x35553555u, 0x35550000u},  // H  (A, C, T)
    {0x35553555u, 0x00003555u},  // V  (A, C, G)
};

// fp16 pair of the soft one-hot of base q of a row; kept out of line so that the hot loop stays small (instruction cache)
static __device__ __noinline__ uint32_t soft_onehot_f16(const uint32_t* packed, const uint16_t* amask, const uint8_t* codes,
                                                        const long long word_off, const int n_avail, const long long q, const int half)
{
    smg_row_t row;
    row.word_off = word_off; row.n_avail = n_avail; row.g_off = 0; row.g_len = 0; row.n_own = 0;
    return c_emb16[smg_fetch_code(packed, amask, codes, row, q)][half];
}

constexpr int kCandStage = 128;

constexpr int kTcQueue = 32;  // per-lane candidate queue slots (a tile adds at most 4 * NB <= 16 per lane)

// Drain of the per-lane candidate queues (out of line, runs a few times per chunk): every lane decodes its tags
// (tile << 5 | block << 2 | accumulator) into (window start, column), the warp compacts them with ballot +
// prefix-popcount into its stage, and the stage is appended to the global candidate list with one atomic per flush.
// Returns the new stage fill.
static __device__ __noinline__ int drain_queues(const uint32_t* s_q, const int qn, unsigned long long* s_cand, int nst, const int c0,
                                                const int c1, const int* __restrict__ col_id, const unsigned long long rowbits,
                                                const int pos_shift, unsigned long long* counter,
                                                unsigned long long* __restrict__ cand, const unsigned long long cand_cap)
{
    const int lane = threadIdx.x;
    const unsigned lt_mask = (1u << lane) - 1u;
    const int q2 = (lane & 3) * 2;
    const int maxn = __reduce_max_sync(SMG_FULL_MASK, qn);
    for (int j = 0; j < maxn; ++j) {
        const bool have = j < qn;
        const unsigned tag = have ? s_q[j * 32 + lane] : 0u;
        const int i = (int)(tag & 3u), blk = (int)((tag >> 2) & 7u);
        const int s = c0 + (int)(tag >> 5) * 8 + q2 + (i & 1);
        const bool is_cand = have && (s < c1);
        const unsigned m = __ballot_sync(SMG_FULL_MASK, is_cand);
        if (m) {
            if (is_cand) {
                const int cid = col_id[blk * 16 + (lane >> 2) + ((i < 2) ? 0 : 8)];
                s_cand[nst + __popc(m & lt_mask)] =
                    rowbits | ((unsigned long long)(unsigned)s << pos_shift) | (unsigned long long)(unsigned)cid;
            }
            nst += __popc(m);
            __syncwarp();
            if (nst > kCandStage - 32) {
                unsigned long long base = 0;
                if (lane == 0) base = atomicAdd(counter, (unsigned long long)nst);
                base = __shfl_sync(SMG_FULL_MASK, base, 0);
                for (int k = lane; k < nst; k += 32)
                    if (base + k < cand_cap) cand[base + k] = s_cand[k];
                nst = 0;
                __syncwarp();
            }
        }
    }
    return nst;
}

// register caps on the occupancy steps of an SM sub-partition (see scan_regtab.cuh): 80/96/128/168 registers =
// 6/5/4/3 one-warp CTAs
template <int KS, int NB> struct HmmaMaxRegs {
    static constexpr int v = NB > 6 ? 240 : NB > 4 ? (KS <= 2 ? 128 : 168)
                             : NB == 4 ? (KS == 1 ? 80 : KS == 2 ? 96 : KS == 3 ? 128 : 168)
                             : NB == 3 ? (KS == 1 ? 64 : KS == 2 ? 80 : KS == 3 ? 96 : KS <= 5 ? 128 : 168)
                                       : (KS == 1 ? 64 : KS == 2 ? 72 : KS == 3 ? 80 : 128);
};

// KS = k16 steps (padded width = 4 * KS), NB = 16-column blocks per warp.  One warp per CTA; the warp walks its chunk
// in tiles of 8 positions: the B fragments of the next tile are generated (integer pipe) while the tensor pipe works on
// the current one; a tile that passes the vote pushes its candidates to per-lane queues, drained out of line.
template <int KS, int NB>
__global__ void __launch_bounds__(32) __maxnreg__((HmmaMaxRegs<KS, NB>::v)) scan_hmma_kernel(const HmmaParams hp)
{
    const ScanParams& p = hp.sp;
    __shared__ unsigned long long s_cand[kCandStage];
    __shared__ uint32_t s_q[kTcQueue * 32];  // per-lane candidate queues, [slot][lane]
    static_assert(4 * NB <= kTcQueue - 8, "a tile must fit in the queue");
    const int lane = threadIdx.x;
    const int colgroup = hp.first_colgroup + blockIdx.y;
    const smg_chunk_t ch = p.chunks[blockIdx.x];
    const smg_row_t row = p.rows[ch.row];
    const int c0 = ch.start;
    const int c1 = min(c0 + p.chunk_len, row.n_own);
    if (c1 <= c0) return;

    // ---- weights: fp16 A fragments of NB x 16 columns in registers
    uint32_t a[NB][KS][4];
    {
        const uint32_t* ga = hp.afrag + hp.cg_afrag_off[colgroup];
#pragma unroll
        for (int blk = 0; blk < NB; ++blk)
#pragma unroll
            for (int ks = 0; ks < KS; ++ks)
#pragma unroll
                for (int r = 0; r < 4; ++r) a[blk][ks][r] = ga[((blk * KS + ks) * 4 + r) * 32 + lane];
    }
    const int colbase = colgroup * (NB * 16);
    float tA[NB], tB[NB];  // prefilter thresholds of the two columns this lane sees in each block
#pragma unroll
    for (int blk = 0; blk < NB; ++blk) {
        tA[blk] = hp.thr_pre_col[colbase + blk * 16 + (lane >> 2)];
        tB[blk] = hp.thr_pre_col[colbase + blk * 16 + (lane >> 2) + 8];
    }
    const int n = lane >> 2;         // position offset of this lane's B-fragment column within a tile
    const int kk = (lane & 3) >> 1;  // PWM-row offset of its k-pair within a k16 slice (second register: +2)
    const int bb = (lane & 1) * 2;   // its base pair {bb, bb+1}
    const uint32_t* __restrict__ wp = p.packed + row.word_off;
    const uint16_t* __restrict__ ap = p.amask + row.word_off;
    const unsigned long long rowbits = (unsigned long long)(unsigned)ch.row << hp.cand_shift;
    int nst = 0;

    auto flush = [&]() {
        unsigned long long base = 0;
        if (lane == 0) base = atomicAdd(&p.counters[SMG_CTR_CANDIDATES], (unsigned long long)nst);
        base = __shfl_sync(SMG_FULL_MASK, base, 0);
        for (int i = lane; i < nst; i += 32)
            if (base + i < hp.cand_cap) hp.cand[base + i] = s_cand[i];
        nst = 0;
        __syncwarp();
    };

    // Soft one-hot B fragments of the tile of 8 positions starting at pt (this lane: bases pt + n + 4*ks + kk (+2)).
    // A tile starts at a multiple of 8 and n <= 7, so every lane of the warp reads the SAME three packed words
    // w0..w2 (48 bases from the word that holds pt): they are loaded once per 16 positions at a warp-uniform address
    // and the per-lane part is two funnel shifts by sft = 2 * (n + kk) (+16 for the second tile of the word).
    const uint32_t bb16 = (uint32_t)bb << 4;  // 16 * bb
    const int sft0 = 2 * (n + kk);
    auto make_b = [&](const int pt, const uint32_t w0, const uint32_t w1, const uint32_t w2, const unsigned amb, const int sft,
                      uint32_t (&b)[KS][2]) {
        if (amb == 0u) {
            // d = base ^ bb selects the half of the fp16 pair that is 1.0 (d = 0: low, 1: high, 2 and 3: neither):
            // value = 0x3c00 << (16 * d) with the shift clamped, 16 * d taken straight from the window bits:
            // three instructions per fragment register (shift, and-xor, shift).
            const uint32_t klo = __funnelshift_rc(w0, w1, sft), khi = __funnelshift_rc(w1, w2, sft);
#pragma unroll
            for (int ks = 0; ks < KS; ++ks) {
#pragma unroll
                for (int h = 0; h < 2; ++h) {
                    const int bit = 2 * (4 * ks + 2 * h);  // bit offset of this row's base in the shifted window
                    uint32_t x;                            // 2-bit base code moved to bits 4..5 (fields never straddle)
                    if (bit == 0) x = klo << 4;
                    else if (bit < 32) x = klo >> (bit - 4);
                    else if (bit == 32) x = khi << 4;
                    else x = khi >> (bit - 36);
                    b[ks][h] = shl_one_f16((x & 0x30u) ^ bb16);
                }
            }
        } else {  // IUPAC codes / 'Z' / padding somewhere near: soft one-hot from the code table (out of line, rare)
#pragma unroll
            for (int ks = 0; ks < KS; ++ks) {
#pragma unroll
                for (int h = 0; h < 2; ++h)
                    b[ks][h] = soft_onehot_f16(p.packed, p.amask, p.codes, row.word_off, row.n_avail,
                                               (long long)(pt + n) + 4 * ks + kk + 2 * h, bb >> 1);
            }
        }
    };
    // NB x KS tensor-core MMAs of one tile (the first slice accumulates onto RZ: no accumulator initialisation)
    auto issue = [&](float (&c)[NB][4], const uint32_t (&b)[KS][2]) {
#pragma unroll
        for (int ks = 0; ks < KS; ++ks)
#pragma unroll
            for (int blk = 0; blk < NB; ++blk) {
                if (ks == 0) mma_f16_zero(c[blk], a[blk][ks], b[ks][0], b[ks][1]);
                else mma_f16(c[blk], a[blk][ks], b[ks][0], b[ks][1]);
            }
    };
    // one max-reduction + one vote for the whole tile (NB x 16 columns x 8 positions); per block and per accumulator
    // only when something passed.  Candidates are rare: ballot + prefix-popcount compaction into the warp stage.
    // Candidate handling in two levels (as in scan_regtab.cuh): a lane whose accumulator exceeds the prefilter threshold
    // appends a 32-bit tag to ITS OWN queue in shared memory (predicated store + pointer bump: no ballot, shuffle or
    // call in the scan loop); the queues are drained together when one may overflow and at the end of the chunk.
    const unsigned qbase = (unsigned)__cvta_generic_to_shared(&s_q[lane]);
    unsigned qaddr = qbase;
    auto drain = [&]() {
        nst = drain_queues(s_q, (int)((qaddr - qbase) >> 7), s_cand, nst, c0, c1, hp.col_id + colbase, rowbits, p.motif_bits + 1,
                           &p.counters[SMG_CTR_CANDIDATES], hp.cand, hp.cand_cap);
        qaddr = qbase;
        __syncwarp();
    };
    auto threshold = [&](const float (&c)[NB][4], const int pt) {
        // first level: a predicate chain and one vote for the whole tile (NB x 16 columns x 8 positions)
        bool any = false;
#pragma unroll
        for (int blk = 0; blk < NB; ++blk)
            any = any || (fmaxf(c[blk][0], c[blk][1]) > tA[blk]) || (fmaxf(c[blk][2], c[blk][3]) > tB[blk]);
        if (!__any_sync(SMG_FULL_MASK, any)) return;
        const unsigned tagbase = (unsigned)((pt - c0) >> 3) << 5;
#pragma unroll
        for (int blk = 0; blk < NB; ++blk) {
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                if (c[blk][i] > ((i < 2) ? tA[blk] : tB[blk])) {
                    asm volatile("st.shared.b32 [%0], %1;" ::"r"(qaddr), "r"(tagbase + (unsigned)(blk * 4 + i)) : "memory");
                    qaddr += 128u;
                }
            }
        }
        if (__any_sync(SMG_FULL_MASK, qaddr > qbase + (unsigned)(kTcQueue - 4 * NB) * 128u)) drain();
    };

    // c0 is a multiple of 16 (chunks are cut on word boundaries); two tiles per packed word
    int wi = c0 >> 4;
    uint32_t w0 = wp[wi], w1 = wp[wi + 1], w2 = wp[wi + 2];
    unsigned a0 = ap[wi], a1 = ap[wi + 1], a2 = ap[wi + 2];
    uint32_t b[KS][2];
    float c[NB][4];
    make_b(c0, w0, w1, w2, a0 | a1 | a2, sft0, b);
    for (int pt = c0; pt < c1; pt += 16) {
        issue(c, b);
        make_b(pt + 8, w0, w1, w2, a0 | a1 | a2, sft0 + 16, b);  // next tile (ALU) while the tensor pipe drains this one
        threshold(c, pt);
        issue(c, b);
        ++wi;  // rows are padded: reading one word ahead of the chunk is safe
        w0 = w1; w1 = w2; w2 = wp[wi + 2];
        a0 = a1; a1 = a2; a2 = ap[wi + 2];
        make_b(pt + 16, w0, w1, w2, a0 | a1 | a2, sft0, b);
        threshold(c, pt + 8);
    }
    drain();
    if (nst > 0) flush();
}

template <int KS, int NB>
void launch_scan_hmma(const HmmaParams& hp, int n_chunks, int n_colgroups, cudaStream_t stream)
{
    dim3 grid((unsigned)n_chunks, (unsigned)n_colgroups, 1);
    scan_hmma_kernel<KS, NB><<<grid, 32, 0, stream>>>(hp);
}

}  // namespace smg
