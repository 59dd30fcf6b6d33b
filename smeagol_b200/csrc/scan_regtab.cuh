// Fused PWM scan kernel, register-table variant (the hot path).
//
// Replaces PWMModel.predict + np.where threshold + end trim (smeagol/models.py:92-117) and the per-(row, motif,
// strand) counting of scan._count_sites (smeagol/scan.py:114-147) for both strands in ONE pass.
//
// Mapping (B200-first, not a Conv1D): one warp per CTA; every LANE owns NM (1 or 2) PWMs, both strands each, with the
// log-odds tables in REGISTERS.  The sequence is the warp-uniform operand: all lanes consume the same base at the same
// step, so the table row is selected by uniform control flow (bit tests on the packed word) and no per-lane
// indexing, shared-memory lookup or bank conflict exists.  The window sum is a shift register
// r[k] <- r[k-1] + T[k][b]  (ascending-k summation order, bit-identical to a sequential fp32 sum), advanced with
// packed FADD2 (sm_100 fp32x2): one issue slot performs the add for the forward AND the reverse-complement strand
// (RC strand = forward string scanned with Trc[k][b] = T[W-1-k][3-b]).  Completed window scores are checked against
// the per-PWM threshold every 4 steps; candidates go to per-lane queues in shared memory and are drained with warp
// ballot + prefix popcount into a per-warp stage that is appended to the global hit list with one atomic per flush;
// per-lane counters give the count table.
//
// Algorithmic work: 2*W lookup-adds per base x motif (both strands).  The binding unit is the FP32 pipe
// (128 lanes/clk/SM; FADD2 issues 64 adds per slot), not shared memory.
#pragma once
#include "common.cuh"

namespace smg {

__device__ __forceinline__ float2 vadd(const float2 a, const float2 b) { return __fadd2_rn(a, b); }
__device__ __forceinline__ float vmax3(const float m, const float2 a)
{
    float d;
    asm("max.f32 %0, %1, %2, %3;" : "=f"(d) : "f"(m), "f"(a.x), "f"(a.y));
    return d;
}
__device__ __forceinline__ float2 vmix(const float e0, const float e1, const float e2, const float e3, const float2 t0,
                                       const float2 t1, const float2 t2, const float2 t3)
{
    return make_float2(smg_mix(e0, e1, e2, e3, t0.x, t1.x, t2.x, t3.x), smg_mix(e0, e1, e2, e3, t0.y, t1.y, t2.y, t3.y));
}

// Register-resident log-odds table of one PWM of one lane: pairs P(k,b) = {T[k][b] (forward), Trc[k][b] (reverse
// complement)}.  With the RC table padded at its HEAD (rows k < WB-W are zero) Trc[k][b] = T[WB-1-k][3-b] holds for
// every k, hence P(WB-1-k, 3-b) is P(k,b) with its halves swapped -- and FADD2 takes a half-swapped operand for free
// (SASS operand modifier .F32x2.LO_HI).  Only half of the pairs are kept: 4*WB registers instead of 8*WB, which is
// what buys the extra resident warps.
template <int WB> struct Tab {
    static constexpr int H = WB / 2;
    float2 h[H][4];
    float2 hm[2];  // middle row (odd WB): bases A, C; G and T are their swapped partners
    __device__ __forceinline__ void load(const float* __restrict__ gt_, const int lane)
    {
        const float2* gt = reinterpret_cast<const float2*>(gt_);
#pragma unroll
        for (int k = 0; k < H; ++k)
#pragma unroll
            for (int b = 0; b < 4; ++b) h[k][b] = gt[(k * 4 + b) * 32 + lane];
        if (WB & 1) {
            hm[0] = gt[(H * 4 + 0) * 32 + lane];
            hm[1] = gt[(H * 4 + 1) * 32 + lane];
        } else {
            hm[0] = hm[1] = make_float2(0.f, 0.f);
        }
    }
    __device__ __forceinline__ float2 get(const int k, const int b) const
    {
        if (k < H) return h[k][b];
        if (k >= WB - H) {
            const float2 p = h[WB - 1 - k][3 - b];
            return make_float2(p.y, p.x);
        }
        if (b < 2) return hm[b];
        const float2 p = hm[3 - b];
        return make_float2(p.y, p.x);
    }
};

constexpr int kQCap = 16;        // per-lane candidate queue slots (32 with two PWMs per lane)
constexpr int kStageCap = 128;   // per-warp staged hits
constexpr int kStageFlush = 64;  // flush when more than this many are staged (one pass adds <= 32)
constexpr int kU = 4;            // steps between threshold checks on the ACGT fast path (divides 16)

// one shift-register step for base B (compile-time) of all NM PWMs of the lane; o[m][U] receives the completed scores
#define SMG_STEP(B, U)                                                                                     \
    {                                                                                                      \
        _Pragma("unroll") for (int m_ = 0; m_ < NM; ++m_) {                                                \
            o[m_][U] = vadd(r[m_][WB - 2], t[m_].get(WB - 1, B));                                          \
            _Pragma("unroll") for (int k = WB - 2; k >= 1; --k) r[m_][k] = vadd(r[m_][k - 1], t[m_].get(k, B)); \
            r[m_][0] = t[m_].get(0, B);                                                                    \
        }                                                                                                  \
    }

// Threaded dispatch: the body of step u first evaluates the two bit tests of step u+1's base straight off the packed
// word (so their predicate latency hides behind its own FADD2s) and then jumps directly into the next step's body --
// there is no common dispatch block inside a group of kU = 4 steps.  +12% over a per-step switch (measured).
#define SMG_GOTO_NEXT(u, w, P)                                                            \
    if ((w) & (2u << (2 * ((u) + 1)))) {                                                  \
        if ((w) & (1u << (2 * ((u) + 1)))) goto P##3; else goto P##2;                     \
    } else {                                                                              \
        if ((w) & (1u << (2 * ((u) + 1)))) goto P##1; else goto P##0;                     \
    }

// Explicit register caps per instantiation (an uncapped __maxnreg__ makes ptxas spend MORE registers than its
// default heuristic and lose occupancy; measured).  Registers are partitioned per SM sub-partition (16K each), so
// the occupancy steps are 64/72/80/96/128/168 registers = 8/7/6/5/4/3 one-warp CTAs per sub-partition; the caps
// below are the measured best step for each width (tools/bench_widths.py).
#ifndef SMG_REGCAPS
#define SMG_REGCAPS(WB) ((WB) <= 6 ? 72 : (WB) <= 8 ? 80 : (WB) <= 10 ? 96 : (WB) <= 12 ? 104 : (WB) <= 16 ? 128 : (WB) <= 24 ? 168 : 240)
#endif
#ifndef SMG_REGCAPS2
#define SMG_REGCAPS2(WB) ((WB) <= 8 ? 128 : 168)
#endif
template <int WB, int NM> struct MaxRegs { static constexpr int v = NM == 1 ? SMG_REGCAPS(WB) : SMG_REGCAPS2(WB); };

template <int WB, int NM>
__global__ void __launch_bounds__(32) __maxnreg__((MaxRegs<WB, NM>::v)) scan_regtab_kernel(const ScanParams p)
{
    static_assert(WB >= 2 && WB <= SMG_MAX_REGTAB_WIDTH && NM >= 1 && NM <= SMG_MAX_PWMS_PER_LANE, "unsupported shape");
    using V = float2;
    __shared__ unsigned long long s_keys[kStageCap];
    __shared__ float s_scores[kStageCap];
    __shared__ uint2 s_q[kQCap * NM][32];  // per-lane candidate queues, [slot][lane] (bank-conflict free)

    const int lane = threadIdx.x;
    const unsigned lt_mask = (1u << lane) - 1u;
    const int group = p.first_group + blockIdx.y;
    const smg_chunk_t ch = p.chunks[blockIdx.x];
    const smg_row_t row = p.rows[ch.row];
    const float pinf = __int_as_float(0x7f800000);

    // ---- per-lane PWMs: tables into registers, thresholds, trim bounds
    Tab<WB> t[NM];
    int motif[NM], wm[NM];
    float thr_lo[NM], thr_hi[NM];
#pragma unroll
    for (int m = 0; m < NM; ++m) {
        t[m].load(p.group_tab + p.group_tab_off[group] + (size_t)m * (WB * 4 * 32 * 2), lane);
        motif[m] = p.lane_motif[(group * SMG_MAX_PWMS_PER_LANE + m) * 32 + lane];
        wm[m] = motif[m] >= 0 ? p.motif_width[motif[m]] : WB;
        thr_lo[m] = motif[m] >= 0 ? p.thr_lo[motif[m]] : pinf;
        thr_hi[m] = motif[m] >= 0 ? p.thr_hi[motif[m]] : pinf;
    }
    const long long avail64 = row.g_len - row.g_off;  // smax (largest window start that fits) = avail64 - W
    const bool want_hits = p.hit_keys != nullptr;
    const int key_shift = p.pos_bits + p.motif_bits;
    const unsigned long long rowkey0 = (unsigned long long)(unsigned)p.out_rows[ch.row * 2 + 0] << key_shift;
    const unsigned long long rowkey1 = (unsigned long long)(unsigned)p.out_rows[ch.row * 2 + 1] << key_shift;

    const int c0 = ch.start;
    const int c1 = min(c0 + p.chunk_len, row.n_own);  // owned window starts [c0, c1)
    const int n_steps = (c1 - c0) + WB - 1;
    const int n_words = (n_steps + 15) >> 4;
    const uint32_t* __restrict__ wp = p.packed + row.word_off + (c0 >> 4);
    const uint16_t* __restrict__ ap = p.amask + row.word_off + (c0 >> 4);
    const uint8_t* __restrict__ cp = p.codes ? p.codes + row.word_off * 16 : nullptr;

    V r[NM][WB - 1];
#pragma unroll
    for (int m = 0; m < NM; ++m)
#pragma unroll
        for (int k = 0; k < WB - 1; ++k) r[m][k] = make_float2(0.f, 0.f);
    int cnt[NM][2];  // definite hits of this lane per PWM and strand
#pragma unroll
    for (int m = 0; m < NM; ++m) cnt[m][0] = cnt[m][1] = 0;
    int nst = 0;  // staged hits (warp-uniform)

    auto flush = [&]() {
        unsigned long long base = 0;
        if (lane == 0) base = atomicAdd(&p.counters[SMG_CTR_HITS_APPENDED], (unsigned long long)nst);
        base = __shfl_sync(SMG_FULL_MASK, base, 0);
        for (int i = lane; i < nst; i += 32) {
            if (base + i < p.hit_cap) {
                p.hit_keys[base + i] = s_keys[i];
                p.hit_scores[base + i] = s_scores[i];
            }
        }
        nst = 0;
        __syncwarp();
    };

    // Candidate handling.  Two levels keep the per-block cost low:
    //  (1) push(): a lane whose completed window score exceeds thr_lo appends (step, PWM slot, strand, score) to ITS
    //      OWN queue in shared memory -- predicated store + pointer bump, no ballot / shuffle / loop in the scan loop;
    //  (2) drain(): when any queue may overflow (and at the end of the chunk) the warp walks the queues together:
    //      range / trim checks, definite-vs-near split, per-lane counters, ballot + prefix-popcount compaction into
    //      the warp stage, which is appended to the global hit list with one atomic per flush.
    const int s_origin = c0 - (WB - 1);  // window start completed by the first step of the chunk
    // the queue is addressed through a running shared-memory byte address: one predicated 8-byte store + one
    // predicated add per candidate (tag = step index within the chunk, PWM slot, strand)
    const unsigned qbase = (unsigned)__cvta_generic_to_shared(&s_q[0][lane]);
    unsigned qaddr = qbase;
    constexpr unsigned kQStride = 32 * sizeof(uint2);
    auto push = [&](const float v, const unsigned tag) {
        asm volatile("st.shared.v2.b32 [%0], {%1, %2};" ::"r"(qaddr), "r"(tag), "f"(v) : "memory");
        qaddr += kQStride;
    };
    auto drain = [&]() {
        const int qn = (int)((qaddr - qbase) / kQStride);
        const int maxn = __reduce_max_sync(SMG_FULL_MASK, qn);
        for (int j = 0; j < maxn; ++j) {
            const bool have = j < qn;
            const uint2 e = have ? s_q[j][lane] : make_uint2(0u, 0u);
            const float v = __uint_as_float(e.y);
            const int st = (int)(e.x & 1u);
            const int sub = NM > 1 ? (int)((e.x >> 1) & 1u) : 0;
            const int my_motif = (NM > 1 && sub) ? motif[NM - 1] : motif[0];
            const int my_w = (NM > 1 && sub) ? wm[NM - 1] : wm[0];
            const float my_hi = (NM > 1 && sub) ? thr_hi[NM - 1] : thr_hi[0];
            // the RC table is head-padded: its completed window starts WB - W bases later
            const int s = s_origin + (int)(e.x >> 2) + (st ? (WB - my_w) : 0);
            const long long smax64 = avail64 - (long long)my_w;
            const bool cand = have && (s >= c0) && (s < c1) && ((long long)s <= smax64) && ((p.strands >> st) & 1);
            const bool hit = cand && (v > my_hi);
            if (hit) {
#pragma unroll
                for (int m = 0; m < NM; ++m) {
                    if (sub == m) {
                        if (st) ++cnt[m][1]; else ++cnt[m][0];
                    }
                }
            }
            if (cand && !hit) {  // near-threshold band: defer to the fp64 adjudicator
                const unsigned long long i = atomicAdd(&p.counters[SMG_CTR_NEAR], 1ull);
                if (i < p.near_cap) {
                    smg_near_t n;
                    n.row = ch.row; n.motif = my_motif; n.strand = st; n.start = s; n.score = v; n.reserved = 0;
                    p.near[i] = n;
                }
            }
            if (want_hits) {
                const unsigned hm = __ballot_sync(SMG_FULL_MASK, hit);
                if (hm) {
                    if (hit) {
                        const long long pos = st ? (smax64 - (long long)s) : (row.g_off + (long long)s);
                        const int slot = nst + __popc(hm & lt_mask);
                        s_keys[slot] = (st ? rowkey1 : rowkey0) | ((unsigned long long)pos << p.motif_bits) |
                                       (unsigned long long)(unsigned)my_motif;
                        s_scores[slot] = v;
                    }
                    nst += __popc(hm);
                    __syncwarp();
                    if (nst > kStageFlush) flush();
                }
            }
        }
        qaddr = qbase;
        __syncwarp();
    };

    uint32_t w = wp[0];
    uint32_t am = ap[0];
    int s_base = c0 - (WB - 1);  // window start completed by the first step of the current word
    for (int wi = 0; wi < n_words; ++wi) {
        const uint32_t wn = wp[wi + 1];  // rows are padded, the prefetch never leaves the allocation
        const uint32_t an = ap[wi + 1];
        if (am == 0) {
            // ---- ACGT fast path: 16 bases, threshold check every kU steps
#pragma unroll 1
            for (int blk = 0; blk < 16 / kU; ++blk) {
                V o[NM][kU];
                static_assert(kU == 4, "the threaded dispatch below is written for groups of 4 steps");
                if (w & 2u) { if (w & 1u) goto sA3; else goto sA2; } else { if (w & 1u) goto sA1; else goto sA0; }
                sA0: SMG_STEP(0, 0); SMG_GOTO_NEXT(0, w, sB)
                sA1: SMG_STEP(1, 0); SMG_GOTO_NEXT(0, w, sB)
                sA2: SMG_STEP(2, 0); SMG_GOTO_NEXT(0, w, sB)
                sA3: SMG_STEP(3, 0); SMG_GOTO_NEXT(0, w, sB)
                sB0: SMG_STEP(0, 1); SMG_GOTO_NEXT(1, w, sC)
                sB1: SMG_STEP(1, 1); SMG_GOTO_NEXT(1, w, sC)
                sB2: SMG_STEP(2, 1); SMG_GOTO_NEXT(1, w, sC)
                sB3: SMG_STEP(3, 1); SMG_GOTO_NEXT(1, w, sC)
                sC0: SMG_STEP(0, 2); SMG_GOTO_NEXT(2, w, sD)
                sC1: SMG_STEP(1, 2); SMG_GOTO_NEXT(2, w, sD)
                sC2: SMG_STEP(2, 2); SMG_GOTO_NEXT(2, w, sD)
                sC3: SMG_STEP(3, 2); SMG_GOTO_NEXT(2, w, sD)
                sD0: SMG_STEP(0, 3); goto sEnd;
                sD1: SMG_STEP(1, 3); goto sEnd;
                sD2: SMG_STEP(2, 3); goto sEnd;
                sD3: SMG_STEP(3, 3);
                sEnd:
                w >>= 2 * kU;
                // two-level check: one vote for the whole group of 4 steps, then one per half, so that a group with a
                // single candidate issues 4 predicated pushes instead of 8
                bool any_lo = false, any_hi = false;
#pragma unroll
                for (int m = 0; m < NM; ++m) {
                    const float m01 = vmax3(fmaxf(o[m][0].x, o[m][0].y), o[m][1]);
                    const float m23 = vmax3(fmaxf(o[m][2].x, o[m][2].y), o[m][3]);
                    any_lo = any_lo || (m01 > thr_lo[m]);
                    any_hi = any_hi || (m23 > thr_lo[m]);
                }
                if (__any_sync(SMG_FULL_MASK, any_lo || any_hi)) {
                    const unsigned tag0 = (unsigned)(wi * 16 + blk * kU) << 2;
#pragma unroll
                    for (int half = 0; half < 2; ++half) {
                        if (__any_sync(SMG_FULL_MASK, half ? any_hi : any_lo)) {
#pragma unroll
                            for (int m = 0; m < NM; ++m) {
#pragma unroll
                                for (int u = 2 * half; u < 2 * half + 2; ++u) {
                                    if (o[m][u].x > thr_lo[m]) push(o[m][u].x, tag0 + (unsigned)((u << 2) | (m << 1)));
                                    if (o[m][u].y > thr_lo[m]) push(o[m][u].y, tag0 + (unsigned)((u << 2) | (m << 1) | 1));
                                }
                            }
                        }
                    }
                    // a group adds at most 2*kU entries per PWM slot
                    if (__any_sync(SMG_FULL_MASK, qaddr > qbase + (kQCap * NM - 2 * kU * NM) * kQStride)) drain();
                }
            }
        } else {
            // ---- generic path: IUPAC ambiguity codes, 'Z', and the padding past the end of the row
#pragma unroll 1
            for (int j = 0; j < 16; ++j) {
                V o[NM][1];
                if (((am >> j) & 1u) == 0) {
                    switch ((w >> (2 * j)) & 3u) {
                    case 0: SMG_STEP(0, 0); break;
                    case 1: SMG_STEP(1, 0); break;
                    case 2: SMG_STEP(2, 0); break;
                    default: SMG_STEP(3, 0); break;
                    }
                } else {
                    const int q = s_base + (WB - 1) + j;  // row-local position of this base
                    const int code = (cp != nullptr && q < row.n_avail) ? (int)cp[q] : 0;
                    const float e0 = c_emb[code][0], e1 = c_emb[code][1], e2 = c_emb[code][2], e3 = c_emb[code][3];
#pragma unroll
                    for (int m = 0; m < NM; ++m) {
                        o[m][0] = vadd(r[m][WB - 2], vmix(e0, e1, e2, e3, t[m].get(WB - 1, 0), t[m].get(WB - 1, 1),
                                                          t[m].get(WB - 1, 2), t[m].get(WB - 1, 3)));
#pragma unroll
                        for (int k = WB - 2; k >= 1; --k)
                            r[m][k] = vadd(r[m][k - 1], vmix(e0, e1, e2, e3, t[m].get(k, 0), t[m].get(k, 1), t[m].get(k, 2),
                                                             t[m].get(k, 3)));
                        r[m][0] = vmix(e0, e1, e2, e3, t[m].get(0, 0), t[m].get(0, 1), t[m].get(0, 2), t[m].get(0, 3));
                    }
                }
                bool any_c = false;
#pragma unroll
                for (int m = 0; m < NM; ++m) any_c = any_c || (fmaxf(o[m][0].x, o[m][0].y) > thr_lo[m]);
                if (__any_sync(SMG_FULL_MASK, any_c)) {
                    const unsigned tag0 = (unsigned)(wi * 16 + j) << 2;
#pragma unroll
                    for (int m = 0; m < NM; ++m) {
                        if (o[m][0].x > thr_lo[m]) push(o[m][0].x, tag0 + (unsigned)(m << 1));
                        if (o[m][0].y > thr_lo[m]) push(o[m][0].y, tag0 + (unsigned)((m << 1) | 1));
                    }
                    if (__any_sync(SMG_FULL_MASK, qaddr > qbase + (kQCap * NM - 2 * kU * NM) * kQStride)) drain();
                }
            }
        }
        w = wn;
        am = an;
        s_base += 16;
    }

    drain();
    if (want_hits && nst > 0) flush();
    int tot = 0;
#pragma unroll
    for (int m = 0; m < NM; ++m) {
        if (p.counts != nullptr && motif[m] >= 0) {
            int32_t* c = p.counts + ((size_t)ch.row * p.n_motifs + motif[m]) * 2;
            if (cnt[m][0]) atomicAdd(c + 0, cnt[m][0]);
            if (cnt[m][1]) atomicAdd(c + 1, cnt[m][1]);
        }
        tot += cnt[m][0] + cnt[m][1];
    }
#pragma unroll
    for (int d = 16; d >= 1; d >>= 1) tot += __shfl_xor_sync(SMG_FULL_MASK, tot, d);
    if (lane == 0 && tot) atomicAdd(&p.counters[SMG_CTR_DEFINITE], (unsigned long long)tot);
}

template <int WB, int NM>
void launch_scan_regtab(const ScanParams& p, int n_chunks, int n_groups, cudaStream_t stream)
{
    dim3 grid((unsigned)n_chunks, (unsigned)n_groups, 1);
    scan_regtab_kernel<WB, NM><<<grid, 32, 0, stream>>>(p);
}

}  // namespace smg
