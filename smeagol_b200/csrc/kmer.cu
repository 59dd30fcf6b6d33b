// K-MER INDEX ENGINE for narrow PWMs (smg_kmer_index_build / smg_scan_kmer).
//
// A window of a PWM of width W over A/C/G/T has only 4^W possible contents.  For W <= 12 (and a scan long enough to pay
// for it) the engine therefore scores every k-mer ONCE per call -- in fp32 in the fixed ascending-k order of the
// register-table kernel, with the near-threshold band adjudicated in fp64 on the spot -- and keeps, per k-mer, the short
// list of (PWM, strand, score) that are hits (hit density 1e-4 .. 6e-3: the lists are mostly empty).  The scan itself is
// then OUTPUT-BOUND: per window start one 2W-bit mask of the packed bases per width class, one table lookup, and one
// list walk that emits finished hits with scores bit-identical to every other engine.  No per-cell arithmetic at all.
//
//   index   head[class offset + x]        first entry of k-mer x (or -1); x = sum_k base(p + k) << 2k
//           entry {col | flags, score, next}   col = motif * 2 + strand; singly linked (built in one pass with atomicExch)
//   build   one warp = one subtree of (W-3)-base prefixes x 16 motifs x 2 strands (lane = column), walked depth first: a
//           prefix sum is one rounded addition on top of its parent's (same rounding as a sequential ascending-k sum),
//           and an exact upper bound of the 64 extensions prunes almost every prefix before any of them is formed
//   scan    one thread = one window start; a CTA stages its hits in shared memory and appends them to the global hit list
//           with one atomic per flush; per-(motif, strand) counts in a shared-memory histogram, flushed per chunk
//   windows that contain an IUPAC code, 'Z' or the end of the row are scored directly (rare generic path).
//
// SEED ENGINE (smg_seed_index_build / smg_scan_seed, at the end of this file): the same index machinery for the WIDE
// PWMs -- each (PWM, strand) enters a one-class index as a forward-only virtual PWM made of its most selective S
// positions with a relaxed threshold; matches become candidate windows for the exact second pass (kmer.h).
#include <algorithm>
#include <string>
#include <vector>

#include "common.cuh"
#include "kmer.h"

namespace smg {

// ---------------------------------------------------------------------------------------------------- build
// BRANCH AND BOUND: rounded fp32 addition is monotone in each operand, so the score of every k-mer that extends a
// prefix is bounded by the SAME chain of additions fed with the per-position maxima,
//     ub = ((acc + max_b T[W-3][b]) + max_b T[W-2][b]) + max_b T[W-1][b]      (fixed order, fp32, every step rounded)
// -- an exact bound, not an estimate.  At hit densities of 1e-4 .. 6e-3 almost every prefix fails `ub > thr_lo` and its
// 64 extensions are never formed; the survivors are expanded by the whole warp (lane = extension).
//
// Warp-private allocation state of the index build.  Entry slots are reserved kSlotBlock at a time (ONE atomic round trip
// per block; hits take their slots by ballot + prefix count), and the link of an entry -- the previous head returned by
// atomicExch -- is stored one emission LATER, so that the round trip of the exchange overlaps the next expansion
// instead of stalling the in-order warp (the build was latency-bound on exactly these two atomics).  The lists are only
// read by the scan kernel, i.e. after this kernel has completed, so no ordering between entry and head is needed here.
constexpr int kSlotBlock = 64;
struct KmerBuildState {
    unsigned long long slot_base;  // warp-uniform
    int slot_used;                 // warp-uniform
    long long pend_idx;            // per lane: entry whose `next` is still in flight (-1: none)
    int pend_prev;
};

__device__ __forceinline__ void kmer_flush_link(const KmerParams& kp, KmerBuildState& bs)
{
    if (bs.pend_idx >= 0) kp.entries[bs.pend_idx].next = bs.pend_prev;
    bs.pend_idx = -1;
}

// called by a converged warp; `hit` lanes append one entry each
__device__ __forceinline__ void kmer_emit_entries(const KmerParams& kp, const ScanParams& p, KmerBuildState& bs, const bool hit,
                                                  const int lane, const int W, const long long x, const float s, const float* tab,
                                                  const int st, const int last, const int motif, const float hi, const double t64,
                                                  int32_t* head)
{
    const unsigned hm = __ballot_sync(SMG_FULL_MASK, hit);
    if (!hm) return;
    const int n = __popc(hm);
    if (bs.slot_used + n > kSlotBlock) {  // next block (the unused tail of the old one stays empty: never linked)
        unsigned long long nb = 0;
        if (lane == 0) nb = atomicAdd(&p.counters[SMG_CTR_KMER_ENTRIES], (unsigned long long)kSlotBlock);
        bs.slot_base = __shfl_sync(SMG_FULL_MASK, nb, 0);
        bs.slot_used = 0;
    }
    const unsigned long long idx = bs.slot_base + bs.slot_used + __popc(hm & ((1u << lane) - 1u));
    bs.slot_used += n;
    if (!hit) return;
    uint32_t flags = 0;
    if (!(s > hi)) {  // near-threshold band: decide in fp64 (same operands, same order)
        double a64 = 0.0;
        for (int k = 0; k < W; ++k) {
            const int j = 4 * k + (int)((x >> (2 * k)) & 3);
            const double term = (double)tab[st ? (last - j) : j];
            a64 = (k == 0) ? term : __dadd_rn(a64, term);
        }
        flags = kEntNear | ((a64 > t64) ? 0u : kEntRej);
    }
    if (idx < kp.entry_cap) {
        kmer_flush_link(kp, bs);  // the previous exchange of this lane has long returned
        smg_kmer_entry_t e;
        e.col = (uint32_t)(motif * 2 + st) | flags;
        e.score = s;
        e.next = -1;
        e.reserved = 0;
        kp.entries[idx] = e;
        bs.pend_prev = atomicExch(&head[x], (int32_t)idx);
        bs.pend_idx = (long long)idx;
        atomicOr(reinterpret_cast<unsigned*>(kp.head + kp.bm_off[W]) + (x >> 5), 1u << (x & 31));
        if (W >= kKmerBitmapMinW && kp.max_w > kKmerBitmapMinW) {
            // combined bitmap of the wide classes, keyed by the max_w-mer: every extension of this k-mer
            unsigned* cb = reinterpret_cast<unsigned*>(kp.head + kp.cb_off);
            const long long n_ext = 1ll << (2 * (kp.max_w - W));
            for (long long t2 = 0; t2 < n_ext; ++t2) {
                const long long y = x | (t2 << (2 * W));
                atomicOr(cb + (y >> 5), 1u << (y & 31));
            }
        }
    }
}

// One WARP = one subtree of prefixes x one tile of 16 motifs; lane = (motif of the tile, strand).  The prefix tree
// (bases k = 0 .. pre-1, the ascending-k order of the sums) is walked depth first, so a node costs every lane ONE rounded
// addition for its own column -- a prefix sum is never recomputed from k = 0 -- and a leaf costs the bound chain.
constexpr int kBuildTile = 16;
constexpr int kBuildRow = kKmerMaxW * 4 + 1;  // padded: the 16 rows of a tile start in 16 different banks

struct KmerBuildCtx {
    const KmerParams* kp;
    const float* tab;      // this lane's column (shared memory)
    const float* tab_all;  // the tile's tables
    const float* s_hi;
    const double* s_t64;
    const int* s_m;
    int32_t* head;
    float lo, mx0, mx1, mx2;
    int W, pre, d0, last, st, lane;
    bool on;
};

template <int TAIL>
__device__ __forceinline__ void kmer_build_leaf(const KmerBuildCtx& c, KmerBuildState& bs, const float acc, const long long q)
{
    // exact upper bound of all 4^TAIL extensions (same chain of rounded additions, maxima as operands)
    float ub = (c.pre == 0) ? c.mx0 : __fadd_rn(acc, c.mx0);
    if (TAIL >= 2) ub = __fadd_rn(ub, c.mx1);
    if (TAIL >= 3) ub = __fadd_rn(ub, c.mx2);
    unsigned surv = __ballot_sync(SMG_FULL_MASK, c.on && ub > c.lo);
    // surviving (prefix, column) pairs are expanded by the WHOLE warp, one at a time (lane = extension)
    constexpr int kExt = 1 << (2 * TAIL);
    while (surv) {
        const int src = __ffs(surv) - 1;
        surv &= surv - 1;
        const float a = __shfl_sync(SMG_FULL_MASK, acc, src);
        const float lo = __shfl_sync(SMG_FULL_MASK, c.lo, src);
        const int t = src >> 1, st = src & 1;
        const float* tab = c.tab_all + t * kBuildRow;
#pragma unroll
        for (int e0 = 0; e0 < kExt; e0 += 32) {
            const int e = e0 + c.lane;
            float sc = a;
#pragma unroll
            for (int u = 0; u < TAIL; ++u) {
                const int j = 4 * (c.pre + u) + ((e >> (2 * u)) & 3);
                const float term = tab[st ? (c.last - j) : j];
                sc = (c.pre + u == 0) ? term : __fadd_rn(sc, term);
            }
            kmer_emit_entries(*c.kp, c.kp->sp, bs, e < kExt && sc > lo, c.lane, c.W, q | ((long long)e << (2 * c.pre)), sc, tab, st,
                              c.last, c.s_m[t], c.s_hi[t], c.s_t64[t], c.head);
        }
    }
}

template <int DEPTH, int SUB, int TAIL> struct KmerBuildDfs {
    static __device__ __forceinline__ void run(const KmerBuildCtx& c, KmerBuildState& bs, const float acc, const long long q)
    {
        const int k = c.d0 + DEPTH;
#pragma unroll 1
        for (int b = 0; b < 4; ++b) {
            const int j = 4 * k + b;
            const float term = c.tab[c.st ? (c.last - j) : j];
            KmerBuildDfs<DEPTH + 1, SUB, TAIL>::run(c, bs, k == 0 ? term : __fadd_rn(acc, term), q | ((long long)b << (2 * k)));
        }
    }
};
template <int SUB, int TAIL> struct KmerBuildDfs<SUB, SUB, TAIL> {
    static __device__ __forceinline__ void run(const KmerBuildCtx& c, KmerBuildState& bs, const float acc, const long long q)
    {
        kmer_build_leaf<TAIL>(c, bs, acc, q);
    }
};

// grid.x: subtrees (the first d0 = pre - SUB prefix bases), 4 per CTA; grid.y: tiles of 16 motifs of the class
template <int SUB, int TAIL>
__global__ void __launch_bounds__(128) kmer_build_kernel(const __grid_constant__ KmerParams kp, const int W)
{
    const ScanParams& p = kp.sp;
    __shared__ float s_tab[kBuildTile * kBuildRow];
    __shared__ float s_hi[kBuildTile];
    __shared__ double s_t64[kBuildTile];
    __shared__ int s_m[kBuildTile];
    const int c0 = kp.cls_off[W], n_cm = kp.cls_off[W + 1] - c0;
    const int m0 = blockIdx.y * kBuildTile;
    const int nt = min(kBuildTile, n_cm - m0);
    for (int i = threadIdx.x; i < nt * W * 4; i += blockDim.x) {
        const int t = i / (W * 4), j = i % (W * 4);
        const int m = kp.cls_motifs[c0 + m0 + t];
        s_tab[t * kBuildRow + j] = p.nat_tab[p.motif_off[m] * 4 + j];
    }
    if (threadIdx.x < nt) {
        const int m = kp.cls_motifs[c0 + m0 + threadIdx.x];
        s_m[threadIdx.x] = m;
        s_hi[threadIdx.x] = p.thr_hi[m];
        s_t64[threadIdx.x] = kp.thr64[m];
    }
    __syncthreads();
    KmerBuildCtx c;
    c.kp = &kp;
    c.W = W; c.pre = W - TAIL; c.d0 = c.pre - SUB; c.last = 4 * W - 1;
    c.lane = threadIdx.x & 31;
    const int t = c.lane >> 1;
    c.st = c.lane & 1;
    c.on = t < nt && ((p.strands >> c.st) & 1);
    c.tab_all = s_tab; c.tab = s_tab + (t < nt ? t : 0) * kBuildRow;
    c.s_hi = s_hi; c.s_t64 = s_t64; c.s_m = s_m;
    c.head = kp.head + kp.head_off[W];
    c.lo = c.on ? p.thr_lo[s_m[t]] : 0.f;
    float mx[3] = {0.f, 0.f, 0.f};
#pragma unroll
    for (int u = 0; u < TAIL; ++u) {  // forward: T[k][b] = tab[4k + b]; reverse complement: tab[last - (4k + b)]
        const int j = 4 * (c.pre + u);
        const float4 v = c.st ? make_float4(c.tab[c.last - j], c.tab[c.last - j - 1], c.tab[c.last - j - 2], c.tab[c.last - j - 3])
                              : make_float4(c.tab[j], c.tab[j + 1], c.tab[j + 2], c.tab[j + 3]);
        mx[u] = fmaxf(fmaxf(v.x, v.y), fmaxf(v.z, v.w));
    }
    c.mx0 = mx[0]; c.mx1 = mx[1]; c.mx2 = mx[2];
    const long long n_sub = 1ll << (2 * c.d0);
    const long long sub = (long long)blockIdx.x * 4 + (threadIdx.x >> 5);
    if (sub >= n_sub) return;  // whole warps
    float acc = 0.f;
    for (int k = 0; k < c.d0; ++k) {
        const int j = 4 * k + (int)((sub >> (2 * k)) & 3);
        const float term = c.tab[c.st ? (c.last - j) : j];
        acc = (k == 0) ? term : __fadd_rn(acc, term);
    }
    KmerBuildState bs;
    bs.slot_base = 0; bs.slot_used = kSlotBlock;  // the first emission reserves a block
    bs.pend_idx = -1; bs.pend_prev = -1;
    KmerBuildDfs<0, SUB, TAIL>::run(c, bs, acc, sub);
    kmer_flush_link(kp, bs);
}

// ---------------------------------------------------------------------------------------------------- scan
// Per-warp hit stage, double buffered: when a buffer is full its slots in the global hit list are reserved (one atomic,
// result not awaited), the warp goes on filling the other buffer, and the first one is copied out when the warp next
// switches -- by then the atomic has long returned.  Warps never wait for one another inside a chunk.
struct KmerWarpStage {
    unsigned long long keys[2][kKmerStage];
    float scores[2][kKmerStage];
    int n;  // fill of the active buffer
};

struct KmerEmit {  // per-thread view of its warp's stage (registers)
    KmerWarpStage* st;
    int cur;                       // active buffer
    int pend_n;                    // hits of the other buffer awaiting copy-out (0: none)
    unsigned long long pend_base;  // their first slot in the global list (valid in lane 0 until broadcast)
};

__device__ __forceinline__ void kmer_push(const ScanParams& p, KmerEmit& em, const unsigned long long key, const float score)
{
    const int slot = atomicAdd(&em.st->n, 1);  // one address per warp: aggregated by the compiler
    if (slot < kKmerStage) {
        em.st->keys[em.cur][slot] = key;
        em.st->scores[em.cur][slot] = score;
    } else {  // buffer full: straight to the global list (the warp switches buffers when one is half full, so this is rare)
        const unsigned long long gi = atomicAdd(&p.counters[SMG_CTR_HITS_APPENDED], 1ull);
        if (gi < p.hit_cap) {
            p.hit_keys[gi] = key;
            p.hit_scores[gi] = score;
        }
    }
}

// copy out the pending buffer (if any), then make the active buffer pending (reserve its slots) and switch.  Converged warp.
__device__ __forceinline__ void kmer_switch(const ScanParams& p, KmerEmit& em, const int lane)
{
    __syncwarp();
    if (em.pend_n > 0) {
        const unsigned long long base = __shfl_sync(SMG_FULL_MASK, em.pend_base, 0);
        const int ob = em.cur ^ 1;
        for (int i = lane; i < em.pend_n; i += 32)
            if (base + i < p.hit_cap) {
                p.hit_keys[base + i] = em.st->keys[ob][i];
                p.hit_scores[base + i] = em.st->scores[ob][i];
            }
        em.pend_n = 0;
    }
    const int n = min(em.st->n, kKmerStage);
    __syncwarp();
    if (n > 0) {
        if (lane == 0) {
            em.pend_base = atomicAdd(&p.counters[SMG_CTR_HITS_APPENDED], (unsigned long long)n);
            em.st->n = 0;
        }
        em.pend_n = n;
        em.cur ^= 1;
    }
    __syncwarp();
}

// direct scoring of one window that the index cannot serve (IUPAC code, 'Z', end padding): every PWM of the class
__device__ __noinline__ void kmer_slow_window(const KmerParams& kp, const smg_row_t& row, const int r, const int W, const long long s,
                                              KmerEmit& em, int* s_hist, unsigned* tallies)
{
    const ScanParams& p = kp.sp;
    const int key_shift = p.pos_bits + p.motif_bits;
    const long long smax = row.g_len - row.g_off - (long long)W;
    for (int i = kp.cls_off[W]; i < kp.cls_off[W + 1]; ++i) {
        const int m = kp.cls_motifs[i];
        for (int st = 0; st < 2; ++st) {
            if (!((p.strands >> st) & 1)) continue;
            const float v = window_score32(p, row, m, st, s);
            if (!(v > p.thr_lo[m])) continue;
            bool hit = true;
            if (!(v > p.thr_hi[m])) {
                ++tallies[1];
                const double s64 = window_score64(p.packed, p.amask, p.codes, row, p.nat_tab + p.motif_off[m] * 4, W, st, s, -1, 0);
                hit = s64 > kp.thr64[m];
                if (hit) ++tallies[2];
            } else {
                ++tallies[0];
            }
            if (!hit) continue;
            if (s_hist) atomicAdd(&s_hist[m * 2 + st], 1);
            else if (p.counts) atomicAdd(p.counts + ((size_t)r * p.n_motifs + m) * 2 + st, 1);
            if (p.hit_keys) {
                const long long pos = st ? (smax - s) : (row.g_off + s);
                kmer_push(p, em, ((unsigned long long)(unsigned)p.out_rows[r * 2 + st] << key_shift) |
                                     ((unsigned long long)pos << p.motif_bits) | (unsigned long long)(unsigned)m, v);
            }
        }
    }
}

// dynamic shared memory: [2 * n_motifs] int histogram (hist_in_smem) -- counts of this chunk's row
__global__ void __launch_bounds__(kKmerThreads, 3) scan_kmer_kernel(const __grid_constant__ KmerParams kp, const int hist_in_smem)
{
    const ScanParams& p = kp.sp;
    extern __shared__ int s_hist_dyn[];
    __shared__ KmerWarpStage s_stage[kKmerThreads / 32];
    __shared__ int2 s_wq[kKmerThreads / 32][kKmerQueue];  // per-warp work queue of non-empty lists
    __shared__ unsigned s_tally[3];
    __shared__ uint32_t s_bm[kKmerSmemBitmapWords];  // bitmaps of the classes narrower than kKmerBitmapMinW
    int* s_hist = hist_in_smem && p.counts ? s_hist_dyn : nullptr;
    const smg_chunk_t ch = p.chunks[blockIdx.x];
    const smg_row_t row = p.rows[ch.row];
    const int c0 = ch.start, c1 = min(c0 + p.chunk_len, row.n_own);
    if (c1 <= c0) return;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const unsigned lt_mask = (1u << lane) - 1u;
    int2* wq = s_wq[warp];
    KmerEmit em;
    em.st = &s_stage[warp]; em.cur = 0; em.pend_n = 0; em.pend_base = 0;
    const int key_shift = p.pos_bits + p.motif_bits;
    if (s_hist)
        for (int i = tid; i < 2 * p.n_motifs; i += kKmerThreads) s_hist[i] = 0;
    if (lane == 0) s_stage[warp].n = 0;
    if (tid < 3) s_tally[tid] = 0;
#pragma unroll
    for (int W = 1; W < kKmerBitmapMinW; ++W) {
        if (W > kp.max_w || kp.cls_off[W + 1] == kp.cls_off[W]) continue;
        const int nw = kmer_bm_words(W);
        for (int i = tid; i < nw; i += kKmerThreads) s_bm[kmer_sbm_off(W) + i] = (uint32_t)kp.head[kp.bm_off[W] + i];
    }
    __syncthreads();
    const uint32_t* __restrict__ wp = p.packed + row.word_off;
    const uint16_t* __restrict__ ap = p.amask + row.word_off;
    const unsigned long long orow0 = (unsigned long long)(unsigned)p.out_rows[ch.row * 2] << key_shift;
    const unsigned long long orow1 = (unsigned long long)(unsigned)p.out_rows[ch.row * 2 + 1] << key_shift;
    unsigned tallies[3] = {0, 0, 0};  // definite, near, accepted
    unsigned present = 0;
    // (the per-class offsets kp.head_off[W] / kp.bm_off[W] are read from the constant bank where they are used: as
    // arrays of registers they pushed the hot loop into local-memory spills -- 40 local accesses per 32 window starts)
#pragma unroll
    for (int W = 1; W <= kKmerMaxW; ++W)
        if (W <= kp.max_w && kp.cls_off[W + 1] > kp.cls_off[W]) present |= 1u << W;
    const long long lim64 = (long long)row.g_len - (long long)row.g_off;
    const int lim = lim64 > 0x7fffffffll ? 0x7fffffff : (int)lim64;  // windows must end inside the sequence
    const int32_t* __restrict__ hd = kp.head;
    const bool combined = kp.max_w > kKmerBitmapMinW;
    const uint32_t cb32 = (uint32_t)kp.cb_off;
    const uint32_t cb_mask = (uint32_t)((1ull << (2 * kp.max_w)) - 1ull);

    for (int pb = c0 + warp * 32; pb < c1; pb += kKmerThreads) {  // this warp's 32 window starts of the CTA's stride
        const int s = pb + lane;
        const bool active = s < c1;
        uint32_t win = 0, am = 0xffffu;
        if (active) {
            const int w = s >> 4, sh = s & 15;
            const uint32_t p0 = wp[w], p1 = wp[w + 1];
            const uint32_t a0 = ap[w], a1 = ap[w + 1];
            win = __funnelshift_r(p0, p1, 2 * sh);
            am = ((a0 | (a1 << 16)) >> sh) & 0xffffu;
        }
        // which classes does this window start take through the index?  bit W set = class present, window inside the
        // sequence (models.py:110 trim: W <= lim - s) and free of ambiguity codes (W <= index of the first flagged base)
        unsigned ok = 0, slow = 0;
        if (active) {
            const int room = lim - s;                              // widest window that still fits
            const unsigned m_trim = room >= kKmerMaxW ? 0xffffffffu : (room < 1 ? 0u : ((2u << room) - 1u));
            const int clean = am ? (__ffs((int)am) - 1) : 16;      // bases before the first IUPAC code / 'Z' / padding
            const unsigned m_clean = clean >= kKmerMaxW ? 0xffffffffu : ((2u << clean) - 1u);
            ok = present & m_trim & m_clean;
            slow = present & m_trim & ~m_clean;
        }
        // ---- class lookups: one bitmap probe (wide classes: 95 % of the k-mers have no entry, and the 1-bit tables stay
        // cache resident while the head tables do not) and one head load per class; the class loop is fully unrolled
        // narrow classes: bitmap in shared memory; wide classes: ONE probe of the combined bitmap (keyed by the widest
        // class's k-mer: set iff any wide class has a list for its prefix) decides whether their own bitmaps are read
        if (combined && (ok >> kKmerBitmapMinW)) {
            const uint32_t y = win & cb_mask;
            if (!(((uint32_t)hd[cb32 + (y >> 5)] >> (y & 31)) & 1u)) ok &= (1u << kKmerBitmapMinW) - 1u;
        }
        uint32_t bmw[kKmerMaxW + 1];
#pragma unroll
        for (int W = 1; W <= kKmerMaxW; ++W) {
            bmw[W] = (ok >> W) & 1u;
            if (bmw[W]) {
                const uint32_t x = win & (uint32_t)((1ull << (2 * W)) - 1ull);
                if (W >= kKmerBitmapMinW) bmw[W] = (uint32_t)hd[(uint32_t)kp.bm_off[W] + (x >> 5)] >> (x & 31);
                else bmw[W] = s_bm[kmer_sbm_off(W) + (x >> 5)] >> (x & 31);
            }
        }
        int heads[kKmerMaxW + 1];
#pragma unroll
        for (int W = 1; W <= kKmerMaxW; ++W) {
            heads[W] = -1;
            if (bmw[W] & 1u) heads[W] = hd[(uint32_t)kp.head_off[W] + (win & (uint32_t)((1ull << (2 * W)) - 1ull))];
        }
        // one list walk: every entry becomes a hit (or a tally, for k-mers the fp64 adjudicator rejected)
        auto walk = [&](int e, const int W, const int sp) {
            const int smax = lim - W;
            do {
                const smg_kmer_entry_t ent = kp.entries[e];
                e = ent.next;
                if (ent.col & kEntNear) ++tallies[1];
                if (ent.col & kEntRej) continue;
                if (ent.col & kEntNear) ++tallies[2]; else ++tallies[0];
                const uint32_t col = ent.col & 0x1fffffffu;
                const int m = (int)(col >> 1), sd = (int)(col & 1u);
                if (s_hist) atomicAdd(&s_hist[col], 1);
                else if (p.counts) atomicAdd(p.counts + ((size_t)ch.row * p.n_motifs + m) * 2 + sd, 1);
                if (p.hit_keys) {
                    const long long pos = sd ? (long long)(smax - sp) : (row.g_off + (long long)sp);
                    kmer_push(p, em, (sd ? orow1 : orow0) | ((unsigned long long)pos << p.motif_bits) | (unsigned long long)(unsigned)m,
                              ent.score);
                }
            } while (e >= 0);
        };
        // queue the non-empty lists: item = {entry index, (lane << 4) | W}; ballot + prefix popcount, no atomics.  A list
        // that does not fit in the queue (more than kKmerQueue non-empty lists in 32 window starts) is walked in place.
        int qn = 0;
#pragma unroll
        for (int W = 1; W <= kKmerMaxW; ++W) {
            if (!((present >> W) & 1u)) continue;
            const bool has = heads[W] >= 0;
            const unsigned bm = __ballot_sync(SMG_FULL_MASK, has);
            if (has) {
                const int qi = qn + __popc(bm & lt_mask);
                if (qi < kKmerQueue) wq[qi] = make_int2(heads[W], (lane << 4) | W);
                else walk(heads[W], W, s);
            }
            qn += __popc(bm);
        }
        qn = min(qn, kKmerQueue);
        __syncwarp();
        // ---- list walks with every lane busy: lane i takes item i, i + 32, ...
        for (int qi = lane; qi < qn; qi += 32) {
            const int2 item = wq[qi];
            walk(item.x, item.y & 15, pb + (item.y >> 4));
        }
        __syncwarp();
        // ---- generic path for windows the index cannot serve (rare)
        if (__any_sync(SMG_FULL_MASK, slow != 0u)) {
            if (slow) {
                for (int W = 1; W <= kp.max_w; ++W)
                    if ((slow >> W) & 1u) kmer_slow_window(kp, row, ch.row, W, s, em, s_hist, tallies);
            }
            __syncwarp();
        }
        if (p.hit_keys && em.st->n > kKmerStage / 2) kmer_switch(p, em, lane);
    }
    if (p.hit_keys) {  // drain: the active buffer becomes pending, then the last pending buffer is copied out
        kmer_switch(p, em, lane);
        kmer_switch(p, em, lane);
    }
    __syncthreads();
    if (s_hist && p.counts) {
        for (int i = tid; i < 2 * p.n_motifs; i += kKmerThreads) {
            const int v = s_hist[i];
            if (v) atomicAdd(p.counts + (size_t)ch.row * p.n_motifs * 2 + i, v);
        }
    }
#pragma unroll
    for (int i = 0; i < 3; ++i) tallies[i] = __reduce_add_sync(SMG_FULL_MASK, tallies[i]);
    if (lane == 0) {
        if (tallies[0]) atomicAdd(&s_tally[0], tallies[0]);
        if (tallies[1]) atomicAdd(&s_tally[1], tallies[1]);
        if (tallies[2]) atomicAdd(&s_tally[2], tallies[2]);
    }
    __syncthreads();
    if (tid == 0) {
        if (s_tally[0]) atomicAdd(&p.counters[SMG_CTR_DEFINITE], (unsigned long long)s_tally[0]);
        if (s_tally[1]) atomicAdd(&p.counters[SMG_CTR_NEAR_TC], (unsigned long long)s_tally[1]);
        if (s_tally[2]) atomicAdd(&p.counters[SMG_CTR_ADJ_ACCEPTED], (unsigned long long)s_tally[2]);
    }
}

// ---------------------------------------------------------------------------------------------------- seed scan
// One thread = one seed position q.  Clean S-mers: bitmap probe, head, list walk; every entry is a virtual PWM whose
// window starts at q - offset and becomes a 64-bit candidate for finalize_candidates_kernel (exact fp32 / fp64 scoring).
// S-mers with IUPAC codes cannot be looked up: the thread scores the soft seed of every virtual PWM directly (rare).
// A window is owned by the row segment that contains its START: the last chunk of a row also visits the o_max seed
// positions beyond its owned starts, and candidates whose start falls outside [0, n_own) are dropped.
struct SeedStage {
    unsigned long long c[256];
    int n;
};

__device__ __forceinline__ void seed_push(const SeedParams& sp, SeedStage* stg, const unsigned long long cand)
{
    const int slot = atomicAdd(&stg->n, 1);
    if (slot < 256) {
        stg->c[slot] = cand;
    } else {  // stage full (long lists / the slow path): straight to the global list
        const unsigned long long idx = atomicAdd(&sp.sp.counters[SMG_CTR_CANDIDATES], 1ull);
        if (idx < sp.cand_cap) sp.cand[idx] = cand;
    }
}

__device__ __forceinline__ void seed_flush(const SeedParams& sp, SeedStage* stg, const int lane)
{
    __syncwarp();
    const int n = min(stg->n, 256);
    if (n > 0) {
        unsigned long long base = 0;
        if (lane == 0) base = atomicAdd(&sp.sp.counters[SMG_CTR_CANDIDATES], (unsigned long long)n);
        base = __shfl_sync(SMG_FULL_MASK, base, 0);
        for (int i = lane; i < n; i += 32)
            if (base + i < sp.cand_cap) sp.cand[base + i] = stg->c[i];
    }
    __syncwarp();
    if (lane == 0) stg->n = 0;
    __syncwarp();
}

__global__ void __launch_bounds__(256) seed_scan_kernel(const __grid_constant__ SeedParams sp)
{
    const ScanParams& p = sp.sp;
    __shared__ SeedStage s_stage[8];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    SeedStage* stg = &s_stage[warp];
    if (lane == 0) stg->n = 0;
    __syncwarp();
    const smg_chunk_t ch = p.chunks[blockIdx.x];
    const smg_row_t row = p.rows[ch.row];
    const int c0 = ch.start, c1 = min(c0 + p.chunk_len, row.n_own);
    if (c1 <= c0) return;
    const int S = sp.S;
    // seed positions of this chunk; the row's last chunk also covers the seeds of windows that start just before n_own
    const int q_last = min((c1 == row.n_own) ? c1 + sp.o_max : c1, row.n_avail - S + 1);  // exclusive
    const uint32_t* __restrict__ wp = p.packed + row.word_off;
    const uint16_t* __restrict__ ap = p.amask + row.word_off;
    const uint32_t kmask = (uint32_t)((1ull << (2 * S)) - 1ull), amask_s = (1u << S) - 1u;
    const int32_t* __restrict__ hd = sp.head;
    const unsigned long long rowbits = (unsigned long long)(unsigned)ch.row << sp.cand_shift;
    const int pos_shift = p.motif_bits + 1;
    for (int qb = c0 + warp * 32; qb < q_last; qb += 256) {
        const int q = qb + lane;
        if (q < q_last) {
            const int w = q >> 4, sh = q & 15;
            const uint32_t p0 = wp[w], p1 = wp[w + 1];
            const uint32_t a0 = ap[w], a1 = ap[w + 1];
            const uint32_t x = __funnelshift_r(p0, p1, 2 * sh) & kmask;
            const uint32_t am = ((a0 | (a1 << 16)) >> sh) & amask_s;
            if (am == 0u) {
                if (((uint32_t)hd[sp.bm_off + (x >> 5)] >> (x & 31)) & 1u) {
                    int e = hd[sp.head_off + x];
                    while (e >= 0) {
                        const smg_kmer_entry_t ent = sp.entries[e];
                        e = ent.next;
                        const int v = (int)((ent.col & 0x1fffffffu) >> 1);
                        const int start = q - sp.voffset[v];
                        if (start >= 0 && start < row.n_own)
                            seed_push(sp, stg, rowbits | ((unsigned long long)(unsigned)start << pos_shift) | (unsigned long long)(unsigned)sp.vcol[v]);
                    }
                }
            } else {
                // soft one-hot seed (encode.py:11-29) of every virtual PWM, any summation order: the threshold carries a margin
                for (int v = 0; v < sp.nv; ++v) {
                    const int start = q - sp.voffset[v];
                    if (start < 0 || start >= row.n_own) continue;
                    const float* tab = sp.vtab + (size_t)v * S * 4;
                    float acc = 0.f;
                    for (int k = 0; k < S; ++k) {
                        const int c = smg_fetch_code(p.packed, p.amask, p.codes, row, (int64_t)q + k);
                        acc += smg_mix(c_emb[c][0], c_emb[c][1], c_emb[c][2], c_emb[c][3], tab[4 * k], tab[4 * k + 1], tab[4 * k + 2],
                                       tab[4 * k + 3]);
                    }
                    if (acc > sp.vthr[v])
                        seed_push(sp, stg, rowbits | ((unsigned long long)(unsigned)start << pos_shift) | (unsigned long long)(unsigned)sp.vcol[v]);
                }
            }
        }
        __syncwarp();
        if (stg->n > 128) seed_flush(sp, stg, lane);
    }
    seed_flush(sp, stg, lane);
}

}  // namespace smg

// ---------------------------------------------------------------------------------------------------- host side
int smg_kmer_fill_params(smg::KmerParams& kp, const std::vector<int>& widths, const int32_t* d_cls_motifs, int max_w)
{
    if (max_w < 1 || max_w > smg::kKmerMaxW) return 1;
    kp.cls_motifs = d_cls_motifs;
    kp.max_w = max_w;
    int off = 0;
    int64_t hoff = 0;
    for (int W = 0; W <= smg::kKmerMaxW + 1; ++W) {
        kp.cls_off[W] = off;
        kp.head_off[W] = hoff;
        if (W >= 1 && W <= smg::kKmerMaxW) {
            int n = 0;
            for (int w : widths) n += (w == W);
            off += n;
            if (n > 0 && W <= max_w) hoff += 1ll << (2 * W);
        }
    }
    for (int W = 0; W <= smg::kKmerMaxW + 1; ++W) {  // bitmaps (one bit per k-mer, at least one word) follow the heads
        kp.bm_off[W] = hoff;
        if (W >= 1 && W <= max_w && kp.cls_off[W + 1] > kp.cls_off[W]) hoff += std::max<int64_t>(1, (1ll << (2 * W)) >> 5);
    }
    kp.cb_off = hoff;  // combined bitmap of the wide classes, keyed by the max_w-mer
    return 0;
}

int64_t smg_kmer_head_len(const std::vector<int>& widths, int max_w, int64_t* n_heads)
{
    int64_t h = 0, b = 0;
    for (int W = 1; W <= std::min(max_w, smg::kKmerMaxW); ++W) {
        bool any = false;
        for (int w : widths) any = any || (w == W);
        if (any) {
            h += 1ll << (2 * W);
            b += std::max<int64_t>(1, (1ll << (2 * W)) >> 5);
        }
    }
    if (n_heads) *n_heads = h;
    if (h && max_w > smg::kKmerBitmapMinW) b += (1ll << (2 * std::min(max_w, smg::kKmerMaxW))) >> 5;  // combined bitmap
    return h ? h + b : 0;
}

cudaError_t smg_kmer_launch_build(const smg::KmerParams& kp, cudaStream_t st, uint64_t* n_launches)
{
    for (int W = 1; W <= kp.max_w; ++W) {
        const int n_cm = kp.cls_off[W + 1] - kp.cls_off[W];
        if (n_cm == 0) continue;
        const int tail = std::min(W, 3);
        const int pre = W - tail;
        // prefix levels walked depth first by one warp (4^sub leaves): as deep as leaves enough warps to fill the GPU
        const long long tiles = (n_cm + smg::kBuildTile - 1) / smg::kBuildTile;
        int sub = std::min(pre, 4);
        while (sub > 1 && (1ll << (2 * (pre - sub))) * tiles < 148 * 32) --sub;
        const long long n_sub = 1ll << (2 * (pre - sub));
        dim3 grid((unsigned)((n_sub + 3) / 4), (unsigned)tiles, 1);
        if (tail == 1) smg::kmer_build_kernel<0, 1><<<grid, 128, 0, st>>>(kp, W);
        else if (tail == 2) smg::kmer_build_kernel<0, 2><<<grid, 128, 0, st>>>(kp, W);
        else if (sub == 0) smg::kmer_build_kernel<0, 3><<<grid, 128, 0, st>>>(kp, W);
        else if (sub == 1) smg::kmer_build_kernel<1, 3><<<grid, 128, 0, st>>>(kp, W);
        else if (sub == 2) smg::kmer_build_kernel<2, 3><<<grid, 128, 0, st>>>(kp, W);
        else if (sub == 3) smg::kmer_build_kernel<3, 3><<<grid, 128, 0, st>>>(kp, W);
        else smg::kmer_build_kernel<4, 3><<<grid, 128, 0, st>>>(kp, W);
        ++*n_launches;
        cudaError_t e = cudaGetLastError();
        if (e != cudaSuccess) return e;
    }
    return cudaSuccess;
}

cudaError_t smg_seed_launch_scan(const smg::SeedParams& sp, cudaStream_t st)
{
    smg::seed_scan_kernel<<<(unsigned)sp.n_chunks, 256, 0, st>>>(sp);
    return cudaGetLastError();
}

cudaError_t smg_kmer_launch_scan(const smg::KmerParams& kp, int n_chunks, cudaStream_t st)
{
    const size_t hist_bytes = (size_t)kp.sp.n_motifs * 2 * sizeof(int);
    const int in_smem = hist_bytes <= 96 * 1024;
    if (in_smem) {  // static + dynamic shared memory may exceed the 48 KB default
        cudaError_t e = cudaFuncSetAttribute(smg::scan_kmer_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)hist_bytes);
        if (e != cudaSuccess) return e;
    }
    smg::scan_kmer_kernel<<<(unsigned)n_chunks, smg::kKmerThreads, in_smem ? hist_bytes : 0, st>>>(kp, in_smem);
    return cudaGetLastError();
}
