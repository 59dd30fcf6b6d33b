// tcgen05 / TMEM PREFILTER for the PWM scan -- the "im2col tensor-core variant" of north_star, built so that the im2col
// matrix never exists and the result stays exact (same contract as scan_hmma.cuh: it only has to emit a SUPERSET of the
// true hits as 64-bit candidates; finalize_candidates_kernel re-scores them in fp32 in the fixed order and adjudicates).
//
//   score[col][p] = sum_{k < W, b < 4} T_col[k][b] * X[p + k][b]              (models.py:50-68: Embedding + Conv1D)
//
// is the GEMM  D[128 columns][256 window starts] += A[128][16] * B[256][16]^T  per tcgen05.mma (kind::f16, fp32
// accumulate in TMEM), KS = ceil(W / 4) instructions per accumulator:
//   * A = the fp16 log-odds tables of 128 (PWM, strand) columns, K-major, resident in shared memory for the whole kernel
//     (loaded once per CTA with cp.async.bulk); the reverse-complement strand is a column with the table
//     Trc[k][b] = T[W-1-k][3-b] (matrices.py:224-243), so one forward one-hot operand serves both strands;
//   * B = the soft one-hot of the sequence, 4 fp16 (8 bytes) per base, written ONCE per tile as a linear strip.  Row p of
//     the im2col matrix is the strip shifted by 8*p bytes (Toeplitz), so the shared-memory matrix descriptor simply
//     OVERLAPS its core matrices: 8-row groups 128 B apart (SBO), the two K halves of an instruction 16 B apart (LBO),
//     rows 16 B apart = every second window start; a second strip shifted by one base serves the odd starts.  A tile of
//     512 window starts therefore costs 8.7 KB of shared memory and ~550 16-byte stores instead of a 512 x 4W matrix
//     (checked against a host contraction in tools/tcgen05_probe.cu, profiles/r02_tcgen05_probe.txt);
//   * D lives in TMEM (2 x 256 columns, double buffered): the MMA of the next accumulator runs while 16 epilogue warps
//     read the current one (tcgen05.ld 32x32b.x64: thread = one (PWM, strand) column, 64 window starts in registers),
//     release the buffer, reduce with 3-input max and compare with the thread's prefilter threshold.  Groups of 16
//     cells that hold a candidate are TRANSPOSED through a small per-warp scratch in shared memory: the flagged lanes
//     store their 16 cells, then the warp re-reads them with lane = cell (two flagged lanes per ballot) and writes the
//     survivors as 64-bit candidates (row, start, motif, strand) straight into the global list, in blocks of 128 slots
//     that the warp reserves one block AHEAD (the atomic's latency never stalls it; the unused tail of the last blocks
//     is padded with an invalid marker that the exact pass skips).  The epilogue warps synchronise on every
//     accumulator, so their per-step work must be short and even: a version that queued the groups and drained the
//     queue every few steps ran 3.3x slower at hit density 1.5e-3 (one draining warp stalls all 16), and two
//     asynchronous consumer warps could not keep up (a single warp's dependent chain: 275 cycles per pair of groups).
//
// Warp roles (576 threads, 1 CTA per SM, persistent over chunks): warps 0..15 = epilogue, warp 16 = strip producer,
// warp 17 = MMA issuer (one elected lane).  The two service warps carry the HIGHEST warp ids on purpose: the SMSP
// arbiter prefers the highest eligible warp id, so the single thread that feeds the tensor pipe is never starved by
// the four epilogue warps that share its scheduler while they grind through the max tree.  Synchronisation is mbarrier-only: xfull/xempty (strips), tfull/tempty (TMEM).
#pragma once
#include "common.cuh"
#include "scan_tc5_params.cuh"
#include "tc5.cuh"

namespace smg {

struct Tc5Misc {
    uint64_t xfull[2], xempty[2], tfull[2], tempty[2], wbar;
    uint32_t tmem_base;
    int32_t n_blk;
    int32_t blk_ks[kTc5MaxBlk];
    int32_t blk_off[kTc5MaxBlk];
};
static_assert(sizeof(Tc5Misc) <= 1024, "misc block");

__device__ __forceinline__ float tc5_max3(const float a, const float b, const float c)
{
    float d;
    asm("max.f32 %0, %1, %2, %3;" : "=f"(d) : "f"(a), "f"(b), "f"(c));
    return d;
}
__device__ __forceinline__ uint32_t tc5_shl_one(const uint32_t s)  // 0x3c00 << s, 0 for s >= 32 (PTX shl clamps)
{
    uint32_t d;
    asm("shl.b32 %0, 0x3c00, %1;" : "=r"(d) : "r"(s));
    return d;
}

// fp16 pairs {e[c][0], e[c][1]}, {e[c][2], e[c][3]} of the soft one-hot table (encode.py:11-29)
static __constant__ uint32_t c_tc5_emb16[17][2] = {
    {0x00000000u, 0x00000000u}, {0x00003c00u, 0x00000000u}, {0x3c000000u, 0x00000000u}, {0x00000000u, 0x00003c00u},
    {0x00000000u, 0x3c000000u}, {0x00000000u, 0x3c000000u}, {0x34003400u, 0x34003400u}, {0x00003800u, 0x38000000u},
    {0x38000000u, 0x00003800u}, {0x38003800u, 0x00000000u}, {0x00000000u, 0x38003800u}, {0x00003800u, 0x00003800u},
    {0x38000000u, 0x38000000u}, {0x35550000u, 0x35553555u}, {0x00003555u, 0x35553555u}, {0x35553555u, 0x35550000u},
    {0x35553555u, 0x00003555u},
};

// ---- producer: the two one-hot strips of the tile whose first window start is row-local base t0 (a multiple of 16).
// Strip e, entry q (8 bytes) = soft one-hot of base t0 + e + q.  Lane j writes the 16 bases of packed word j.
__device__ __forceinline__ void tc5_build_strips(const ScanParams& p, const smg_row_t& row, const int t0, uint8_t* strip0,
                                                 uint8_t* strip1, const int lane)
{
    for (int j = lane; j < kTc5StripBases / 16; j += 32) {
        const int b0 = t0 + 16 * j;  // first base of this word (row-local)
        bool fast = false;
        if (b0 + 17 <= row.n_avail) {
            const long long w = row.word_off + (b0 >> 4);
            const uint32_t am = (uint32_t)p.amask[w] | (((uint32_t)p.amask[w + 1] & 1u) << 16);
            if (am == 0u) {
                fast = true;
                const uint32_t pw = p.packed[w], pn = p.packed[w + 1];
                uint2 v[17];
#pragma unroll
                for (int i = 0; i < 17; ++i) {
                    uint32_t d;  // 16 * (2-bit base code): selects which fp16 of the four is 1.0
                    if (i == 16) d = (pn << 4) & 0x30u;
                    else if (i < 2) d = (pw << (4 - 2 * i)) & 0x30u;
                    else d = (pw >> (2 * i - 4)) & 0x30u;
                    v[i].x = tc5_shl_one(d);
                    v[i].y = tc5_shl_one(d - 32u);
                }
                uint4* d0 = reinterpret_cast<uint4*>(strip0 + 128 * j);
                uint4* d1 = reinterpret_cast<uint4*>(strip1 + 128 * j);
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                    d0[i] = make_uint4(v[2 * i].x, v[2 * i].y, v[2 * i + 1].x, v[2 * i + 1].y);
                    d1[i] = make_uint4(v[2 * i + 1].x, v[2 * i + 1].y, v[2 * i + 2].x, v[2 * i + 2].y);
                }
            }
        }
        if (!fast) {  // IUPAC codes, 'Z', or the end of the row (bases at or beyond n_avail are 'Z' = zeros)
            uint2* e0 = reinterpret_cast<uint2*>(strip0) + 16 * j;
            uint2* e1 = reinterpret_cast<uint2*>(strip1) + 16 * j - 1;
#pragma unroll 1
            for (int i = 0; i < 17; ++i) {
                const int c = (b0 + i < row.n_avail) ? smg_fetch_code(p.packed, p.amask, p.codes, row, (int64_t)b0 + i) : 0;
                const uint2 v = make_uint2(c_tc5_emb16[c][0], c_tc5_emb16[c][1]);
                if (i < 16) e0[i] = v;
                if (i >= 1) e1[i] = v;
            }
        }
    }
}

__global__ void __launch_bounds__(kTc5Threads, 1) scan_tc5_kernel(const __grid_constant__ Tc5Params hp)
{
    extern __shared__ __align__(1024) uint8_t smem[];
    const ScanParams& p = hp.sp;
    Tc5Misc& ms = *reinterpret_cast<Tc5Misc*>(smem + kTc5OffMisc);
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int cg = blockIdx.y;
    const int first_blk = hp.cg_first_blk[cg];
    const int n_blk = hp.cg_first_blk[cg + 1] - first_blk;

    // ---- one-time setup: barriers, block tables, TMEM, weight image (bulk async copy, completes on wbar)
    if (tid < n_blk) {
        ms.blk_ks[tid] = hp.blk_ks[first_blk + tid];
        ms.blk_off[tid] = hp.blk_off[first_blk + tid];
    }
    if (tid == 0) {
        for (int i = 0; i < 2; ++i) {
            tc5::mbar_init(&ms.xfull[i], 1);
            tc5::mbar_init(&ms.xempty[i], 1);
            tc5::mbar_init(&ms.tfull[i], 1);
            tc5::mbar_init(&ms.tempty[i], kTc5EpiWarps);
        }
        tc5::mbar_init(&ms.wbar, 1);
        tc5::mbar_fence_init();
    }
    if (warp == kTc5ProdWarp) tc5::tmem_alloc(&ms.tmem_base, 512);
    tc5::fence_before_sync();
    __syncthreads();
    tc5::fence_after_sync();
    const uint32_t tb = ms.tmem_base;
    if (tid == 32 * kTc5MmaWarp) {  // the MMA thread owns the weight image: it issues the copy and is the only one to wait for it
        uint32_t total = 0;
        for (int b = 0; b < n_blk; ++b) total += (uint32_t)ms.blk_ks[b] * kTc5UnitBytes;
        asm volatile("{\n\t.reg .b64 st;\n\tmbarrier.arrive.expect_tx.shared::cta.b64 st, [%0], %1;\n\t}" ::"r"(tc5::smem_u32(&ms.wbar)),
                     "r"(total)
                     : "memory");
        const uint8_t* src = hp.wimg + hp.cg_img_off[cg];
        for (int b = 0; b < n_blk; ++b) {
            const uint32_t bytes = (uint32_t)ms.blk_ks[b] * kTc5UnitBytes;
            asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                             tc5::smem_u32(smem + kTc5OffW + ms.blk_off[b])),
                         "l"(src + ms.blk_off[b]), "r"(bytes), "r"(tc5::smem_u32(&ms.wbar))
                         : "memory");
        }
    }

    if (warp == kTc5ProdWarp) {
        // =========================================================================== strip producer
        int tcount = 0;
        for (int ci = blockIdx.x; ci < hp.n_chunks; ci += gridDim.x) {
            const smg_chunk_t ch = p.chunks[ci];
            const smg_row_t row = p.rows[ch.row];
            const int c0 = ch.start, c1 = min(c0 + p.chunk_len, row.n_own);
            for (int t0 = c0; t0 < c1; t0 += kTc5Tile, ++tcount) {
                const int stage = tcount & 1;
                if (tcount >= 2) tc5::mbar_wait(&ms.xempty[stage], (uint32_t)(((tcount >> 1) - 1) & 1));
                tc5_build_strips(p, row, t0, smem + kTc5OffStrip + (stage * 2) * kTc5StripBytes,
                                 smem + kTc5OffStrip + (stage * 2 + 1) * kTc5StripBytes, lane);
                tc5::fence_proxy_async();
                __syncwarp();
                if (lane == 0) tc5::mbar_arrive(&ms.xfull[stage]);
            }
        }
    } else if (warp == kTc5MmaWarp) {
        // =========================================================================== MMA issuer (one thread)
        if (lane == 0) {
            tc5::mbar_wait(&ms.wbar, 0);
            const uint32_t idesc = tc5::make_idesc_f16_f32(128, 256);
            const uint32_t wbase = tc5::smem_u32(smem + kTc5OffW);
            const uint32_t sbase = tc5::smem_u32(smem + kTc5OffStrip);
            int tcount = 0, gcount = 0;
            for (int ci = blockIdx.x; ci < hp.n_chunks; ci += gridDim.x) {
                const smg_chunk_t ch = p.chunks[ci];
                const int c0 = ch.start, c1 = min(c0 + p.chunk_len, p.rows[ch.row].n_own);
                for (int t0 = c0; t0 < c1; t0 += kTc5Tile, ++tcount) {
                    const int stage = tcount & 1;
                    tc5::mbar_wait(&ms.xfull[stage], (uint32_t)((tcount >> 1) & 1));
                    tc5::fence_after_sync();
                    for (int b = 0; b < n_blk; ++b) {
                        const int ks = ms.blk_ks[b];
                        const uint32_t wa = wbase + (uint32_t)ms.blk_off[b];
                        for (int e = 0; e < 2; ++e, ++gcount) {
                            const int buf = gcount & 1, k = gcount >> 1;
                            if (k >= 1) {
                                tc5::mbar_wait(&ms.tempty[buf], (uint32_t)((k - 1) & 1));
                                tc5::fence_after_sync();
                            }
                            const uint32_t sa = sbase + (uint32_t)(stage * 2 + e) * kTc5StripBytes;
                            for (int s = 0; s < ks; ++s)
                                tc5::mma_f16_ss(tb + (uint32_t)buf * 256u, tc5::make_smem_desc(wa + s * kTc5UnitBytes, 2048, 128),
                                                tc5::make_smem_desc(sa + s * 32, 16, 128), idesc, s > 0 ? 1u : 0u);
                            tc5::mma_commit(&ms.tfull[buf]);
                        }
                    }
                    tc5::mma_commit(&ms.xempty[stage]);  // strips of this stage are free once these MMAs have read them
                }
            }
            // every asynchronous arrival must have landed before the CTA may exit
            if (tcount >= 1) tc5::mbar_wait(&ms.xempty[(tcount - 1) & 1], (uint32_t)(((tcount - 1) >> 1) & 1));
            if (tcount >= 2) tc5::mbar_wait(&ms.xempty[(tcount - 2) & 1], (uint32_t)(((tcount - 2) >> 1) & 1));
        }
    } else {
        // =========================================================================== epilogue warps
        const int ew = warp;
        const int quad = warp & 3;      // TMEM lanes 32 * (warp % 4) .. + 31 are the ones this warp may read
        const int colchunk = ew >> 2;   // 64 of the accumulator's 256 columns (window starts)
        uint8_t* scratch = smem + kTc5OffScratch + ew * kTc5ScratchBytes;
        const unsigned lt_mask = (1u << lane) - 1u;
        const int cell = lane & 15, sub = lane >> 4;
        const int pos_shift = p.motif_bits + 1;
        const uint32_t taddr0 = tb + ((uint32_t)(quad * 32) << 16) + (uint32_t)(colchunk * 64);
        const size_t col0 = (size_t)first_blk * 128 + quad * 32 + lane;
        // candidate slots: the block being filled [cbase, cbase + kTc5CandBlock) and the next one, reserved in advance
        unsigned long long cbase = 0, nbase = 0;
        if (lane == 0) {
            cbase = atomicAdd(&p.counters[SMG_CTR_CANDIDATES], 2ull * kTc5CandBlock);
            nbase = cbase + kTc5CandBlock;
        }
        cbase = __shfl_sync(SMG_FULL_MASK, cbase, 0);
        nbase = __shfl_sync(SMG_FULL_MASK, nbase, 0);
        int cused = 0;
        int gcount = 0;
        for (int ci = blockIdx.x; ci < hp.n_chunks; ci += gridDim.x) {
            const smg_chunk_t ch = p.chunks[ci];
            const int c0 = ch.start, c1 = min(c0 + p.chunk_len, p.rows[ch.row].n_own);
            const unsigned long long rowbits = (unsigned long long)(unsigned)ch.row << hp.cand_shift;
            for (int t0 = c0; t0 < c1; t0 += kTc5Tile) {
                for (int b = 0; b < n_blk; ++b) {
                    const float thr = hp.thr_pre_col[col0 + b * 128];
                    for (int e = 0; e < 2; ++e, ++gcount) {
                        const int buf = gcount & 1;
                        tc5::mbar_wait(&ms.tfull[buf], (uint32_t)((gcount >> 1) & 1));
                        tc5::fence_after_sync();
                        // both halves of the 64 cells are requested back to back (one TMEM round trip instead of two), then the
                        // accumulator is handed back to the MMA issuer before any arithmetic
                        uint32_t r[2][32];
                        tc5::tmem_ld_32x32b_x32(r[0], taddr0 + (uint32_t)buf * 256u);
                        tc5::tmem_ld_32x32b_x32(r[1], taddr0 + (uint32_t)buf * 256u + 32u);
                        tc5::tmem_ld_wait();
                        tc5::fence_before_sync();
                        __syncwarp();
                        if (lane == 0) tc5::mbar_arrive(&ms.tempty[buf]);
#pragma unroll
                        for (int half = 0; half < 2; ++half) {
                            // common path: the maximum of each group of 16 cells (3-input max), ONE vote for the 32 cells
                            const float* v = reinterpret_cast<const float*>(r[half]);
                            float gm[2];
#pragma unroll
                            for (int gh = 0; gh < 2; ++gh) {
                                const float* u = v + 16 * gh;
                                const float m0 = tc5_max3(u[0], u[1], u[2]), m1 = tc5_max3(u[3], u[4], u[5]);
                                const float m2 = tc5_max3(u[6], u[7], u[8]), m3 = tc5_max3(u[9], u[10], u[11]);
                                const float m4 = tc5_max3(u[12], u[13], u[14]);
                                gm[gh] = fmaxf(tc5_max3(m0, m1, m2), tc5_max3(m3, m4, u[15]));
                            }
                            if (!__any_sync(SMG_FULL_MASK, fmaxf(gm[0], gm[1]) > thr)) continue;
                            // ---- rare path: transpose the flagged groups through the scratch and extract the candidates
#pragma unroll
                            for (int gh = 0; gh < 2; ++gh) {
                                const float* u = v + 16 * gh;
                                const bool hot = gm[gh] > thr;
                                const unsigned bm = __ballot_sync(SMG_FULL_MASK, hot);
                                if (!bm) continue;
                                const int k = __popc(bm);
                                if (hot) {
                                    float4* dst = reinterpret_cast<float4*>(scratch + __popc(bm & lt_mask) * kTc5EntryBytes);
                                    dst[0] = make_float4(u[0], u[1], u[2], u[3]);
                                    dst[1] = make_float4(u[4], u[5], u[6], u[7]);
                                    dst[2] = make_float4(u[8], u[9], u[10], u[11]);
                                    dst[3] = make_float4(u[12], u[13], u[14], u[15]);
                                    reinterpret_cast<uint2*>(dst)[8] = make_uint2(__float_as_uint(thr), (uint32_t)hp.col_id[col0 + b * 128]);
                                }
                                __syncwarp();
                                const int s0 = t0 + 2 * (colchunk * 64 + (half * 2 + gh) * 16) + e;  // window start of cell 0
                                const bool cell_ok = s0 + 2 * cell < c1;                           // starts owned by this chunk
                                for (int j = 0; j < k; j += 2) {  // lane = cell, two flagged lanes per ballot
                                    const uint8_t* ent = scratch + (j + sub) * kTc5EntryBytes;
                                    const uint2 tc = reinterpret_cast<const uint2*>(ent)[8];
                                    const float val = reinterpret_cast<const float*>(ent)[cell];
                                    const bool is = cell_ok && (j + sub < k) && (val > __uint_as_float(tc.x));
                                    const unsigned m = __ballot_sync(SMG_FULL_MASK, is);
                                    if (m) {
                                        if (is) {
                                            const int i = cused + __popc(m & lt_mask);
                                            const unsigned long long idx = i < kTc5CandBlock ? cbase + i : nbase + (i - kTc5CandBlock);
                                            if (idx < hp.cand_cap)
                                                hp.cand[idx] = rowbits | ((unsigned long long)(unsigned)(s0 + 2 * cell) << pos_shift) |
                                                               (unsigned long long)tc.y;
                                        }
                                        cused += __popc(m);
                                        if (cused >= kTc5CandBlock) {  // move on to the reserved block, reserve another one
                                            cused -= kTc5CandBlock;
                                            cbase = nbase;
                                            unsigned long long nb = 0;
                                            if (lane == 0) nb = atomicAdd(&p.counters[SMG_CTR_CANDIDATES], (unsigned long long)kTc5CandBlock);
                                            nbase = __shfl_sync(SMG_FULL_MASK, nb, 0);
                                        }
                                    }
                                }
                                __syncwarp();  // the scratch is rewritten by the next flagged group
                            }
                        }
                    }
                }
            }
        }
        // pad what was reserved and not used (the exact pass walks every reserved slot)
        for (int i = cused + lane; i < 2 * kTc5CandBlock; i += 32) {
            const unsigned long long idx = i < kTc5CandBlock ? cbase + i : nbase + (i - kTc5CandBlock);
            if (idx < hp.cand_cap) hp.cand[idx] = SMG_CAND_INVALID;
        }
    }
    tc5::fence_before_sync();
    __syncthreads();
    if (warp == kTc5ProdWarp) tc5::tmem_dealloc(tb, 512);
}

inline size_t tc5_smem_bytes(const int max_image_bytes) { return (size_t)kTc5OffW + (size_t)max_image_bytes; }

}  // namespace smg
