// Hit-list sort into the reference's np.where order (models.py:108: row-major (row, pos, motif)), bucketed.
//
// cub::DeviceRadixSort needs 5 onesweep passes over the 39..45-bit keys (128 B of DRAM traffic per hit).  The keys are
// far from random, though: a hit's high bits are (output row, position) and the scan engines emit hits chunk by chunk.
// So: (1) histogram of the hits over BUCKETS = key >> (motif_bits + B) (one output row x 2^B positions each, sized on
// the host for ~1-2 K hits), (2) exclusive scan, (3) ONE scatter pass that moves every hit into its bucket as a 32-bit
// (position-in-bucket, motif) key + score, (4) one WARP per bucket orders its hits in shared memory (counting sort over the
// 2^B positions + insertion sort by motif inside a position) and writes the final 64-bit keys and scores.  48 B of DRAM
// traffic per hit instead of 128.
// Warps aggregate their bucket-counter atomics with __match_any_sync: 32 consecutive hits of the unsorted list come
// from one or two chunks.  A bucket may hold any number of hits (only its per-position counters live in shared memory);
// *d_overflow is kept in the interface and always reports 0.
#include <string>

#include <cub/device/device_scan.cuh>

#include "common.cuh"

int smg_fail_from(int code, const std::string& msg);
void smg_count_launch();

namespace smg {

constexpr int kSortThreads = 256;  // 8 warps = 8 buckets in flight per CTA
constexpr int kSortMaxBins = 512;  // positions a bucket may span (2^9)

__global__ void __launch_bounds__(256) bucket_hist_kernel(const unsigned long long* __restrict__ keys, const unsigned long long n,
                                                          const int shift, unsigned* __restrict__ counts)
{
    const unsigned lane = threadIdx.x & 31u;
    const unsigned long long stride = (unsigned long long)gridDim.x * blockDim.x;
    const unsigned long long n_round = (n + 31ull) & ~31ull;  // whole warps stay converged for match_any
    for (unsigned long long i = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x; i < n_round; i += stride) {
        const bool have = i < n;
        const unsigned b = have ? (unsigned)(keys[i] >> shift) : 0xffffffffu;
        const unsigned m = __match_any_sync(SMG_FULL_MASK, b);
        if (have && lane == (unsigned)(__ffs((int)m) - 1)) atomicAdd(&counts[b], (unsigned)__popc(m));
    }
}

__global__ void __launch_bounds__(256) bucket_scatter_kernel(const unsigned long long* __restrict__ keys, const float* __restrict__ scores,
                                                             const unsigned long long n, const int shift, unsigned* __restrict__ cursor,
                                                             uint32_t* __restrict__ low, float* __restrict__ sc)
{
    const unsigned lane = threadIdx.x & 31u;
    const unsigned lt_mask = (1u << lane) - 1u;
    const unsigned long long low_mask = (1ull << shift) - 1ull;
    // kU independent elements per thread and pass: the cursor atomics return a value, so with one element per pass
    // every warp sat out a full L2 round trip per 32 hits (17 % issue utilisation); now kU round trips overlap
    constexpr int kU = 4;
    const unsigned long long stride = (unsigned long long)gridDim.x * blockDim.x * kU;
    for (unsigned long long blk = (unsigned long long)blockIdx.x * blockDim.x * kU; blk < n; blk += stride) {  // CTA-uniform trips
        const unsigned long long i0 = blk + threadIdx.x;
        unsigned long long k[kU];
        float v[kU];
        unsigned base[kU], m[kU];
        bool have[kU];
#pragma unroll
        for (int u = 0; u < kU; ++u) {
            const unsigned long long i = i0 + (unsigned long long)u * blockDim.x;
            have[u] = i < n;
            k[u] = have[u] ? keys[i] : 0ull;
            v[u] = have[u] ? scores[i] : 0.f;
        }
#pragma unroll
        for (int u = 0; u < kU; ++u) {
            const unsigned b = have[u] ? (unsigned)(k[u] >> shift) : 0xffffffffu;
            m[u] = __match_any_sync(SMG_FULL_MASK, b);
            base[u] = 0;
            if (have[u] && (int)lane == __ffs((int)m[u]) - 1) base[u] = atomicAdd(&cursor[b], (unsigned)__popc(m[u]));
        }
#pragma unroll
        for (int u = 0; u < kU; ++u) {
            const unsigned bs = __shfl_sync(SMG_FULL_MASK, base[u], __ffs((int)m[u]) - 1);
            if (have[u]) {
                const unsigned slot = bs + (unsigned)__popc(m[u] & lt_mask);
                low[slot] = (uint32_t)(k[u] & low_mask);
                sc[slot] = v[u];
            }
        }
    }
}

// offsets[b] .. offsets[b + 1]: the bucket's hits in `low` / `sc`; output at the same offsets of keys_out / scores_out.
// ONE WARP PER BUCKET (no block-wide barriers, 8 independent warps per CTA, 2 KB of shared memory each): a bucket spans
// 2^pb <= 512 positions of one output row, so its hits are ordered with one counting sort over the position (shared-memory
// histogram, warp scan, scatter straight to the final arrays) followed by an insertion sort of the few hits that share a
// position (ordered by motif, in place in global memory): ~30 instructions per hit, against ~240 for a block radix sort.
__global__ void __launch_bounds__(kSortThreads) bucket_sort_kernel(const unsigned* __restrict__ offsets, const unsigned n_buckets,
                                                                   const uint32_t* __restrict__ low, const float* __restrict__ sc,
                                                                   const int shift, const int motif_bits,
                                                                   unsigned long long* __restrict__ keys_out,
                                                                   float* __restrict__ scores_out)
{
    constexpr int kWarps = kSortThreads / 32;
    __shared__ unsigned s_cnt[kWarps][kSortMaxBins + 32];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    unsigned* cnt = s_cnt[warp];
    const int n_bins = 1 << (shift - motif_bits);
    constexpr int kPer = kSortMaxBins / 32;  // bins per lane in the scan
    for (unsigned b = blockIdx.x * kWarps + warp; b < n_buckets; b += gridDim.x * kWarps) {
        const unsigned o0 = offsets[b], n = offsets[b + 1] - o0;
        if (n == 0) continue;
        const unsigned long long hi_bits = (unsigned long long)b << shift;
        if (n == 1) {
            if (lane == 0) {
                keys_out[o0] = hi_bits | low[o0];
                scores_out[o0] = sc[o0];
            }
            continue;
        }
        for (int i = lane; i < kSortMaxBins + 32; i += 32) cnt[i] = 0u;
        __syncwarp();
        for (unsigned j = lane; j < n; j += 32) atomicAdd(&cnt[low[o0 + j] >> motif_bits], 1u);
        __syncwarp();
        // exclusive scan of the per-position counts: kPer consecutive bins per lane, then a warp scan of the lane totals
        unsigned c[kPer], tot = 0;
#pragma unroll
        for (int i = 0; i < kPer; ++i) {
            c[i] = cnt[lane * kPer + i];
            tot += c[i];
        }
        unsigned incl = tot;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const unsigned t = __shfl_up_sync(SMG_FULL_MASK, incl, d);
            if (lane >= d) incl += t;
        }
        unsigned run = incl - tot;
        __syncwarp();
#pragma unroll
        for (int i = 0; i < kPer; ++i) {
            cnt[lane * kPer + i] = run;  // start of the bin; advanced by the scatter below, so it ends as the bin's END
            run += c[i];
        }
        __syncwarp();
        for (unsigned j = lane; j < n; j += 32) {
            const uint32_t k = low[o0 + j];
            const unsigned slot = atomicAdd(&cnt[k >> motif_bits], 1u);
            keys_out[o0 + slot] = hi_bits | k;
            scores_out[o0 + slot] = sc[o0 + j];
        }
        __syncwarp();  // orders the stores above before the reads below (same warp)
        // hits sharing a position: order by motif (a handful at most).  cnt[bin] is now the END of bin, cnt[bin - 1] its start.
        for (int bin = lane; bin < n_bins; bin += 32) {
            const unsigned hi = cnt[bin], lo = bin ? cnt[bin - 1] : 0u;
            if (hi - lo < 2) continue;
            unsigned long long* kp = keys_out + o0;
            float* sp = scores_out + o0;
            for (unsigned x = lo + 1; x < hi; ++x) {
                const unsigned long long kk = kp[x];
                const float vv = sp[x];
                unsigned y = x;
                while (y > lo && kp[y - 1] > kk) {
                    kp[y] = kp[y - 1];
                    sp[y] = sp[y - 1];
                    --y;
                }
                kp[y] = kk;
                sp[y] = vv;
            }
        }
        __syncwarp();
    }
}

}  // namespace smg

extern "C" {

int smg_sort_hits_bucketed(const uint64_t* d_keys_in, const float* d_scores_in, uint64_t* d_keys_out, float* d_scores_out, uint64_t n,
                           int key_bits, int motif_bits, int bucket_shift, void* d_temp, uint64_t* temp_bytes, uint32_t* d_overflow,
                           void* stream)
{
    if (!temp_bytes || key_bits <= 0 || key_bits > 63 || bucket_shift <= 0 || bucket_shift > 31 || bucket_shift >= key_bits || motif_bits <= 0 || motif_bits > bucket_shift ||
        bucket_shift - motif_bits > 9 ||
        key_bits - bucket_shift > 28 || n >= (1ull << 32))
        return smg_fail_from(SMG_ERR_INVALID, "smg_sort_hits_bucketed: bad argument");
    const uint64_t n_buckets = 1ull << (key_bits - bucket_shift);
    size_t scan_bytes = 0;
    cub::DeviceScan::ExclusiveSum(nullptr, scan_bytes, (unsigned*)nullptr, (unsigned*)nullptr, (int)(n_buckets + 1));
    const size_t a256 = 256;
    auto up = [&](size_t x) { return (x + a256 - 1) / a256 * a256; };
    const size_t off_counts = 0, off_offsets = up((n_buckets + 1) * 4), off_low = off_offsets + up((n_buckets + 1) * 4);
    const size_t off_sc = off_low + up((size_t)n * 4), off_scan = off_sc + up((size_t)n * 4);
    const size_t need = off_scan + up(scan_bytes);
    if (!d_temp) {
        *temp_bytes = (uint64_t)need;
        return SMG_OK;
    }
    if (*temp_bytes < need || !d_keys_in || !d_scores_in || !d_keys_out || !d_scores_out || !d_overflow)
        return smg_fail_from(SMG_ERR_INVALID, "smg_sort_hits_bucketed: buffers missing or temporary storage too small");
    if (n == 0) return SMG_OK;
    cudaStream_t st = (cudaStream_t)stream;
    char* t = (char*)d_temp;
    unsigned* counts = (unsigned*)(t + off_counts);
    unsigned* offsets = (unsigned*)(t + off_offsets);
    uint32_t* low = (uint32_t*)(t + off_low);
    float* sc = (float*)(t + off_sc);
    cudaError_t e = cudaMemsetAsync(counts, 0, (n_buckets + 1) * 4, st);
    if (e == cudaSuccess) e = cudaMemsetAsync(d_overflow, 0, 4, st);
    if (e != cudaSuccess) return smg_fail_from(SMG_ERR_CUDA, std::string("smg_sort_hits_bucketed: ") + cudaGetErrorString(e));
    const unsigned blocks = (unsigned)std::min<uint64_t>((n + 255) / 256, 148 * 16);
    smg::bucket_hist_kernel<<<blocks, 256, 0, st>>>(reinterpret_cast<const unsigned long long*>(d_keys_in), n, bucket_shift, counts);
    smg_count_launch();
    e = cub::DeviceScan::ExclusiveSum(t + off_scan, scan_bytes, counts, offsets, (int)(n_buckets + 1), st);
    if (e != cudaSuccess) return smg_fail_from(SMG_ERR_CUDA, std::string("cub::DeviceScan: ") + cudaGetErrorString(e));
    // the scatter advances a copy of the offsets (the counts array is reused as the cursor)
    e = cudaMemcpyAsync(counts, offsets, (n_buckets + 1) * 4, cudaMemcpyDeviceToDevice, st);
    if (e != cudaSuccess) return smg_fail_from(SMG_ERR_CUDA, std::string("smg_sort_hits_bucketed: ") + cudaGetErrorString(e));
    smg::bucket_scatter_kernel<<<blocks, 256, 0, st>>>(reinterpret_cast<const unsigned long long*>(d_keys_in), d_scores_in, n,
                                                       bucket_shift, counts, low, sc);
    smg_count_launch();
    const unsigned sblocks = (unsigned)std::min<uint64_t>((n_buckets + 7) / 8, 148 * 64);
    smg::bucket_sort_kernel<<<sblocks, smg::kSortThreads, 0, st>>>(offsets, (unsigned)n_buckets, low, sc, bucket_shift, motif_bits,
                                                                   reinterpret_cast<unsigned long long*>(d_keys_out), d_scores_out);
    smg_count_launch();
    e = cudaGetLastError();
    if (e != cudaSuccess) return smg_fail_from(SMG_ERR_CUDA, std::string("bucket sort kernels: ") + cudaGetErrorString(e));
    return SMG_OK;
}

}  // extern "C"
