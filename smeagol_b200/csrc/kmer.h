// Host-side interface of the k-mer index engine (kmer.cu), used by the C-ABI entry points in api.cu.
#pragma once
#include <cstdint>
#include <vector>
#include <cuda_runtime.h>

#include "common.cuh"

namespace smg {

constexpr int kKmerMaxW = 12;
constexpr int kKmerStage = 128;          // hits per buffer of a warp's double-buffered stage
constexpr int kKmerQueue = 160;          // work items (non-empty lists) a warp queues per 32 window starts
constexpr int kKmerBitmapMinW = 9;       // classes at least this wide are looked up through a 1-bit-per-k-mer bitmap first
constexpr int kKmerThreads = 256;
constexpr uint32_t kEntNear = 1u << 30;  // decided by the fp64 adjudicator
constexpr uint32_t kEntRej = 1u << 29;   // inside the band, rejected in fp64 (kept so that the scan can count it)

// bitmap words of class W, and where the narrow classes' bitmaps sit in the scan kernel's shared memory
__host__ __device__ constexpr int kmer_bm_words(const int W) { return (1 << (2 * W)) >= 32 ? (1 << (2 * W)) / 32 : 1; }
__host__ __device__ constexpr int kmer_sbm_off(const int W)
{
    int o = 0;
    for (int w = 1; w < W; ++w) o += kmer_bm_words(w);
    return o;
}
constexpr int kKmerSmemBitmapWords = kmer_sbm_off(kKmerBitmapMinW);

struct KmerParams {
    ScanParams sp;
    const int32_t* cls_motifs;        // motif ids sorted by width
    int32_t cls_off[kKmerMaxW + 2];   // class W = motifs cls_motifs[cls_off[W] .. cls_off[W + 1])
    int64_t head_off[kKmerMaxW + 2];  // first head slot of class W
    int64_t bm_off[kKmerMaxW + 2];    // first bitmap word of class W (in the same int32 buffer, after all heads)
    int64_t cb_off;                   // combined bitmap of the classes >= kKmerBitmapMinW (max_w > kKmerBitmapMinW only)
    int32_t max_w;                    // classes 1 .. max_w
    int32_t* head;
    smg_kmer_entry_t* entries;
    unsigned long long entry_cap;
    const double* thr64;
};

// ---- seed engine (wide PWMs through the k-mer machinery): a window of a PWM of width W > S can only be a hit if its
// most selective S consecutive positions (the SEED, at offset o of the window) score more than thr - (maximum of the
// rest); every (PWM, strand) becomes a VIRTUAL forward-only PWM of width S in a one-class k-mer index built per call.
constexpr int kSeedMaxRest = 14;  // a PWM is seeded when W <= S + kSeedMaxRest
struct SeedParams {
    ScanParams sp;                 // the REAL PWM set (rows, chunks, candidate key layout)
    const int32_t* head;           // one-class index of the virtual PWMs
    int64_t head_off, bm_off;      // heads / bitmap of class S inside `head`
    const smg_kmer_entry_t* entries;
    const int32_t* vcol;           // [nv] real motif * 2 + strand
    const int32_t* voffset;        // [nv] offset of the seed inside the window
    const float* vtab;             // [nv][S][4] seed tables (strand already applied)
    const float* vthr;             // [nv] seed thresholds of this call
    int32_t nv, S, o_max, n_chunks;
    int32_t cand_shift;            // candidate = row << cand_shift | start << (motif_bits + 1) | motif << 1 | strand
    unsigned long long* cand;
    unsigned long long cand_cap;
};

}  // namespace smg

cudaError_t smg_seed_launch_scan(const smg::SeedParams& sp, cudaStream_t st);
int smg_kmer_fill_params(smg::KmerParams& kp, const std::vector<int>& widths, const int32_t* d_cls_motifs, int max_w);
// int32 slots of the index buffer: list heads of every class, then the bitmaps; *n_heads = the head part
int64_t smg_kmer_head_len(const std::vector<int>& widths, int max_w, int64_t* n_heads = nullptr);
cudaError_t smg_kmer_launch_build(const smg::KmerParams& kp, cudaStream_t st, uint64_t* n_launches);
cudaError_t smg_kmer_launch_scan(const smg::KmerParams& kp, int n_chunks, cudaStream_t st);
