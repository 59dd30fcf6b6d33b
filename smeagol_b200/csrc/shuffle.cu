// k-mer preserving shuffles on the device (the step before the background scan of config c3).
//
// Replaces utils.shuffle_records (smeagol/utils.py:36-83) for the per-base work: simK = 2 calls
// deeplift.dinuc_shuffle (third-party, not under /root/reference; published algorithm: for every symbol the list of its
// successors is permuted except for its last element, then the sequence is re-walked from the first base consuming the
// lists in order -- an Eulerian path, so every dinucleotide count, the first and the last base are preserved);
// simK = 1 is a uniform permutation of the bases (random.shuffle).  The random streams differ from numpy's /
// Python's (counter-based splitmix64 here), so parity is on the PROPERTIES and on the distribution, not on the bytes.
//
// One warp per copy.  The successor lists of the ORIGINAL sequence are bucketed once (host, O(L)); every copy gets
// them in shared memory (global scratch when a sequence does not fit), lanes shuffle one symbol's list each
// (Fisher-Yates), lane 0 re-walks.  All copies of a genome are produced in one launch and never leave the device.
#include <cstdint>
#include <string>

#include <cuda_runtime.h>

#include "../../include/smeagol_b200.h"

extern int smg_fail_from(int code, const std::string& msg);
extern void smg_count_launch();

namespace {

constexpr int kSymbols = 17;  // Z,A,C,G,T,U,N,W,S,M,K,R,Y,B,D,H,V (encode.py:11-29)
__constant__ char c_alphabet[kSymbols + 1] = "ZACGTUNWSMKRYBDHV";

__device__ __forceinline__ uint64_t splitmix64(uint64_t x)
{
    x += 0x9e3779b97f4a7c15ull;
    x = (x ^ (x >> 30)) * 0xbf58476d1ce4e5b9ull;
    x = (x ^ (x >> 27)) * 0x94d049bb133111ebull;
    return x ^ (x >> 31);
}
// uniform integer in [0, n) from one 64-bit draw (multiply-high; bias < n / 2^64)
__device__ __forceinline__ uint32_t bounded(const uint64_t r, const uint32_t n) { return (uint32_t)__umul64hi(r, (uint64_t)n); }

__global__ void __launch_bounds__(32) kmer_shuffle_kernel(const uint8_t* __restrict__ tok, const int64_t L,
                                                          const uint8_t* __restrict__ succ, const int64_t* __restrict__ off,
                                                          const int k, const uint64_t seed, uint8_t* __restrict__ scratch,
                                                          uint8_t* __restrict__ out, const int64_t stride, const int use_smem)
{
    extern __shared__ uint8_t s_work[];
    __shared__ int s_cnt[kSymbols];
    const int lane = threadIdx.x;
    const int64_t copy = blockIdx.x;
    uint8_t* work = use_smem ? s_work : scratch + copy * L;
    uint8_t* o = out + copy * stride;
    const uint64_t stream0 = splitmix64(seed ^ (0xd1342543de82ef95ull * (uint64_t)(copy + 1)));
    if (L <= 0) return;
    if (k == 1) {
        // uniform permutation of the bases: Fisher-Yates by one lane on the copy's private array
        for (int64_t i = lane; i < L; i += 32) work[i] = tok[i];
        __syncwarp();
        if (lane == 0) {
            for (int64_t i = L - 1; i > 0; --i) {
                const uint32_t j = bounded(splitmix64(stream0 + (uint64_t)i), (uint32_t)(i + 1));
                const uint8_t a = work[i], b = work[j];
                work[i] = b;
                work[j] = a;
            }
        }
        __syncwarp();
        for (int64_t i = lane; i < L; i += 32) o[i] = (uint8_t)c_alphabet[work[i]];
        return;
    }
    // ---- k == 2: successor lists (bucketed by symbol, original order) -> private copy
    for (int64_t i = lane; i < L - 1; i += 32) work[i] = succ[i];
    if (lane < kSymbols) s_cnt[lane] = 0;
    __syncwarp();
    if (lane < kSymbols) {  // one symbol per lane: permute all but the LAST successor of the symbol
        const int64_t a = off[lane], n = off[lane + 1] - off[lane];
        const uint64_t st = splitmix64(stream0 ^ (0x2545f4914f6cdd1dull * (uint64_t)(lane + 1)));
        for (int64_t i = n - 2; i > 0; --i) {
            const uint32_t j = bounded(splitmix64(st + (uint64_t)i), (uint32_t)(i + 1));
            const uint8_t x = work[a + i], y = work[a + j];
            work[a + i] = y;
            work[a + j] = x;
        }
    }
    __syncwarp();
    if (lane == 0) {  // re-walk from the first base, consuming every symbol's list in order
        int cur = tok[0];
        o[0] = (uint8_t)c_alphabet[cur];
        for (int64_t j = 1; j < L; ++j) {
            const int c = s_cnt[cur];
            s_cnt[cur] = c + 1;
            cur = work[off[cur] + c];
            o[j] = (uint8_t)c_alphabet[cur];
        }
    }
}

}  // namespace

extern "C" {

int smg_kmer_shuffle(const uint8_t* d_tok, int64_t len, const uint8_t* d_succ, const int64_t* d_off, int32_t n_copies, int32_t k,
                     uint64_t seed, uint8_t* d_scratch, uint8_t* d_out, int64_t stride, void* stream)
{
    if (!d_tok || len < 0 || n_copies < 0 || !d_out || stride < len) return smg_fail_from(SMG_ERR_INVALID, "smg_kmer_shuffle: bad argument");
    if (k != 1 && k != 2) return smg_fail_from(SMG_ERR_INVALID, "smg_kmer_shuffle: k must be 1 or 2");
    if (k == 2 && (!d_succ || !d_off)) return smg_fail_from(SMG_ERR_INVALID, "smg_kmer_shuffle: successor lists missing");
    if (n_copies == 0 || len == 0) return SMG_OK;
    const size_t need = (size_t)len;
    int use_smem = need <= 200 * 1024 ? 1 : 0;
    if (!use_smem && !d_scratch) return smg_fail_from(SMG_ERR_INVALID, "smg_kmer_shuffle: scratch of n_copies * len bytes needed");
    cudaError_t e;
    if (use_smem && need > 48 * 1024) {
        e = cudaFuncSetAttribute(kmer_shuffle_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)need);
        if (e != cudaSuccess) return smg_fail_from(SMG_ERR_CUDA, cudaGetErrorString(e));
    }
    kmer_shuffle_kernel<<<(unsigned)n_copies, 32, use_smem ? need : 0, (cudaStream_t)stream>>>(d_tok, len, d_succ, d_off, k, seed,
                                                                                               d_scratch, d_out, stride, use_smem);
    e = cudaGetLastError();
    if (e != cudaSuccess) return smg_fail_from(SMG_ERR_CUDA, std::string("kmer_shuffle_kernel: ") + cudaGetErrorString(e));
    smg_count_launch();
    return SMG_OK;
}

}  // extern "C"
