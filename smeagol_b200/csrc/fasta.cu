// FASTA text -> per-record ASCII sequences on the device (the step before encode).
//
// Replaces the reference's read_fasta (smeagol/io.py:16-41: Biopython SeqIO.parse(handle, "fasta")) for the part that
// touches every byte: locating the header lines and removing line breaks / blanks from the sequence lines.  The
// header TEXT (id, name, description) is parsed by the host from its own copy of the file at the offsets found here.
// HBM-bound byte work: 16 bytes per thread (128-bit loads), one pass to find the headers, one to count the kept bytes
// per record, one stream compaction (cub::DeviceSelect, library) -- about 4 bytes of traffic per input byte.
//
// Semantics (Biopython's SimpleFastaParser): a record starts at a '>' that is the first byte of a line; everything
// before the first header is ignored; inside a record '\n', '\r' and ' ' are dropped ('\t' too), every other byte is
// sequence (validated later by smg_pack, which reports the reference's KeyError for bytes outside its alphabet).
#include <cstdint>
#include <string>

#include <cub/device/device_select.cuh>
#include <thrust/iterator/transform_iterator.h>
#include <cuda_runtime.h>

#include "../../include/smeagol_b200.h"

extern int smg_fail_from(int code, const std::string& msg);
extern void smg_count_launch();

namespace {

__device__ __forceinline__ bool fasta_keep(const uint8_t c) { return c != '\n' && c != '\r' && c != ' ' && c != '\t'; }

struct KeepOp {
    __host__ __device__ __forceinline__ bool operator()(const uint8_t c) const
    {
        return c != '\n' && c != '\r' && c != ' ' && c != '\t';
    }
};
struct UpperOp {
    __host__ __device__ __forceinline__ uint8_t operator()(const uint8_t c) const
    {
        return (c >= 'a' && c <= 'z') ? (uint8_t)(c - 32) : c;
    }
};

__device__ __forceinline__ void load16(const uint8_t* __restrict__ b, const int64_t i0, const int64_t n, uint8_t (&raw)[16])
{
    if (i0 + 16 <= n) {
        const uint4 v = *reinterpret_cast<const uint4*>(b + i0);  // the buffer is 16-byte aligned (checked by the caller)
        memcpy(raw, &v, 16);
    } else {
#pragma unroll
        for (int j = 0; j < 16; ++j) raw[j] = (i0 + j < n) ? b[i0 + j] : (uint8_t)'\n';
    }
}

// header starts: '>' that is the first byte of a line.  Unordered append (headers are rare), warp-aggregated.
__global__ void __launch_bounds__(256) fasta_find_headers_kernel(const uint8_t* __restrict__ b, const int64_t n,
                                                                 int64_t* __restrict__ hdr, const unsigned long long cap,
                                                                 unsigned long long* __restrict__ n_hdr)
{
    const int64_t stride = (int64_t)gridDim.x * blockDim.x * 16;
    for (int64_t i0 = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) * 16; i0 < n; i0 += stride) {
        uint8_t raw[16];
        load16(b, i0, n, raw);
        uint8_t prev = (i0 == 0) ? (uint8_t)'\n' : b[i0 - 1];
#pragma unroll
        for (int j = 0; j < 16; ++j) {
            if (raw[j] == '>' && prev == '\n' && i0 + j < n) {
                const unsigned long long k = atomicAdd(n_hdr, 1ull);
                if (k < cap) hdr[k] = i0 + j;
            }
            prev = raw[j];
        }
    }
}

// one warp per header: find the end of the header line, then overwrite the line with '\n' so that the byte-wise
// keep predicate drops it.  hdr_end[r] = offset of the terminating '\n' (or n).
__global__ void __launch_bounds__(128) fasta_blank_headers_kernel(uint8_t* __restrict__ b, const int64_t n,
                                                                  const int64_t* __restrict__ hdr, const int64_t n_hdr,
                                                                  int64_t* __restrict__ hdr_end)
{
    const int lane = threadIdx.x & 31;
    const int64_t w = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (w >= n_hdr) return;
    const int64_t start = hdr[w];
    int64_t p = start, end = n;
    while (p < n) {
        const uint8_t c = (p + lane < n) ? b[p + lane] : (uint8_t)'\n';
        const unsigned m = __ballot_sync(0xffffffffu, c == '\n');
        if (m) {
            end = min(p + (int64_t)(__ffs((int)m) - 1), n);
            break;
        }
        p += 32;
    }
    __syncwarp();
    for (int64_t q = start + lane; q < end; q += 32) b[q] = '\n';
    if (lane == 0) hdr_end[w] = end;
    // everything before the first header belongs to no record (SimpleFastaParser skips it): blank it as well
    if (w == 0)
        for (int64_t q = lane; q < start; q += 32) b[q] = '\n';
}

// kept bytes per record (after the headers were blanked); record of a byte = last header at or before it
__global__ void __launch_bounds__(256) fasta_count_kernel(const uint8_t* __restrict__ b, const int64_t n,
                                                          const int64_t* __restrict__ hdr, const int64_t n_hdr,
                                                          unsigned long long* __restrict__ seq_len)
{
    __shared__ unsigned long long s_sum;
    __shared__ long long s_rec;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x * 16;
    // the trip count is the same for every thread of a block (block-wide barriers below)
    const int64_t n_round = (n + (int64_t)blockDim.x * 16 - 1) / ((int64_t)blockDim.x * 16) * ((int64_t)blockDim.x * 16);
    for (int64_t base = (int64_t)blockIdx.x * blockDim.x * 16; base < n_round; base += stride) {
        const int64_t i0 = base + (int64_t)threadIdx.x * 16;
        long long rec = -1;
        unsigned cnt = 0;
        bool straddles = false;
        if (i0 < n) {
            uint8_t raw[16];
            load16(b, i0, n, raw);
            int64_t lo = -1, hi = n_hdr - 1;  // largest r with hdr[r] <= i0 (-1: before the first header)
            while (lo < hi) {
                const int64_t mid = (lo + hi + 1) >> 1;
                if (hdr[mid] <= i0) lo = mid; else hi = mid - 1;
            }
            rec = lo;
            straddles = (rec + 1 < n_hdr) && (hdr[rec + 1] < i0 + 16);
            if (!straddles) {
#pragma unroll
                for (int j = 0; j < 16; ++j) cnt += (i0 + j < n && fasta_keep(raw[j])) ? 1u : 0u;
                if (rec < 0) cnt = 0;
            } else {  // a header starts inside these 16 bytes: attribute byte by byte
                long long r = rec;
                for (int j = 0; j < 16 && i0 + j < n; ++j) {
                    while (r + 1 < n_hdr && hdr[r + 1] <= i0 + j) ++r;
                    if (r >= 0 && fasta_keep(raw[j])) atomicAdd(&seq_len[r], 1ull);
                }
                cnt = 0;
            }
        }
        // common case: the whole block lies inside one record -> one atomic per block
        if (threadIdx.x == 0) { s_sum = 0ull; s_rec = rec; }
        __syncthreads();
        const bool same = __syncthreads_and((i0 >= n) || (rec == s_rec && !straddles) || cnt == 0);
        if (same) {
            unsigned w = __reduce_add_sync(0xffffffffu, cnt);
            if ((threadIdx.x & 31) == 0 && w) atomicAdd(&s_sum, (unsigned long long)w);
            __syncthreads();
            if (threadIdx.x == 0 && s_sum && s_rec >= 0) atomicAdd(&seq_len[s_rec], s_sum);
        } else if (cnt && rec >= 0) {
            atomicAdd(&seq_len[rec], (unsigned long long)cnt);
        }
        __syncthreads();
    }
}

}  // namespace

extern "C" {

int smg_fasta_index(const uint8_t* d_bytes, int64_t n, int64_t* d_hdr_start, uint64_t hdr_cap, uint64_t* d_n_hdr, void* stream)
{
    if (!d_bytes || n < 0 || !d_hdr_start || !d_n_hdr) return smg_fail_from(SMG_ERR_INVALID, "smg_fasta_index: bad argument");
    if ((reinterpret_cast<uintptr_t>(d_bytes) & 15) != 0) return smg_fail_from(SMG_ERR_INVALID, "smg_fasta_index: buffer must be 16-byte aligned");
    cudaStream_t st = (cudaStream_t)stream;
    cudaError_t e = cudaMemsetAsync(d_n_hdr, 0, sizeof(uint64_t), st);
    if (e != cudaSuccess) return smg_fail_from(SMG_ERR_CUDA, cudaGetErrorString(e));
    if (n == 0) return SMG_OK;
    const int64_t threads = (n + 15) / 16;
    const unsigned blocks = (unsigned)std::min<int64_t>((threads + 255) / 256, 148 * 32);
    fasta_find_headers_kernel<<<blocks, 256, 0, st>>>(d_bytes, n, d_hdr_start, hdr_cap, reinterpret_cast<unsigned long long*>(d_n_hdr));
    e = cudaGetLastError();
    if (e != cudaSuccess) return smg_fail_from(SMG_ERR_CUDA, std::string("fasta_find_headers_kernel: ") + cudaGetErrorString(e));
    smg_count_launch();
    return SMG_OK;
}

int smg_fasta_compact(uint8_t* d_bytes, int64_t n, const int64_t* d_hdr_start, int64_t n_hdr, int64_t* d_hdr_end,
                      uint64_t* d_seq_len, uint8_t* d_out, int64_t* d_n_out, int upper, void* d_temp, uint64_t* temp_bytes,
                      void* stream)
{
    if (!temp_bytes || n < 0 || n_hdr < 0) return smg_fail_from(SMG_ERR_INVALID, "smg_fasta_compact: bad argument");
    if (n >= (1ll << 31)) return smg_fail_from(SMG_ERR_UNSUPPORTED, "smg_fasta_compact: at most 2^31 - 1 bytes per call");
    cudaStream_t st = (cudaStream_t)stream;
    size_t tb = d_temp ? (size_t)*temp_bytes : 0;
    thrust::transform_iterator<UpperOp, const uint8_t*> up_it((const uint8_t*)d_bytes, UpperOp());
    cudaError_t e;
    if (!d_temp) {  // size query
        e = upper ? cub::DeviceSelect::If(nullptr, tb, up_it, d_out, d_n_out, (int)n, KeepOp(), st)
                  : cub::DeviceSelect::If(nullptr, tb, (const uint8_t*)d_bytes, d_out, d_n_out, (int)n, KeepOp(), st);
        if (e != cudaSuccess) return smg_fail_from(SMG_ERR_CUDA, std::string("cub::DeviceSelect: ") + cudaGetErrorString(e));
        *temp_bytes = (uint64_t)tb;
        return SMG_OK;
    }
    if (!d_bytes || !d_hdr_start || !d_hdr_end || !d_seq_len || !d_out || !d_n_out)
        return smg_fail_from(SMG_ERR_INVALID, "smg_fasta_compact: bad argument");
    if ((reinterpret_cast<uintptr_t>(d_bytes) & 15) != 0) return smg_fail_from(SMG_ERR_INVALID, "smg_fasta_compact: buffer must be 16-byte aligned");
    e = cudaMemsetAsync(d_seq_len, 0, sizeof(uint64_t) * (size_t)std::max<int64_t>(n_hdr, 1), st);
    if (e != cudaSuccess) return smg_fail_from(SMG_ERR_CUDA, cudaGetErrorString(e));
    if (n_hdr > 0 && n > 0) {
        const unsigned hb = (unsigned)((n_hdr * 32 + 127) / 128);
        fasta_blank_headers_kernel<<<hb, 128, 0, st>>>(d_bytes, n, d_hdr_start, n_hdr, d_hdr_end);
        smg_count_launch();
        const int64_t threads = (n + 15) / 16;
        const unsigned blocks = (unsigned)std::min<int64_t>((threads + 255) / 256, 148 * 32);
        fasta_count_kernel<<<blocks, 256, 0, st>>>(d_bytes, n, d_hdr_start, n_hdr, reinterpret_cast<unsigned long long*>(d_seq_len));
        smg_count_launch();
        e = cudaGetLastError();
        if (e != cudaSuccess) return smg_fail_from(SMG_ERR_CUDA, std::string("fasta kernels: ") + cudaGetErrorString(e));
    }
    if (n_hdr == 0 || n == 0) {  // no header line: no record (whatever the bytes are)
        e = cudaMemsetAsync(d_n_out, 0, sizeof(int64_t), st);
        if (e != cudaSuccess) return smg_fail_from(SMG_ERR_CUDA, cudaGetErrorString(e));
        return SMG_OK;
    }
    e = upper ? cub::DeviceSelect::If(d_temp, tb, up_it, d_out, d_n_out, (int)n, KeepOp(), st)
              : cub::DeviceSelect::If(d_temp, tb, (const uint8_t*)d_bytes, d_out, d_n_out, (int)n, KeepOp(), st);
    if (e != cudaSuccess) return smg_fail_from(SMG_ERR_CUDA, std::string("cub::DeviceSelect: ") + cudaGetErrorString(e));
    smg_count_launch();
    return SMG_OK;
}

}  // extern "C"
