// tcgen05 / TMEM probe for the PWM-scan prefilter.  Four parts, all on one B200:
//   A. known-answer test of the shared-memory descriptors used by scan_tc5.cuh: weights as the K-major A operand,
//      the one-hot sequence as a TOEPLITZ K-major B operand (row p+2 = row p shifted by 16 bytes, so the "im2col" matrix
//      is never materialised: LBO = 16 B overlaps the core matrices), checked against a host contraction;
//   B. TMEM -> register read bandwidth (tcgen05.ld 32x32b.xN from 4 / 8 warps, 1 / 2 CTAs per SM), bytes/clk/SM;
//   C. a realistic epilogue loop (tcgen05.ld x64 + 3-input max tree + threshold compare + rare queue push), cells/clk/SM;
//   D. tcgen05.mma issue rate for M=128, N=256, K=16*KS from shared memory, alone and with C running beside it.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo -I smeagol_b200/csrc tools/tcgen05_probe.cu
#include <cuda_fp16.h>
#include <cuda_runtime.h>

#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <vector>

#include "tc5.cuh"

#define CK(x)                                                                                  \
    do {                                                                                       \
        cudaError_t e_ = (x);                                                                  \
        if (e_ != cudaSuccess) {                                                               \
            printf("CUDA error %s at line %d\n", cudaGetErrorString(e_), __LINE__);            \
            exit(1);                                                                           \
        }                                                                                      \
    } while (0)

using namespace tc5;

// ------------------------------------------------------------------------------------------------ part A
// out[m][e * 256 + j] = sum_{k < KW, b < 4} w[m][4k + b] * onehot(seq[2j + e + k])[b],  m < 128, j < 256, e < 2
struct KatArgs {
    const uint8_t* seq;   // base codes 0..3, >= 512 + 64 entries
    const __half* w;      // [128][K] (K = 16 * ks)
    int ks;
    uint32_t a_lbo, a_sbo, b_lbo, b_sbo;
    int b_mode;           // 0: Toeplitz one-hot strip, 1: materialised im2col [256][K]
    float* out;           // [128][512]
};

__global__ void __launch_bounds__(128) kat_kernel(const KatArgs g)
{
    extern __shared__ __align__(1024) uint8_t smem[];
    __shared__ __align__(8) uint64_t bar;
    __shared__ uint32_t tmem_base;
    const int K = 16 * g.ks;
    uint8_t* sA = smem;                             // (K / 8) panels of 128 rows x 16 B
    uint8_t* sB0 = smem + 16384;                    // even window starts
    uint8_t* sB1 = sB0 + 40960;                     // odd window starts
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    for (int i = tid; i < 128 * K; i += 128) {
        const int m = i / K, k = i % K;
        *reinterpret_cast<__half*>(sA + (k / 8) * 2048 + m * 16 + (k % 8) * 2) = g.w[i];
    }
    if (g.b_mode == 0) {
        // strip e: base q of the strip is sequence base q + e; 4 halves (8 B) per base
        for (int q = tid; q < 512 + 64; q += 128)
            for (int e = 0; e < 2; ++e) {
                const int c = g.seq[q + e];
                uint8_t* dst = (e ? sB1 : sB0) + q * 8;
                for (int b = 0; b < 4; ++b) *reinterpret_cast<__half*>(dst + 2 * b) = __float2half(b == c ? 1.f : 0.f);
            }
    } else {
        for (int i = tid; i < 2 * 256 * K; i += 128) {
            const int e = i / (256 * K), n = (i / K) % 256, k = i % K;
            const int c = g.seq[2 * n + e + k / 4];
            *reinterpret_cast<__half*>((e ? sB1 : sB0) + (k / 8) * 4096 + n * 16 + (k % 8) * 2) = __float2half((k % 4) == c ? 1.f : 0.f);
        }
    }
    if (tid == 0) {
        mbar_init(&bar, 1);
        mbar_fence_init();
    }
    fence_proxy_async();
    if (warp == 0) tmem_alloc(&tmem_base, 512);
    fence_before_sync();
    __syncthreads();
    fence_after_sync();
    const uint32_t tb = tmem_base;
    if (tid == 0) {
        const uint32_t idesc = make_idesc_f16_f32(128, 256);
        for (int e = 0; e < 2; ++e)
            for (int s = 0; s < g.ks; ++s) {
                const uint64_t da = make_smem_desc(smem_u32(sA) + s * 4096, g.a_lbo, g.a_sbo);
                const uint32_t boff = g.b_mode == 0 ? s * 32 : s * 8192;
                const uint64_t db = make_smem_desc(smem_u32(e ? sB1 : sB0) + boff, g.b_lbo, g.b_sbo);
                mma_f16_ss(tb + e * 256, da, db, idesc, s > 0);
            }
        mma_commit(&bar);
    }
    mbar_wait(&bar, 0);
    fence_after_sync();
    for (int c = 0; c < 8; ++c) {
        uint32_t r[64];
        tmem_ld_32x32b_x64(r, tb + ((uint32_t)(warp * 32) << 16) + c * 64);
        tmem_ld_wait();
        for (int i = 0; i < 64; ++i) g.out[(size_t)(warp * 32 + lane) * 512 + c * 64 + i] = __uint_as_float(r[i]);
    }
    fence_before_sync();
    __syncthreads();
    if (warp == 0) tmem_dealloc(tb, 512);
}

// ------------------------------------------------------------------------------------------------ part B
template <int NREG>
__device__ __forceinline__ uint32_t ld_n(const uint32_t taddr)  // ptxas drops a tcgen05.ld whose results are all dead
{
    uint32_t r[NREG];
    if constexpr (NREG == 16) tmem_ld_32x32b_x16(r, taddr);
    if constexpr (NREG == 32) tmem_ld_32x32b_x32(r, taddr);
    if constexpr (NREG == 64) tmem_ld_32x32b_x64(r, taddr);
    if constexpr (NREG == 128) tmem_ld_32x32b_x128(r, taddr);
    return r[NREG - 1];
}

template <int NREG>
__global__ void ldtm_bw_kernel(const int iters, const int ncols, unsigned long long* cycles, uint32_t* sink)
{
    __shared__ uint32_t tmem_base;
    const int warp = threadIdx.x >> 5;
    if (warp == 0) tmem_alloc(&tmem_base, ncols);
    fence_before_sync();
    __syncthreads();
    fence_after_sync();
    const uint32_t tb = tmem_base + ((uint32_t)((warp & 3) * 32) << 16);
    uint32_t acc = 0;
    const long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
        uint32_t v[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) v[u] = ld_n<NREG>(tb + ((it * 4 + u) * NREG) % ncols);
        tmem_ld_wait();
        acc ^= (v[0] ^ v[1]) ^ (v[2] ^ v[3]);
    }
    const long long t1 = clock64();
    if (acc == 0x9e3779b9u) sink[threadIdx.x] = acc;
    if ((threadIdx.x & 31) == 0) atomicMax(cycles, (unsigned long long)(t1 - t0));
    fence_before_sync();
    __syncthreads();
    if (warp == 0) tmem_dealloc(tmem_base, ncols);
}

// ------------------------------------------------------------------------------------------------ part C / D
__device__ __forceinline__ float max3(const float a, const float b, const float c)
{
    float d;
    asm("max.f32 %0, %1, %2, %3;" : "=f"(d) : "f"(a), "f"(b), "f"(c));
    return d;
}

// one epilogue step over 64 accumulator cells of this thread's TMEM lane: returns the number of candidates pushed
__device__ __forceinline__ int epilogue_x64(const uint32_t taddr, const float thr, uint32_t* q, int qn, const uint32_t tag)
{
    uint32_t r[64];
    tmem_ld_32x32b_x64(r, taddr);
    tmem_ld_wait();
    float gm[8];
#pragma unroll
    for (int gI = 0; gI < 8; ++gI) {
        const float* v = reinterpret_cast<const float*>(r) + gI * 8;
        gm[gI] = max3(max3(v[0], v[1], v[2]), max3(v[3], v[4], v[5]), fmaxf(v[6], v[7]));
    }
    const float m = max3(max3(gm[0], gm[1], gm[2]), max3(gm[3], gm[4], gm[5]), fmaxf(gm[6], gm[7]));
    if (__any_sync(0xffffffffu, m > thr)) {
#pragma unroll
        for (int gI = 0; gI < 8; ++gI) {
            if (__any_sync(0xffffffffu, gm[gI] > thr)) {
#pragma unroll
                for (int i = 0; i < 8; ++i)
                    if (__uint_as_float(r[gI * 8 + i]) > thr) {
                        q[(qn & 63) * 32] = tag + gI * 8 + i;
                        ++qn;
                    }
            }
        }
    }
    return qn;
}

// warps 0..3: epilogue loop over the 512 TMEM columns; warp 4 (if mma_groups > 0): issues groups of ks MMAs (M=128, N=256)
__global__ void __launch_bounds__(160) epi_mma_kernel(const int epi_iters, const int mma_groups, const int ks, const float thr,
                                                      unsigned long long* epi_cycles, unsigned long long* mma_cycles,
                                                      unsigned long long* n_pushed)
{
    extern __shared__ __align__(1024) uint8_t smem[];
    __shared__ __align__(8) uint64_t bar[2];
    __shared__ uint32_t tmem_base;
    __shared__ uint32_t s_q[64 * 32 * 4];
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    // operands: ZERO weights and always-accumulate, so the MMAs leave the accumulator values (and with them the candidate
    // density of the epilogue beside them) unchanged; a valid one-hot strip.  Operand values do not change the timing.
    for (int i = tid; i < (16384 + 8192) / 4; i += blockDim.x)
        reinterpret_cast<uint32_t*>(smem)[i] = (i < 16384 / 4 || (i & 1)) ? 0u : 0x3c00u;
    if (tid == 0) {
        mbar_init(&bar[0], 1);
        mbar_init(&bar[1], 1);
        mbar_fence_init();
    }
    fence_proxy_async();
    if (warp == 0) tmem_alloc(&tmem_base, 512);
    fence_before_sync();
    __syncthreads();
    fence_after_sync();
    const uint32_t tb = tmem_base;
    if (warp < 4) {
        // fill this warp's 32 lanes x 512 columns with uniform [0,1) values (hash of lane/column)
        for (int c = 0; c < 512; c += 8) {
            uint32_t v[8];
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                uint32_t h = (uint32_t)(blockIdx.x * 131 + (warp * 32 + lane) * 7919 + (c + i)) * 2654435761u;
                h ^= h >> 15; h *= 2246822519u; h ^= h >> 13;
                v[i] = __float_as_uint((float)(h >> 8) * (1.0f / 16777216.0f));
            }
            tmem_st_32x32b_x8(tb + ((uint32_t)(warp * 32) << 16) + c, v);
        }
        tmem_st_wait();
    }
    fence_before_sync();
    __syncthreads();
    fence_after_sync();
    if (warp < 4) {
        int qn = 0;
        const long long t0 = clock64();
        for (int it = 0; it < epi_iters; ++it)
            qn = epilogue_x64(tb + ((uint32_t)(warp * 32) << 16) + (it * 64) % 512, thr, s_q + warp * 2048 + lane, qn, (uint32_t)it << 6);
        const long long t1 = clock64();
        if (lane == 0) atomicMax(epi_cycles, (unsigned long long)(t1 - t0));
        atomicAdd(n_pushed, (unsigned long long)qn);
    } else if (mma_groups > 0 && lane == 0) {
        const uint32_t idesc = make_idesc_f16_f32(128, 256);
        const long long t0 = clock64();
        for (int gI = 0; gI < mma_groups; ++gI) {
            const int buf = gI & 1;
            if (gI >= 2) mbar_wait(&bar[buf], ((gI >> 1) - 1) & 1);  // the group issued two rounds ago on this buffer is done
            for (int s = 0; s < ks; ++s) {
                const uint64_t da = make_smem_desc(smem_u32(smem) + s * 4096, 2048, 128);
                const uint64_t db = make_smem_desc(smem_u32(smem) + 16384 + s * 32, 16, 128);
                mma_f16_ss(tb + buf * 256, da, db, idesc, 1u);
            }
            mma_commit(&bar[buf]);
        }
        for (int buf = 0; buf < 2; ++buf) {
            const int n_on_buf = (mma_groups + 1 - buf) / 2;
            if (n_on_buf > 0) mbar_wait(&bar[buf], (n_on_buf - 1) & 1);
        }
        const long long t1 = clock64();
        atomicMax(mma_cycles, (unsigned long long)(t1 - t0));
    }
    fence_before_sync();
    __syncthreads();
    if (warp == 0) tmem_dealloc(tb, 512);
}

// ------------------------------------------------------------------------------------------------ host
static double run_kat(const int ks, const uint32_t a_lbo, const uint32_t a_sbo, const uint32_t b_lbo, const uint32_t b_sbo,
                      const int b_mode, const char* label)
{
    const int K = 16 * ks, KW = 4 * ks;
    std::vector<uint8_t> seq(1024);
    std::vector<__half> w(128 * K);
    std::vector<float> wf(128 * K);
    uint32_t s = 12345u;
    auto rnd = [&]() { s = s * 1664525u + 1013904223u; return s >> 8; };
    for (auto& c : seq) c = rnd() & 3;
    for (int i = 0; i < 128 * K; ++i) {
        wf[i] = (float)((int)(rnd() % 513) - 384) / 64.f;  // multiples of 1/64 in [-6, 2]: exact in fp16, sums exact in fp32
        w[i] = __float2half(wf[i]);
    }
    uint8_t* d_seq; __half* d_w; float* d_out;
    CK(cudaMalloc(&d_seq, seq.size()));
    CK(cudaMalloc(&d_w, w.size() * 2));
    CK(cudaMalloc(&d_out, 128 * 512 * 4));
    CK(cudaMemcpy(d_seq, seq.data(), seq.size(), cudaMemcpyHostToDevice));
    CK(cudaMemcpy(d_w, w.data(), w.size() * 2, cudaMemcpyHostToDevice));
    CK(cudaMemset(d_out, 0xff, 128 * 512 * 4));
    KatArgs g{d_seq, d_w, ks, a_lbo, a_sbo, b_lbo, b_sbo, b_mode, d_out};
    const int smem_bytes = 16384 + 2 * 40960;
    CK(cudaFuncSetAttribute(kat_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes));
    kat_kernel<<<1, 128, smem_bytes>>>(g);
    CK(cudaGetLastError());
    CK(cudaDeviceSynchronize());
    std::vector<float> out(128 * 512);
    CK(cudaMemcpy(out.data(), d_out, out.size() * 4, cudaMemcpyDeviceToHost));
    double maxerr = 0;
    long bad = 0;
    for (int m = 0; m < 128; ++m)
        for (int e = 0; e < 2; ++e)
            for (int j = 0; j < 256; ++j) {
                float ref = 0;
                for (int k = 0; k < KW; ++k) ref += wf[m * K + 4 * k + seq[2 * j + e + k]];
                const double err = fabs((double)out[m * 512 + e * 256 + j] - ref);
                if (!(err <= 1e-6)) ++bad;
                if (err == err && err > maxerr) maxerr = err;
            }
    printf("KAT %-34s ks=%d A(lbo=%u,sbo=%u) B(lbo=%u,sbo=%u): mismatches %ld / 65536, max |err| %.3g -> %s\n", label, ks, a_lbo,
           a_sbo, b_lbo, b_sbo, bad, maxerr, bad == 0 ? "PASS" : "FAIL");
    cudaFree(d_seq); cudaFree(d_w); cudaFree(d_out);
    return (double)bad;
}

template <int NREG>
static void run_bw(const int warps, const int ctas_per_sm, const int n_sm, const double ghz_hint)
{
    unsigned long long* d_cyc;
    CK(cudaMalloc(&d_cyc, 8 + 4096));
    CK(cudaMemset(d_cyc, 0, 8));
    const int iters = 20000, ncols = 512 / ctas_per_sm;
    cudaEvent_t e0, e1;
    CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    ldtm_bw_kernel<NREG><<<n_sm * ctas_per_sm, warps * 32>>>(100, ncols, d_cyc, (uint32_t*)(d_cyc + 1));  // warm-up
    CK(cudaDeviceSynchronize());
    CK(cudaMemset(d_cyc, 0, 8));
    CK(cudaEventRecord(e0));
    ldtm_bw_kernel<NREG><<<n_sm * ctas_per_sm, warps * 32>>>(iters, ncols, d_cyc, (uint32_t*)(d_cyc + 1));
    CK(cudaEventRecord(e1));
    CK(cudaGetLastError());
    CK(cudaDeviceSynchronize());
    float ms;
    CK(cudaEventElapsedTime(&ms, e0, e1));
    unsigned long long cyc;
    CK(cudaMemcpy(&cyc, d_cyc, 8, cudaMemcpyDeviceToHost));
    const double bytes_per_sm = (double)iters * 4 * NREG * 32 * 4 * warps * ctas_per_sm;
    printf("LDTM 32x32b.x%-3d warps/CTA=%d CTAs/SM=%d: %8.1f B/clk/SM (clock64), %8.1f cells/clk/SM, %.3f ms -> %.1f GB/s/SM (%.2f GHz eff)\n",
           NREG, warps, ctas_per_sm, bytes_per_sm / (double)cyc, bytes_per_sm / 4 / (double)cyc, ms, bytes_per_sm / ms * 1e-6,
           (double)cyc / ms * 1e-6);
    (void)ghz_hint;
    cudaFree(d_cyc);
}

static void run_epi(const int epi_iters, const int mma_groups, const int ks, const float thr, const int n_sm)
{
    unsigned long long* d;
    CK(cudaMalloc(&d, 24));
    const int smem_bytes = 16384 + 8192;
    CK(cudaFuncSetAttribute(epi_mma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes));
    epi_mma_kernel<<<n_sm, 160, smem_bytes>>>(10, mma_groups ? 4 : 0, ks, thr, d, d + 1, d + 2);
    CK(cudaDeviceSynchronize());
    CK(cudaMemset(d, 0, 24));
    cudaEvent_t e0, e1;
    CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    CK(cudaEventRecord(e0));
    epi_mma_kernel<<<n_sm, 160, smem_bytes>>>(epi_iters, mma_groups, ks, thr, d, d + 1, d + 2);
    CK(cudaEventRecord(e1));
    CK(cudaGetLastError());
    CK(cudaDeviceSynchronize());
    float ms;
    CK(cudaEventElapsedTime(&ms, e0, e1));
    unsigned long long h[3];
    CK(cudaMemcpy(h, d, 24, cudaMemcpyDeviceToHost));
    const double cells = (double)epi_iters * 64 * 128;
    printf("EPI+MMA epi_iters=%d mma_groups=%d ks=%d thr=%.5f: ", epi_iters, mma_groups, ks, thr);
    if (epi_iters) printf("epilogue %.1f cells/clk/SM (%.4f cand/cell) ", cells / (double)h[0], (double)h[2] / (cells * n_sm));
    if (mma_groups) printf("| MMA %.1f clk/group of %d (M128 N256 K16) = %.1f cells/clk/SM, %.0f MAC/clk/SM ", (double)h[1] / mma_groups, ks,
                           128.0 * 256 * mma_groups / (double)h[1], 128.0 * 256 * 16 * ks * mma_groups / (double)h[1]);
    printf("| %.3f ms\n", ms);
    cudaFree(d);
}

int main()
{
    cudaDeviceProp prop;
    CK(cudaGetDeviceProperties(&prop, 0));
    printf("device %s, %d SMs, sm_%d%d\n", prop.name, prop.multiProcessorCount, prop.major, prop.minor);
    const int n_sm = prop.multiProcessorCount;
    // ---- A: descriptor conventions
    run_kat(3, 2048, 128, 4096, 128, 1, "materialised B, LBO=K-half stride");
    run_kat(3, 2048, 128, 16, 128, 0, "TOEPLITZ B (LBO=16 overlapping)");
    run_kat(2, 2048, 128, 16, 128, 0, "TOEPLITZ B ks=2");
    run_kat(4, 2048, 128, 16, 128, 0, "TOEPLITZ B ks=4");
    // ---- B: TMEM read bandwidth
    run_bw<16>(4, 1, n_sm, 1.9);
    run_bw<32>(4, 1, n_sm, 1.9);
    run_bw<64>(4, 1, n_sm, 1.9);
    run_bw<128>(4, 1, n_sm, 1.9);
    run_bw<32>(8, 1, n_sm, 1.9);
    run_bw<64>(8, 1, n_sm, 1.9);
    run_bw<32>(4, 2, n_sm, 1.9);
    run_bw<64>(4, 2, n_sm, 1.9);
    run_bw<64>(1, 1, n_sm, 1.9);
    // ---- C: epilogue alone (no candidates / c2-like density 1.5e-3 / 6e-3)
    run_epi(40000, 0, 3, 2.0f, n_sm);
    run_epi(40000, 0, 3, 1.0f - 1.5e-3f, n_sm);
    run_epi(40000, 0, 3, 1.0f - 6e-3f, n_sm);
    // ---- D: MMA alone, then both
    run_epi(0, 20000, 2, 2.0f, n_sm);
    run_epi(0, 20000, 3, 2.0f, n_sm);
    run_epi(0, 20000, 6, 2.0f, n_sm);
    run_epi(40000, 20000, 3, 1.0f - 1.5e-3f, n_sm);
    run_epi(40000, 30000, 2, 1.0f - 1.5e-3f, n_sm);
    return 0;
}
