#include <cstdio>
#include <cstdint>
#include <cstdlib>
#include <vector>
#include <cuda_runtime.h>
#define CK(x) do{cudaError_t e=(x); if(e!=cudaSuccess){printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); exit(1);} }while(0)

__global__ __launch_bounds__(256) void k_fadd(float* out, int iters, float c) {
    float a[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) a[i] = threadIdx.x * 0.001f + i;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < 16; ++i) a[i] = __fadd_rn(a[i], c);
    }
    float s = 0;
#pragma unroll
    for (int i = 0; i < 16; ++i) s += a[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
__global__ __launch_bounds__(256) void k_fadd2(float* out, int iters, float c) {
    float2 a[16];
    float2 cc = make_float2(c, c * 0.5f);
#pragma unroll
    for (int i = 0; i < 16; ++i) a[i] = make_float2(threadIdx.x * 0.001f + i, i);
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < 16; ++i) a[i] = __fadd2_rn(a[i], cc);
    }
    float s = 0;
#pragma unroll
    for (int i = 0; i < 16; ++i) s += a[i].x + a[i].y;
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
// mix: FADD2 + IADD-type overhead to see co-issue
__global__ __launch_bounds__(256) void k_mix(float* out, int iters, float c) {
    float2 a[12];
    float2 cc = make_float2(c, c * 0.5f);
    unsigned x = threadIdx.x, y = 1;
#pragma unroll
    for (int i = 0; i < 12; ++i) a[i] = make_float2(threadIdx.x * 0.001f + i, i);
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < 12; ++i) a[i] = __fadd2_rn(a[i], cc);
#pragma unroll
        for (int i = 0; i < 8; ++i) { x = (x >> 1) ^ (y + i); y = (y << 1) | (x & 3); }
    }
    float s = x + y;
#pragma unroll
    for (int i = 0; i < 12; ++i) s += a[i].x + a[i].y;
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int W, int NT>
__global__ __launch_bounds__(NT) void scan_proto(const uint32_t* __restrict__ packed,
    const float2* __restrict__ tab, const float* __restrict__ thr, int nwords, unsigned long long* out)
{
    const int lane = threadIdx.x & 31;
    const int warp = (NT == 32) ? blockIdx.x : ((blockIdx.x * NT + threadIdx.x) >> 5);
    float2 t[W][4];
#pragma unroll
    for (int k = 0; k < W; ++k)
#pragma unroll
        for (int b = 0; b < 4; ++b) t[k][b] = tab[(k * 4 + b) * 32 + lane];
    float2 r[W];
#pragma unroll
    for (int k = 0; k < W; ++k) r[k] = make_float2(0.f, 0.f);
    const float th = thr[lane];
    unsigned long long acc = 0;
    const uint32_t* wp = packed + (size_t)(warp & 4095) * nwords;
    uint32_t w = wp[0];
    for (int i = 0; i < nwords; ++i) {
        uint32_t wn = wp[i + 1];
#pragma unroll
        for (int j = 0; j < 16; ++j) {
            const uint32_t b = (w >> (2 * j)) & 3u;
            float2 o;
#define STEP(B)                                                     \
    {                                                               \
        o = __fadd2_rn(r[W - 2], t[W - 1][B]);                      \
        _Pragma("unroll") for (int k = W - 2; k >= 1; --k) r[k] = __fadd2_rn(r[k - 1], t[k][B]); \
        r[0] = t[0][B];                                             \
    }
            switch (b) {
            case 0: STEP(0); break;
            case 1: STEP(1); break;
            case 2: STEP(2); break;
            default: STEP(3); break;
            }
            const bool hit = fmaxf(o.x, o.y) > th;
            if (__any_sync(0xffffffffu, hit)) {
                unsigned m = __ballot_sync(0xffffffffu, hit);
                acc += __popc(m);
            }
        }
        w = wn;
    }
    if (lane == 0) out[warp] = acc;
}

template <typename F> float timeit(F f, int reps = 5) {
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    f(); CK(cudaDeviceSynchronize());
    float best = 1e30f;
    for (int r = 0; r < reps; ++r) {
        cudaEventRecord(e0); f(); cudaEventRecord(e1); CK(cudaEventSynchronize(e1));
        float ms; cudaEventElapsedTime(&ms, e0, e1); if (ms < best) best = ms;
    }
    return best;
}

int main() {
    cudaDeviceProp pr; CK(cudaGetDeviceProperties(&pr, 0));
    printf("dev %s sms %d clock %d kHz\n", pr.name, pr.multiProcessorCount, pr.clockRate);
    float* out; CK(cudaMalloc(&out, 148 * 64 * 256 * 4));
    const int iters = 20000;
    for (int bps : {2, 4, 8}) {
        int grid = 148 * bps;
        float ms = timeit([&] { k_fadd<<<grid, 256>>>(out, iters, 1.0f); });
        double adds = (double)grid * 256 * 16 * iters;
        printf("FADD  bps=%d  %.3f ms  %.3e add/s\n", bps, ms, adds / ms * 1e3);
        ms = timeit([&] { k_fadd2<<<grid, 256>>>(out, iters, 1.0f); });
        printf("FADD2 bps=%d  %.3f ms  %.3e add/s\n", bps, ms, 2 * adds / ms * 1e3);
        ms = timeit([&] { k_mix<<<grid, 256>>>(out, iters, 1.0f); });
        printf("MIX   bps=%d  %.3f ms  %.3e add/s (12 FADD2 + 16 int per iter)\n", bps, ms, (double)grid * 256 * 24 * iters / ms * 1e3);
    }
    // scan proto
    const int nwords = 256;  // 4096 bases per warp-item
    std::vector<uint32_t> hp((size_t)4096 * nwords + 16);
    srand(1); for (auto& v : hp) v = ((uint32_t)rand() << 16) ^ rand();
    uint32_t* dp; CK(cudaMalloc(&dp, hp.size() * 4)); CK(cudaMemcpy(dp, hp.data(), hp.size() * 4, cudaMemcpyHostToDevice));
    std::vector<float> ht(32 * 4 * 32 * 2);
    for (auto& v : ht) v = (rand() / (float)RAND_MAX) * 2.f - 1.f;
    float2* dt; CK(cudaMalloc(&dt, ht.size() * 4)); CK(cudaMemcpy(dt, ht.data(), ht.size() * 4, cudaMemcpyHostToDevice));
    unsigned long long* dout; CK(cudaMalloc(&dout, 1 << 22));
    for (float thv : {1e9f, 6.2f, 5.0f}) {
        std::vector<float> hth(32, thv); float* dth; CK(cudaMalloc(&dth, 128)); CK(cudaMemcpy(dth, hth.data(), 128, cudaMemcpyHostToDevice));
        for (int wps : {8, 12, 14}) {
            int nwarps = 148 * wps * 4;
            double la = (double)nwarps * nwords * 16 * 32 * 2.0;
            float ms = timeit([&] { scan_proto<12, 32><<<nwarps, 32>>>(dp, dt, dth, nwords, dout); });
            std::vector<unsigned long long> ho(nwarps); CK(cudaMemcpy(ho.data(), dout, nwarps * 8, cudaMemcpyDeviceToHost));
            unsigned long long tot = 0; for (auto v : ho) tot += v;
            printf("scan W=12 NT=32  thr=%g warps/SM=%d: %.3f ms  %.3e lookup-add/s  hits/cell=%.2e\n", thv, wps, ms, la * 12 / ms * 1e3, tot / (la));
            ms = timeit([&] { scan_proto<12, 128><<<nwarps / 4, 128>>>(dp, dt, dth, nwords, dout); });
            printf("scan W=12 NT=128 thr=%g warps/SM=%d: %.3f ms  %.3e lookup-add/s\n", thv, wps, ms, la * 12 / ms * 1e3);
            ms = timeit([&] { scan_proto<8, 32><<<nwarps, 32>>>(dp, dt, dth, nwords, dout); });
            printf("scan W=8  NT=32  thr=%g warps/SM=%d: %.3f ms  %.3e lookup-add/s\n", thv, wps, ms, la * 8 / ms * 1e3);
        }
    }
    printf("done\n");
    return 0;
}
