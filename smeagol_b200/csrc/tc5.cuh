// Thin inline-PTX wrappers for the sm_100a tensor-core path (tcgen05.mma with TMEM accumulators, mbarrier completion,
// TMEM allocation, TMEM -> register loads).  Used by scan_tc5.cuh (the product prefilter) and by
// tools/tcgen05_probe.cu (descriptor / layout known-answer test and the TMEM read-bandwidth microbenchmark).
//
// Shared-memory matrix descriptor (64 bit), SWIZZLE_NONE, K-major operand, in units of 16 bytes:
//   element (row r, k) of an operand of 16-bit elements lives at
//       start + (r % 8) * 16 B + (r / 8) * SBO + ((k / 8) % 2) * LBO + (k % 8) * 2 B       (one K = 16 instruction)
//   bits [0,14) start >> 4, [16,30) LBO >> 4, [32,46) SBO >> 4, [46,48) version = 1 (Blackwell), [61,64) swizzle = 0.
// Instruction descriptor (32 bit), kind::f16: bits [4,6) D format (1 = f32), [7,10) A format (0 = f16), [10,13) B format,
//   bit 15 / 16 A / B major (0 = K-major), [17,23) N >> 3, [24,29) M >> 4.
#pragma once
#include <cstdint>

namespace tc5 {

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__host__ __device__ constexpr uint64_t make_smem_desc(const uint32_t saddr, const uint32_t lbo_bytes, const uint32_t sbo_bytes)
{
    return (uint64_t)((saddr >> 4) & 0x3fffu) | ((uint64_t)((lbo_bytes >> 4) & 0x3fffu) << 16) |
           ((uint64_t)((sbo_bytes >> 4) & 0x3fffu) << 32) | (1ull << 46);
}

__host__ __device__ constexpr uint32_t make_idesc_f16_f32(const int m, const int n)
{
    return (1u << 4) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(m >> 4) << 24);
}

// ---- TMEM allocation (one warp, converged); the base address lands in shared memory
__device__ __forceinline__ void tmem_alloc(uint32_t* smem_dst, const uint32_t ncols)
{
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(smem_dst)), "r"(ncols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(const uint32_t taddr, const uint32_t ncols)
{
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void fence_before_sync() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void fence_after_sync() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
// generic-proxy shared-memory writes -> visible to the async proxy (tensor core operand reads)
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

// ---- mbarrier
__device__ __forceinline__ void mbar_init(uint64_t* bar, const uint32_t count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_fence_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void mbar_arrive(uint64_t* bar)
{
    asm volatile("{\n\t.reg .b64 st;\n\tmbarrier.arrive.shared::cta.b64 st, [%0];\n\t}" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, const uint32_t parity)
{
    uint32_t ok;
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                 : "=r"(ok)
                 : "r"(smem_u32(bar)), "r"(parity)
                 : "memory");
    return ok != 0;
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, const uint32_t parity)
{
    while (!mbar_try_wait(bar, parity)) {}
}

// ---- MMA: D[tmem] (+)= A[smem] * B[smem]^T, M x N x 16, one thread issues
__device__ __forceinline__ void mma_f16_ss(const uint32_t tmem_d, const uint64_t desc_a, const uint64_t desc_b, const uint32_t idesc,
                                           const uint32_t accumulate)
{
    asm volatile(
        "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(tmem_d), "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
        : "memory");
}
// all MMAs issued so far by this thread arrive on the mbarrier when they complete (implies fence::before_thread_sync)
__device__ __forceinline__ void mma_commit(uint64_t* bar)
{
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}

__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

#include "tc5_ld.inc"

__device__ __forceinline__ void tmem_st_32x32b_x8(const uint32_t taddr, const uint32_t (&r)[8])
{
    asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"r"(taddr), "r"(r[0]), "r"(r[1]),
                 "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7])
                 : "memory");
}

}  // namespace tc5
