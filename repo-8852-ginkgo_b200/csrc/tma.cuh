// tma.cuh — 1-D bulk asynchronous copies (the TMA engine's non-tensor path), mbarrier
// completion and L2 prefetch, as raw PTX for sm_100a.  SASS: UBLKCP.S.G / UBLKPF.L /
// SYNCS.
#pragma once
#include "common.cuh"

namespace gkob200 {
namespace {

__host__ __device__ inline size_t align16(size_t b) { return (b + 15) & ~static_cast<size_t>(15); }

// ---- bulk asynchronous copy (TMA engine, 1-D) + mbarrier, raw PTX ------------
__device__ __forceinline__ unsigned smem_u32(const void* p) { return static_cast<unsigned>(__cvta_generic_to_shared(p)); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, unsigned count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, unsigned bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, unsigned parity)
{
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "WAIT_%=:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra DONE_%=;\n"
        "bra WAIT_%=;\n"
        "DONE_%=:\n"
        "}\n" ::"r"(smem_u32(bar)),
        "r"(parity)
        : "memory");
}
// global -> shared, `bytes` a multiple of 16, both addresses 16-byte aligned; completion
// is signalled on `bar` as transaction bytes.  SASS: UBLKCP.
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, unsigned bytes, uint64_t* bar)
{
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                     smem_u32(dst)),
                 "l"(src), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}

// L2 prefetch of a byte range (16-byte aligned start, size a multiple of 16): the bulk
// copies of a LATER tile then find their data in L2 instead of paying DRAM latency, which
// raises the bytes in flight beyond what fits in shared memory.
__device__ __forceinline__ void bulk_prefetch_l2(const void* src, unsigned bytes)
{
    asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(src), "r"(bytes) : "memory");
}


// ---- L2 residency hints ------------------------------------------------------------
// B200's 126 MB L2 can hold the whole gathered vector x of a 10^7-row problem (80 MB) if
// the streamed matrix arrays do not evict it: the streams are loaded with an evict_first
// policy, the gathers with evict_last.
__device__ __forceinline__ uint64_t l2_policy_evict_first()
{
    uint64_t p;
    asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(p));
    return p;
}
__device__ __forceinline__ uint64_t l2_policy_evict_last(float fraction = 1.0f)
{
    uint64_t p;
    asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, %1;" : "=l"(p) : "f"(fraction));
    return p;
}
__device__ __forceinline__ double ld_hint(const double* a, uint64_t pol)
{
    double v;
    asm volatile("ld.global.nc.L2::cache_hint.f64 %0, [%1], %2;" : "=d"(v) : "l"(a), "l"(pol));
    return v;
}
__device__ __forceinline__ float ld_hint(const float* a, uint64_t pol)
{
    float v;
    asm volatile("ld.global.nc.L2::cache_hint.f32 %0, [%1], %2;" : "=f"(v) : "l"(a), "l"(pol));
    return v;
}
__device__ __forceinline__ int32_t ld_hint(const int32_t* a, uint64_t pol)
{
    int32_t v;
    asm volatile("ld.global.nc.L2::cache_hint.s32 %0, [%1], %2;" : "=r"(v) : "l"(a), "l"(pol));
    return v;
}
__device__ __forceinline__ int64_t ld_hint(const int64_t* a, uint64_t pol)
{
    int64_t v;
    asm volatile("ld.global.nc.L2::cache_hint.s64 %0, [%1], %2;" : "=l"(v) : "l"(a), "l"(pol));
    return v;
}

}  // namespace
}  // namespace gkob200
