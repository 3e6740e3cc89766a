// common.cuh — shared device/host helpers for the sm_100a sparse-solve kernels.
//
// Everything in csrc/ is hand-written for Blackwell B200 (sm_100a).  The
// kernels are HBM-bound SIMT kernels (SpMV ~0.17 flop/B, BLAS-1 <= 0.125 flop/B)
// so no tensor-core path exists here by design; the levers are coalesced /
// vectorised streams, shared-memory (and bulk-async) staging of row blocks,
// grids sized in multiples of the SM count and single-pass grid reductions.
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/gko_b200.h"
#include "p2p.cuh"

namespace gkob200 {

// ---- error handling --------------------------------------------------------
// C-ABI functions never throw: they return 0 or a cudaError_t / negative code.
#define GKOB200_CHECK_LAUNCH()                                   \
    do {                                                         \
        cudaError_t e__ = cudaPeekAtLastError();                 \
        if (e__ != cudaSuccess) return static_cast<int>(e__);    \
    } while (0)

#define GKOB200_CUDA(call)                                       \
    do {                                                         \
        cudaError_t e__ = (call);                                \
        if (e__ != cudaSuccess) return static_cast<int>(e__);    \
    } while (0)

inline cudaStream_t as_stream(void* s) { return reinterpret_cast<cudaStream_t>(s); }

// Number of SMs of the current device (148 on B200), cached per device.
int sm_count();
// Largest opt-in dynamic shared memory per block of the current device.
int max_smem_optin();

inline int64_t ceildiv(int64_t a, int64_t b) { return (a + b - 1) / b; }

// Grid size for a grid-stride elementwise/reduction kernel: a multiple of the
// SM count, capped by the amount of work.
inline int grid_for(int64_t work_items, int block, int ctas_per_sm)
{
    int64_t want = ceildiv(work_items, block);
    int64_t cap = static_cast<int64_t>(sm_count()) * ctas_per_sm;
    if (want < 1) want = 1;
    return static_cast<int>(want < cap ? want : cap);
}

// ---- round-to-nearest, never-contracted arithmetic -------------------------
// The parity oracle (Ginkgo's reference executor built for baseline x86-64)
// evaluates  c += a * b  as a rounded product followed by a rounded sum.  Where
// a kernel keeps the oracle's summation order we also keep its rounding, which
// makes those paths bit-identical to the oracle, by using the _rn intrinsics
// (nvcc never fuses them into FMA).
__device__ __forceinline__ double mul_rn(double a, double b) { return __dmul_rn(a, b); }
__device__ __forceinline__ float mul_rn(float a, float b) { return __fmul_rn(a, b); }
__device__ __forceinline__ double add_rn(double a, double b) { return __dadd_rn(a, b); }
__device__ __forceinline__ float add_rn(float a, float b) { return __fadd_rn(a, b); }
__device__ __forceinline__ double sub_rn(double a, double b) { return __dsub_rn(a, b); }
__device__ __forceinline__ float sub_rn(float a, float b) { return __fsub_rn(a, b); }
__device__ __forceinline__ double div_rn(double a, double b) { return __ddiv_rn(a, b); }
__device__ __forceinline__ float div_rn(float a, float b) { return __fdiv_rn(a, b); }
__device__ __forceinline__ double sqrt_rn(double a) { return __dsqrt_rn(a); }
__device__ __forceinline__ float sqrt_rn(float a) { return __fsqrt_rn(a); }

template <typename T>
__device__ __forceinline__ T ldg(const T* p)
{
    return __ldg(p);
}

// gko::stopping_status is one byte: bit7 converged, bit6 finalized, bits0-5 id
// (reference include/ginkgo/core/stop/stopping_status.hpp:104-108).
__host__ __device__ __forceinline__ bool status_has_stopped(uint8_t s) { return (s & 0x3f) != 0; }
__host__ __device__ __forceinline__ bool status_is_finalized(uint8_t s) { return (s & 0x40) != 0; }

// ---- reductions -------------------------------------------------------------
template <typename T>
__device__ __forceinline__ T warp_sum(T v)
{
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o);
    return v;
}

// Sum over the block; result valid in thread 0.  `red` is >= 32 T of shared mem.
template <typename T>
__device__ __forceinline__ T block_sum(T v, T* red)
{
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    v = warp_sum(v);
    __syncthreads();  // protect `red` from a previous use
    if (lane == 0) red[wid] = v;
    __syncthreads();
    const int nw = (blockDim.x + 31) >> 5;
    if (wid == 0) {
        v = lane < nw ? red[lane] : T(0);
        v = warp_sum(v);
    }
    return v;
}

// Single-pass deterministic grid reduction of NV running sums.
//
// Every block leaves its NV block-sums in `partials[blockIdx.x*NV + i]`, then takes a
// ticket; the block that draws the last ticket re-reads the partials in a fixed order
// (independent of which block happens to be last), reduces them and hands the NV totals
// to `fin(totals)` on its thread 0, then re-arms the ticket.  One launch, no atomics on
// data, run-to-run bit-reproducible for a fixed grid size.  (The reference needs two
// launches + a tmp array: common/cuda_hip/base/kernel_launch_reduction.hpp.inc:33-330.)
//
// Large grids (the SpMV with a fused dot runs one CTA per 128 rows: 62 500 CTAs at 200^3)
// use TWO ticket levels — groups of >= 1024 CTAs, then the groups — because tens of
// thousands of atomics on one address serialise in L2 (measured: +100 us per launch).
constexpr unsigned kTicketGroup = 1024;
constexpr unsigned kMaxTicketGroups = 255;  // ticket words 1..255 of the 1 KB header

template <int NV, typename T, typename Fin>
__device__ __forceinline__ void grid_reduce(T (&v)[NV], T* partials, unsigned* ticket, Fin fin)
{
    __shared__ T red[32];
    __shared__ bool is_last;
#pragma unroll
    for (int i = 0; i < NV; ++i) {
        T s = block_sum(v[i], red);
        if (threadIdx.x == 0) partials[static_cast<size_t>(blockIdx.x) * NV + i] = s;
    }
    const unsigned grid = gridDim.x;
    if (grid <= kTicketGroup) {
        if (threadIdx.x == 0) {
            __threadfence();
            is_last = (atomicAdd(ticket, 1u) == grid - 1);
        }
        __syncthreads();
        if (!is_last) return;
        __threadfence();
        T tot[NV];
#pragma unroll
        for (int i = 0; i < NV; ++i) {
            T s = T(0);
            for (unsigned b = threadIdx.x; b < grid; b += blockDim.x)
                s += __ldcg(&partials[static_cast<size_t>(b) * NV + i]);
            tot[i] = block_sum(s, red);
        }
        if (threadIdx.x == 0) {
            *ticket = 0u;
            fin(tot);
        }
        return;
    }
    // ---- two levels ---------------------------------------------------------------
    unsigned gsize = (grid + kMaxTicketGroups - 1) / kMaxTicketGroups;
    if (gsize < kTicketGroup) gsize = kTicketGroup;
    const unsigned ngroups = (grid + gsize - 1) / gsize;
    const unsigned g = blockIdx.x / gsize;
    const unsigned gbegin = g * gsize;
    const unsigned gend = gbegin + gsize < grid ? gbegin + gsize : grid;
    if (threadIdx.x == 0) {
        __threadfence();
        is_last = (atomicAdd(ticket + 1 + g, 1u) == gend - gbegin - 1);
    }
    __syncthreads();
    if (!is_last) return;
    __threadfence();
#pragma unroll
    for (int i = 0; i < NV; ++i) {
        T s = T(0);
        for (unsigned b = gbegin + threadIdx.x; b < gend; b += blockDim.x)
            s += __ldcg(&partials[static_cast<size_t>(b) * NV + i]);
        s = block_sum(s, red);
        if (threadIdx.x == 0) partials[static_cast<size_t>(grid + g) * NV + i] = s;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        ticket[1 + g] = 0u;
        __threadfence();
        is_last = (atomicAdd(ticket, 1u) == ngroups - 1);
    }
    __syncthreads();
    if (!is_last) return;
    __threadfence();
    T tot[NV];
#pragma unroll
    for (int i = 0; i < NV; ++i) {
        T s = T(0);
        for (unsigned b = threadIdx.x; b < ngroups; b += blockDim.x)
            s += __ldcg(&partials[static_cast<size_t>(grid + b) * NV + i]);
        tot[i] = block_sum(s, red);
    }
    if (threadIdx.x == 0) {
        *ticket = 0u;
        fin(tot);
    }
}

// Deferred variant for kernels whose CTAs must retire as fast as possible (the SpMV with a
// fused dot holds 44 KB of shared memory per CTA: waiting for the ticket atomic's round trip
// before exiting cost ~15 % of its throughput).  The kernel only stores one partial per CTA
// — fire and forget — and `finish_partials` (one CTA, fixed order, deterministic) sums them.
template <typename T>
__device__ __forceinline__ void store_block_partial(T v, T* partials)
{
    __shared__ T red_p[32];
    const T s = block_sum(v, red_p);
    if (threadIdx.x == 0) partials[blockIdx.x] = s;
}
// Same, one partial per WARP and no barrier at all (the row-block SpMV on short rows: its CTAs
// live ~3 us, two __syncthreads per CTA were 8 % of that).  Warp w of CTA b owns slot
// b * warps + w; a second reduction uses the array right behind (gridDim.x * warps entries).
template <typename T>
__device__ __forceinline__ void store_warp_partial(T v, T* partials, int warps_per_cta)
{
    const T s = warp_sum(v);
    if ((threadIdx.x & 31) == 0) partials[static_cast<size_t>(blockIdx.x) * warps_per_cta + (threadIdx.x >> 5)] = s;
}
template <typename T>
__device__ __forceinline__ void store_warp_partial2(T v0, T v1, T* partials, int warps_per_cta)
{
    const T s0 = warp_sum(v0), s1 = warp_sum(v1);
    if ((threadIdx.x & 31) == 0) {
        const size_t slot = static_cast<size_t>(blockIdx.x) * warps_per_cta + (threadIdx.x >> 5);
        partials[slot] = s0;
        partials[static_cast<size_t>(gridDim.x) * warps_per_cta + slot] = s1;
    }
}
// two reductions per CTA: the second array follows the first (gridDim.x entries each)
template <typename T>
__device__ __forceinline__ void store_block_partial2(T v0, T v1, T* partials)
{
    __shared__ T red_p2[2][32];
    const T s0 = block_sum(v0, red_p2[0]);
    const T s1 = block_sum(v1, red_p2[1]);
    if (threadIdx.x == 0) {
        partials[blockIdx.x] = s0;
        partials[gridDim.x + blockIdx.x] = s1;
    }
}

// Second stage of the deferred reduction.  gridDim.x = G CTAs sum G fixed contiguous chunks of
// the per-CTA partials (both arrays when out2 != nullptr: the second array follows the first,
// n entries each); the CTA that draws the last ticket adds the G chunk sums in chunk order,
// writes the result(s) and — inside a distributed solver — all-reduces `p2p_count` values at
// `p2p_buf` over peer memory (p2p.cuh), so SpMV + dot + all-reduce are two launches.
// Deterministic for a fixed (n, G): the association does not depend on which CTA is last.
// (Round 1 used ONE CTA: 8 us for the 62 500 partials of a 200^3 SpMV, on the critical path of
// every iteration.)  `chunk_sums` has room for 2 * G values, `ticket` is one zeroed word.
constexpr int kFinishThreads = 512;
constexpr int kFinishMaxCtas = 64;
inline int finish_grid(int64_t n)
{
    int64_t g = (n + 2047) / 2048;
    return static_cast<int>(g < 1 ? 1 : g > kFinishMaxCtas ? kFinishMaxCtas : g);
}

template <typename T>
__global__ void __launch_bounds__(kFinishThreads)
    finish_partials(int64_t n, const T* __restrict__ partials, T* out, const int* skip, T* out2, T* chunk_sums,
                    unsigned* ticket, const P2pDev* p2p, T* p2p_buf, int p2p_count, int* on_fail)
{
    if (skip && *skip) return;
    __shared__ T red[32];
    __shared__ bool is_last;
    const int G = gridDim.x;
    const int64_t chunk = (n + G - 1) / G;
    const int64_t lo = blockIdx.x * chunk, hi = lo + chunk < n ? lo + chunk : n;
    const int nv = out2 ? 2 : 1;
    for (int a = 0; a < nv; ++a) {
        const T* src = partials + a * n;
        T a0 = T(0), a1 = T(0), a2 = T(0), a3 = T(0);
        int64_t i = lo + threadIdx.x;
        for (; i + 3 * kFinishThreads < hi; i += 4 * kFinishThreads) {
            a0 += src[i];
            a1 += src[i + kFinishThreads];
            a2 += src[i + 2 * kFinishThreads];
            a3 += src[i + 3 * kFinishThreads];
        }
        for (; i < hi; i += kFinishThreads) a0 += src[i];
        const T s = block_sum((a0 + a1) + (a2 + a3), red);
        if (threadIdx.x == 0) chunk_sums[a * G + blockIdx.x] = s;
    }
    if (G > 1) {
        if (threadIdx.x == 0) {
            __threadfence();
            is_last = (atomicAdd(ticket, 1u) == static_cast<unsigned>(G) - 1u);
        }
        __syncthreads();
        if (!is_last) return;
        __threadfence();
    }
    if (threadIdx.x == 0) {
        if (G > 1) *ticket = 0u;
        for (int a = 0; a < nv; ++a) {
            T tot = T(0);
            for (int g = 0; g < G; ++g) tot += __ldcg(&chunk_sums[a * G + g]);
            (a == 0 ? out : out2)[0] = tot;
        }
        if (p2p && !peer_allreduce(*p2p, p2p_buf, p2p_count) && on_fail) *on_fail = 1;
    }
}

// Scratch carried by every reducing kernel: ticket words followed by partials.
// Layout of a reduction workspace (bytes): [0,1024) tickets, [1024, ...) partial sums.
// GKOB200_REDUCE_WS_BYTES covers the largest grid the BLAS-1 kernels launch
// (<= 148*16 blocks) times 8 values; solvers that fuse a dot into the SpMV allocate
// reduce_ws_bytes(blocks) for that grid.
constexpr int kReduceHeader = 1024;
constexpr int kReduceMaxBlocks = 148 * 16;
constexpr int kReduceMaxVals = 8;
static_assert(GKOB200_REDUCE_WS_BYTES >= kReduceHeader + (kReduceMaxBlocks + 256) * kReduceMaxVals * 8, "ws");

// bytes for a grid of `blocks` CTAs (+ room for the group partials of the second level)
inline size_t reduce_ws_bytes(int64_t blocks)
{
    return static_cast<size_t>(kReduceHeader) + static_cast<size_t>(blocks + 256) * kReduceMaxVals * sizeof(double);
}

__host__ __device__ inline unsigned* ws_ticket(void* ws) { return reinterpret_cast<unsigned*>(ws); }
template <typename T>
__host__ __device__ inline T* ws_partials(void* ws)
{
    return reinterpret_cast<T*>(reinterpret_cast<char*>(ws) + kReduceHeader);
}

}  // namespace gkob200
