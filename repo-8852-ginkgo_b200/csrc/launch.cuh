// launch.cuh — grid-stride elementwise launcher over an n x k row-major view.
// (Counterpart of the reference's run_kernel / run_kernel_solver,
//  common/cuda_hip/base/kernel_launch.hpp.inc:33-82, built for 148 SMs:
//  grid = min(work, 148 * 8 CTAs) instead of one thread per element.)
#pragma once
#include "common.cuh"

namespace gkob200 {

template <typename F>
__global__ void __launch_bounds__(256) elementwise_2d(int64_t n, int64_t k, F f)
{
    const int64_t total = n * k;
    const int64_t step = static_cast<int64_t>(gridDim.x) * blockDim.x;
    if (k == 1) {
        for (int64_t i = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; i < total; i += step)
            f(i, int64_t(0));
    } else {
        for (int64_t t = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; t < total; t += step)
            f(t / k, t % k);
    }
}

template <typename F>
inline int launch_2d(cudaStream_t s, int64_t n, int64_t k, F f)
{
    if (n <= 0 || k <= 0) return 0;
    elementwise_2d<<<grid_for(n * k, 256, 8), 256, 0, s>>>(n, k, f);
    GKOB200_CHECK_LAUNCH();
    return 0;
}

// Column-wise reduction of an n x k view: result[j] = fin(sum_i term(i, j)).
// One launch (single-pass grid reduction).  For k == 1 all threads stream rows;
// for k > 1 a thread owns column (tid % k') of a group of rows so that accesses
// stay coalesced along the row-major rows.
template <typename T, typename Term, typename Fin>
__global__ void __launch_bounds__(256) col_reduce_1(int64_t n, Term term, Fin fin, T* partials, unsigned* ticket)
{
    T v[1] = {T(0)};
    const int64_t step = static_cast<int64_t>(gridDim.x) * blockDim.x;
    for (int64_t i = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; i < n; i += step)
        v[0] += term(i, int64_t(0));
    grid_reduce<1>(v, partials, ticket, [&](T(&tot)[1]) { fin(int64_t(0), tot[0]); });
}

// general k: one launch per call, block handles all columns; thread (tid) maps to
// column tid % kk and row lane tid / kk where kk = min(k, 32) columns per pass.
template <typename T, typename Term, typename Fin>
__global__ void __launch_bounds__(256) col_reduce_k(int64_t n, int64_t k, Term term, Fin fin, T* partials,
                                                    unsigned* ticket)
{
    // Each block walks rows blockIdx.x, blockIdx.x + gridDim.x, ...; thread t owns
    // columns t, t + 256, ... of those rows (k <= 256 typical: one column each).
    __shared__ bool is_last;
    for (int64_t j = threadIdx.x; j < k; j += blockDim.x) {
        T s = T(0);
        for (int64_t i = blockIdx.x; i < n; i += gridDim.x) s += term(i, j);
        partials[static_cast<size_t>(blockIdx.x) * k + j] = s;
    }
    __threadfence();
    __syncthreads();
    if (threadIdx.x == 0) {
        unsigned t = atomicAdd(ticket, 1u);
        is_last = (t == gridDim.x - 1);
    }
    __syncthreads();
    if (!is_last) return;
    __threadfence();
    for (int64_t j = threadIdx.x; j < k; j += blockDim.x) {
        T s = T(0);
        for (unsigned b = 0; b < gridDim.x; ++b) s += __ldcg(&partials[static_cast<size_t>(b) * k + j]);
        fin(j, s);
    }
    if (threadIdx.x == 0) *ticket = 0u;
}

template <typename T, typename Term, typename Fin>
inline int launch_col_reduce(cudaStream_t s, int64_t n, int64_t k, void* ws, Term term, Fin fin)
{
    if (k <= 0) return 0;
    if (!ws) return GKOB200_EINVAL;
    if (k == 1) {
        col_reduce_1<T><<<grid_for(n, 256, 4), 256, 0, s>>>(n, term, fin, ws_partials<T>(ws), ws_ticket(ws));
    } else {
        // partials: grid * k values must fit the workspace
        int64_t grid = static_cast<int64_t>(kReduceMaxBlocks) * kReduceMaxVals / k;
        if (grid < 1) return GKOB200_EUNSUPPORTED;  // k > 18944 columns
        const int64_t cap = static_cast<int64_t>(sm_count()) * 8;
        if (grid > cap) grid = cap;
        if (grid > n) grid = n > 0 ? n : 1;
        col_reduce_k<T><<<static_cast<unsigned>(grid), 256, 0, s>>>(n, k, term, fin, ws_partials<T>(ws),
                                                                   ws_ticket(ws));
    }
    GKOB200_CHECK_LAUNCH();
    return 0;
}

}  // namespace gkob200
