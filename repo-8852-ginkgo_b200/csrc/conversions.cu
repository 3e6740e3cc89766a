// conversions.cu — integer / index work around the formats: exclusive scan,
// ptrs<->idxs, slice sets, CSR -> ELL / SELL-P / Hybrid / COO, order statistics of the
// row lengths (Hybrid strategies) and the load_balance `srow` array.  All results are
// BIT-EXACT with the reference executor.
//
// [ref] core/components/prefix_sum_kernels.hpp:67 + reference/components/prefix_sum_kernels.cpp:40-52,
//       common/unified/components/format_conversion_kernels.cpp:49-112,
//       reference/matrix/sellp_kernels.cpp:134-160 (compute_slice_sets),
//       reference/matrix/ell_kernels.cpp:159-168 (compute_max_row_nnz),
//       common/unified/matrix/csr_kernels.cpp:137-243 (convert_to_{sellp,ell,hybrid}),
//       common/unified/matrix/hybrid_kernels.cpp:51-76 (compute_coo_row_ptrs),
//       include/ginkgo/core/matrix/hybrid.hpp:112-380 (strategies, host side),
//       include/ginkgo/core/matrix/csr.hpp:421-511 (load_balance::process / clac_size).
#include "launch.cuh"

namespace gkob200 {
namespace {

// ---------------------------------------------------------------------------
// exclusive prefix sum, in place, n entries (the last input entry is ignored and
// receives the total, as in the reference).  Three phases over tiles of 4096:
// tile sums -> scan of tile sums (recursive, one more level is enough for 2^36
// entries) -> rescan with offsets.  Integer addition: exact, order-independent.
// ---------------------------------------------------------------------------
constexpr int kScanThreads = 512;
constexpr int kScanItems = 8;
constexpr int kScanTile = kScanThreads * kScanItems;

template <typename T>
__device__ __forceinline__ T block_exclusive_scan(T v, T* smem, T& total)
{
    // warp scan then scan of warp totals
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    T incl = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const T up = __shfl_up_sync(0xffffffffu, incl, o);
        if (lane >= o) incl += up;
    }
    if (lane == 31) smem[wid] = incl;
    __syncthreads();
    if (wid == 0) {
        const int nw = kScanThreads / 32;
        T w = lane < nw ? smem[lane] : T(0);
        T wi = w;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const T up = __shfl_up_sync(0xffffffffu, wi, o);
            if (lane >= o) wi += up;
        }
        if (lane < nw) smem[lane] = wi - w;  // exclusive warp offsets
        if (lane == nw - 1) smem[32] = wi;   // block total
    }
    __syncthreads();
    total = smem[32];
    const T res = smem[wid] + incl - v;
    __syncthreads();
    return res;
}

template <typename T>
__global__ void __launch_bounds__(kScanThreads) scan_tile_sums(const T* __restrict__ in, int64_t n, T* __restrict__ sums)
{
    __shared__ T red[32];
    const int64_t base = static_cast<int64_t>(blockIdx.x) * kScanTile;
    T s = T(0);
    for (int k = threadIdx.x; k < kScanTile; k += kScanThreads)
        if (base + k < n) s += in[base + k];
    // block sum (integer)
    for (int o = 16; o > 0; o >>= 1) s += __shfl_down_sync(0xffffffffu, s, o);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = s;
    __syncthreads();
    if (threadIdx.x < 32) {
        s = threadIdx.x < kScanThreads / 32 ? red[threadIdx.x] : T(0);
        for (int o = 16; o > 0; o >>= 1) s += __shfl_down_sync(0xffffffffu, s, o);
        if (threadIdx.x == 0) sums[blockIdx.x] = s;
    }
}

template <typename T>
__global__ void __launch_bounds__(kScanThreads)
    scan_tiles(T* __restrict__ data, int64_t n, const T* __restrict__ tile_offsets)
{
    __shared__ T smem[33];
    const int64_t base = static_cast<int64_t>(blockIdx.x) * kScanTile + static_cast<int64_t>(threadIdx.x) * kScanItems;
    T v[kScanItems];
    T local = T(0);
#pragma unroll
    for (int i = 0; i < kScanItems; ++i) {
        v[i] = base + i < n ? data[base + i] : T(0);
        local += v[i];
    }
    T total;
    T off = block_exclusive_scan(local, smem, total) + (tile_offsets ? tile_offsets[blockIdx.x] : T(0));
#pragma unroll
    for (int i = 0; i < kScanItems; ++i) {
        if (base + i < n) data[base + i] = off;
        off += v[i];
    }
}

template <typename T>
int prefix_sum_impl(cudaStream_t s, T* data, int64_t n, void* ws, size_t ws_bytes)
{
    if (n < 0) return GKOB200_EINVAL;
    if (n == 0) return 0;
    if (!data) return GKOB200_EINVAL;
    const int64_t tiles = ceildiv(n, kScanTile);
    if (tiles == 1) {
        scan_tiles<T><<<1, kScanThreads, 0, s>>>(data, n, nullptr);
        GKOB200_CHECK_LAUNCH();
        return 0;
    }
    const int64_t tiles2 = ceildiv(tiles, kScanTile);
    if (!ws || ws_bytes < static_cast<size_t>(tiles + tiles2 + 2) * sizeof(T)) return GKOB200_EWORKSPACE;
    T* sums = reinterpret_cast<T*>(ws);
    T* sums2 = sums + tiles + 1;
    scan_tile_sums<T><<<static_cast<unsigned>(tiles), kScanThreads, 0, s>>>(data, n, sums);
    GKOB200_CHECK_LAUNCH();
    if (tiles2 == 1) {
        scan_tiles<T><<<1, kScanThreads, 0, s>>>(sums, tiles, nullptr);
    } else {
        if (tiles2 > kScanTile) return GKOB200_EUNSUPPORTED;
        scan_tile_sums<T><<<static_cast<unsigned>(tiles2), kScanThreads, 0, s>>>(sums, tiles, sums2);
        scan_tiles<T><<<1, kScanThreads, 0, s>>>(sums2, tiles2, nullptr);
        scan_tiles<T><<<static_cast<unsigned>(tiles2), kScanThreads, 0, s>>>(sums, tiles, sums2);
    }
    GKOB200_CHECK_LAUNCH();
    scan_tiles<T><<<static_cast<unsigned>(tiles), kScanThreads, 0, s>>>(data, n, sums);
    GKOB200_CHECK_LAUNCH();
    return 0;
}

// ---------------------------------------------------------------------------
template <typename I>
__global__ void __launch_bounds__(256)
    ptrs_to_idxs_kernel(const I* __restrict__ ptrs, int64_t n_rows, I* __restrict__ idxs)
{
    // 8 lanes per row: rows are short, writes stay mostly contiguous
    const int64_t gid = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x;
    const int64_t row = gid >> 3;
    if (row >= n_rows) return;
    const I e = ptrs[row + 1];
    for (I k = ptrs[row] + static_cast<I>(gid & 7); k < e; k += 8) idxs[k] = static_cast<I>(row);
}

template <typename I>
__global__ void slice_lengths_kernel(const I* __restrict__ row_ptrs, int64_t n_rows, int64_t slice_size,
                                     int64_t stride_factor, int64_t n_slices, uint64_t* __restrict__ slice_sets,
                                     uint64_t* __restrict__ slice_lengths)
{
    // one warp per slice
    const int64_t slice = (blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (slice >= n_slices) return;
    uint64_t m = 0;
    for (int64_t lr = lane; lr < slice_size; lr += 32) {
        const int64_t row = slice * slice_size + lr;
        const uint64_t len = row < n_rows ? static_cast<uint64_t>(row_ptrs[row + 1] - row_ptrs[row]) : 0;
        const uint64_t padded = (len + stride_factor - 1) / stride_factor * stride_factor;
        m = max(m, padded);
    }
    for (int o = 16; o > 0; o >>= 1) m = max(m, __shfl_down_sync(0xffffffffu, m, o));
    if (lane == 0) {
        slice_lengths[slice] = m;
        slice_sets[slice] = m;
    }
}

template <typename I>
__global__ void max_row_nnz_kernel(const I* __restrict__ row_ptrs, int64_t n_rows, unsigned long long* out)
{
    unsigned long long m = 0;
    for (int64_t r = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; r < n_rows;
         r += static_cast<int64_t>(gridDim.x) * blockDim.x)
        m = max(m, static_cast<unsigned long long>(row_ptrs[r + 1] - row_ptrs[r]));
    for (int o = 16; o > 0; o >>= 1) m = max(m, __shfl_down_sync(0xffffffffu, m, o));
    if ((threadIdx.x & 31) == 0) atomicMax(out, m);
}

// rows with length in [lo, lo + width*bins) counted into `bins` buckets of `width`
template <typename I>
__global__ void row_len_histogram(const I* __restrict__ row_ptrs, int64_t n_rows, uint64_t lo, uint64_t width,
                                  int bins, unsigned long long* __restrict__ hist)
{
    extern __shared__ unsigned int s_hist[];
    for (int i = threadIdx.x; i < bins; i += blockDim.x) s_hist[i] = 0;
    __syncthreads();
    for (int64_t r = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; r < n_rows;
         r += static_cast<int64_t>(gridDim.x) * blockDim.x) {
        const uint64_t len = static_cast<uint64_t>(row_ptrs[r + 1] - row_ptrs[r]);
        if (len >= lo) {
            const uint64_t bkt = (len - lo) / width;
            if (bkt < static_cast<uint64_t>(bins)) atomicAdd(&s_hist[bkt], 1u);
        }
    }
    __syncthreads();
    for (int i = threadIdx.x; i < bins; i += blockDim.x)
        if (s_hist[i]) atomicAdd(&hist[i], static_cast<unsigned long long>(s_hist[i]));
}

}  // namespace
}  // namespace gkob200

using namespace gkob200;

extern "C" {

size_t gkob200_prefix_sum_workspace_bytes(int64_t n)
{
    const int64_t tiles = ceildiv(n, kScanTile);
    return static_cast<size_t>(tiles + ceildiv(tiles, kScanTile) + 4) * sizeof(uint64_t);
}
int gkob200_prefix_sum_i32(void* st, int32_t* data, int64_t n, void* ws, size_t wsb)
{
    return prefix_sum_impl<int32_t>(as_stream(st), data, n, ws, wsb);
}
int gkob200_prefix_sum_i64(void* st, int64_t* data, int64_t n, void* ws, size_t wsb)
{
    return prefix_sum_impl<int64_t>(as_stream(st), data, n, ws, wsb);
}
int gkob200_prefix_sum_u64(void* st, uint64_t* data, int64_t n, void* ws, size_t wsb)
{
    return prefix_sum_impl<uint64_t>(as_stream(st), data, n, ws, wsb);
}

#define GKOB200_DEF_CONV_I(I, IT)                                                                               \
    int gkob200_convert_ptrs_to_idxs_##I(void* st, const IT* ptrs, int64_t n, IT* idxs)                          \
    {                                                                                                           \
        if (n < 0) return GKOB200_EINVAL;                                                                       \
        if (n == 0) return 0;                                                                                   \
        if (!ptrs) return GKOB200_EINVAL;                                                                       \
        ptrs_to_idxs_kernel<IT><<<static_cast<unsigned>(ceildiv(n * 8, 256)), 256, 0, as_stream(st)>>>(ptrs, n, \
                                                                                                      idxs);    \
        GKOB200_CHECK_LAUNCH();                                                                                 \
        return 0;                                                                                               \
    }                                                                                                           \
    int gkob200_convert_idxs_to_ptrs_##I(void* st, const IT* idxs, int64_t num_idxs, int64_t n, IT* ptrs)        \
    {                                                                                                           \
        if (n < 0 || num_idxs < 0 || !ptrs) return GKOB200_EINVAL;                                              \
        if (num_idxs == 0)                                                                                      \
            return launch_2d(as_stream(st), n + 1, 1, [=] __device__(int64_t i, int64_t) { ptrs[i] = 0; });     \
        return launch_2d(as_stream(st), num_idxs + 1, 1, [=] __device__(int64_t i, int64_t) {                   \
            const int64_t begin = i == 0 ? 0 : static_cast<int64_t>(idxs[i - 1]);                               \
            const int64_t end = i == num_idxs ? n : static_cast<int64_t>(idxs[i]);                              \
            for (int64_t blk = begin; blk < end; ++blk) ptrs[blk + 1] = static_cast<IT>(i);                     \
            if (i == 0) ptrs[0] = 0;                                                                            \
        });                                                                                                     \
    }                                                                                                           \
    int gkob200_convert_ptrs_to_sizes_##I(void* st, const IT* ptrs, int64_t n, uint64_t* sizes)                  \
    {                                                                                                           \
        if (n < 0 || (n > 0 && (!ptrs || !sizes))) return GKOB200_EINVAL;                                       \
        return launch_2d(as_stream(st), n, 1, [=] __device__(int64_t i, int64_t) {                              \
            sizes[i] = static_cast<uint64_t>(ptrs[i + 1] - ptrs[i]);                                            \
        });                                                                                                     \
    }                                                                                                           \
    int gkob200_compute_max_row_nnz_##I(void* st, const IT* row_ptrs, int64_t n, uint64_t* max_nnz)              \
    {                                                                                                           \
        if (n < 0 || !max_nnz) return GKOB200_EINVAL;                                                           \
        GKOB200_CUDA(cudaMemsetAsync(max_nnz, 0, sizeof(uint64_t), as_stream(st)));                             \
        if (n == 0) return 0;                                                                                   \
        max_row_nnz_kernel<IT><<<grid_for(n, 256, 8), 256, 0, as_stream(st)>>>(                                 \
            row_ptrs, n, reinterpret_cast<unsigned long long*>(max_nnz));                                       \
        GKOB200_CHECK_LAUNCH();                                                                                 \
        return 0;                                                                                               \
    }                                                                                                           \
    int gkob200_sellp_compute_slice_sets_##I(void* st, const IT* row_ptrs, int64_t n, int64_t slice_size,        \
                                             int64_t stride_factor, uint64_t* slice_sets,                       \
                                             uint64_t* slice_lengths, void* ws, size_t wsb)                     \
    {                                                                                                           \
        if (n < 0 || slice_size <= 0 || stride_factor <= 0 || !slice_sets) return GKOB200_EINVAL;               \
        const int64_t ns = ceildiv(n, slice_size);                                                              \
        if (ns > 0) {                                                                                           \
            slice_lengths_kernel<IT><<<static_cast<unsigned>(ceildiv(ns * 32, 256)), 256, 0, as_stream(st)>>>(  \
                row_ptrs, n, slice_size, stride_factor, ns, slice_sets, slice_lengths);                         \
            GKOB200_CHECK_LAUNCH();                                                                             \
        }                                                                                                       \
        return prefix_sum_impl<uint64_t>(as_stream(st), slice_sets, ns + 1, ws, wsb);                           \
    }                                                                                                           \
    int gkob200_row_len_histogram_##I(void* st, const IT* row_ptrs, int64_t n, uint64_t lo, uint64_t width,      \
                                      int bins, uint64_t* hist)                                                 \
    {                                                                                                           \
        if (n < 0 || bins <= 0 || bins > 8192 || width == 0 || !hist) return GKOB200_EINVAL;                    \
        GKOB200_CUDA(cudaMemsetAsync(hist, 0, sizeof(uint64_t) * bins, as_stream(st)));                         \
        if (n == 0) return 0;                                                                                   \
        row_len_histogram<IT><<<grid_for(n, 256, 4), 256, bins * sizeof(unsigned), as_stream(st)>>>(            \
            row_ptrs, n, lo, width, bins, reinterpret_cast<unsigned long long*>(hist));                         \
        GKOB200_CHECK_LAUNCH();                                                                                 \
        return 0;                                                                                               \
    }
GKOB200_DEF_CONV_I(i32, int32_t)
GKOB200_DEF_CONV_I(i64, int64_t)

/* coo_row_ptrs[i] = max(0, row_nnz[i] - ell_lim), then exclusive scan (n+1 entries) */
int gkob200_hybrid_compute_coo_row_ptrs(void* st, const uint64_t* row_nnz, int64_t n, uint64_t ell_lim,
                                        int64_t* coo_row_ptrs, void* ws, size_t wsb)
{
    if (n < 0 || !coo_row_ptrs) return GKOB200_EINVAL;
    int rc = launch_2d(as_stream(st), n, 1, [=] __device__(int64_t i, int64_t) {
        const int64_t d = static_cast<int64_t>(row_nnz[i]) - static_cast<int64_t>(ell_lim);
        coo_row_ptrs[i] = d > 0 ? d : 0;
    });
    if (rc) return rc;
    return prefix_sum_impl<int64_t>(as_stream(st), coo_row_ptrs, n + 1, ws, wsb);
}

#define GKOB200_DEF_CONV_VI(V, VT, I, IT)                                                                       \
    int gkob200_csr_convert_to_ell_##V##_##I(void* st, int64_t n, const IT* row_ptrs, const IT* cols,            \
                                             const VT* vals, int64_t ell_width, int64_t ell_stride,             \
                                             IT* ell_cols, VT* ell_vals)                                        \
    {                                                                                                           \
        if (n < 0 || ell_width < 0 || ell_stride < n) return GKOB200_EINVAL;                                    \
        return launch_2d(as_stream(st), n, 1, [=] __device__(int64_t row, int64_t) {                            \
            const int64_t rb = row_ptrs[row], re = row_ptrs[row + 1];                                           \
            int64_t out = row;                                                                                  \
            for (int64_t i = rb; i < rb + ell_width; ++i) {                                                     \
                ell_cols[out] = i < re ? cols[i] : IT(-1);                                                      \
                ell_vals[out] = i < re ? vals[i] : VT(0);                                                       \
                out += ell_stride;                                                                              \
            }                                                                                                   \
        });                                                                                                     \
    }                                                                                                           \
    int gkob200_csr_convert_to_sellp_##V##_##I(void* st, int64_t n, const IT* row_ptrs, const IT* cols,          \
                                               const VT* vals, int64_t slice_size, const uint64_t* slice_sets,  \
                                               IT* out_cols, VT* out_vals)                                      \
    {                                                                                                           \
        if (n < 0 || slice_size <= 0) return GKOB200_EINVAL;                                                    \
        return launch_2d(as_stream(st), n, 1, [=] __device__(int64_t row, int64_t) {                            \
            const int64_t rb = row_ptrs[row], re = row_ptrs[row + 1];                                           \
            const int64_t slice = row / slice_size, local = row % slice_size;                                   \
            const int64_t sb = static_cast<int64_t>(slice_sets[slice]);                                         \
            const int64_t sl = static_cast<int64_t>(slice_sets[slice + 1]) - sb;                                \
            int64_t out = sb * slice_size + local;                                                              \
            for (int64_t i = rb; i < rb + sl; ++i) {                                                            \
                out_cols[out] = i < re ? cols[i] : IT(-1);                                                      \
                out_vals[out] = i < re ? vals[i] : VT(0);                                                       \
                out += slice_size;                                                                              \
            }                                                                                                   \
        });                                                                                                     \
    }                                                                                                           \
    int gkob200_csr_convert_to_hybrid_##V##_##I(void* st, int64_t n, const IT* row_ptrs, const IT* cols,         \
                                                const VT* vals, const int64_t* coo_row_ptrs,                    \
                                                int64_t ell_stride, int64_t ell_width, IT* ell_cols,            \
                                                VT* ell_vals, IT* coo_rows, IT* coo_cols, VT* coo_vals)         \
    {                                                                                                           \
        if (n < 0 || ell_width < 0 || ell_stride < n) return GKOB200_EINVAL;                                    \
        return launch_2d(as_stream(st), n, 1, [=] __device__(int64_t row, int64_t) {                            \
            const int64_t rb = row_ptrs[row];                                                                   \
            const int64_t size = row_ptrs[row + 1] - rb;                                                        \
            for (int64_t i = 0; i < ell_width; ++i) {                                                           \
                const bool use = i < size;                                                                      \
                ell_cols[row + ell_stride * i] = use ? cols[rb + i] : IT(-1);                                   \
                ell_vals[row + ell_stride * i] = use ? vals[rb + i] : VT(0);                                    \
            }                                                                                                   \
            const int64_t cb = coo_row_ptrs[row];                                                               \
            for (int64_t i = ell_width; i < size; ++i) {                                                        \
                const int64_t o = cb + i - ell_width;                                                           \
                coo_rows[o] = static_cast<IT>(row);                                                             \
                coo_cols[o] = cols[rb + i];                                                                     \
                coo_vals[o] = vals[rb + i];                                                                     \
            }                                                                                                   \
        });                                                                                                     \
    }
GKOB200_DEF_CONV_VI(f64, double, i32, int32_t)
GKOB200_DEF_CONV_VI(f32, float, i32, int32_t)
GKOB200_DEF_CONV_VI(f64, double, i64, int64_t)
GKOB200_DEF_CONV_VI(f32, float, i64, int64_t)

}  // extern "C"
