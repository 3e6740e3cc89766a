// distributed_kernels.cu — integer bookkeeping of the row-partitioned matrix:
// partition ranges, the split of a global-index matrix into the local block and the
// non-local (ghost-column) block, and the halo index maps.  Bit-exact with the reference.
//
// [ref] reference/distributed/partition_kernels.cpp:42-160,
//       reference/distributed/matrix_kernels.cpp:49-236 (build_local_nonlocal; replaced
//       common/cuda_hip/distributed/matrix_kernels.hpp.inc:55-288, a thrust pipeline),
//       reference/distributed/vector_kernels.cpp:45-95 (build_local).
// The one sort of the path (ghost columns by (owner part, global index)) uses
// cub::DeviceRadixSort from the CUDA toolkit — setup-time only, like the reference's thrust
// call; everything else is hand-written.
#include <cub/device/device_radix_sort.cuh>

#include "launch.cuh"

namespace gkob200 {
namespace {

struct PartitionView {
    int64_t num_ranges;
    const int64_t* range_bounds;        // [num_ranges + 1]
    const int32_t* part_ids;            // [num_ranges]
    const int32_t* range_starting_idx;  // [num_ranges]
    // upper_bound(range_bounds + 1, range_bounds + num_ranges + 1, idx) - (range_bounds + 1)
    __device__ __forceinline__ int64_t find_range(int64_t idx) const
    {
        int64_t lo = 0, hi = num_ranges;
        while (lo < hi) {
            const int64_t mid = (lo + hi) >> 1;
            if (range_bounds[mid + 1] <= idx) lo = mid + 1; else hi = mid;
        }
        return lo;
    }
    __device__ __forceinline__ int32_t to_local(int64_t idx, int64_t range) const
    {
        return static_cast<int32_t>(idx - range_bounds[range]) + range_starting_idx[range];
    }
};

// flags: 1 = local entry, 2 = non-local entry (row owned, column not), 0 = not ours
__global__ void classify_entries(int64_t nnz, const int64_t* __restrict__ rows, const int64_t* __restrict__ cols,
                                 PartitionView rp, PartitionView cp, int32_t local_part, int32_t* __restrict__ f_loc,
                                 int32_t* __restrict__ f_nl)
{
    const int64_t i = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x;
    if (i > nnz) return;
    if (i == nnz) {
        f_loc[i] = 0;
        f_nl[i] = 0;
        return;
    }
    const int64_t rr = rp.find_range(rows[i]);
    int loc = 0, nl = 0;
    if (rp.part_ids[rr] == local_part) {
        const int64_t cr = cp.find_range(cols[i]);
        if (cp.part_ids[cr] == local_part) loc = 1; else nl = 1;
    }
    f_loc[i] = loc;
    f_nl[i] = nl;
}

// stable compaction using the exclusive scans of the flags
template <typename V>
__global__ void scatter_entries(int64_t nnz, const int64_t* __restrict__ rows, const int64_t* __restrict__ cols,
                                const V* __restrict__ vals, PartitionView rp, PartitionView cp,
                                const int32_t* __restrict__ s_loc, const int32_t* __restrict__ s_nl,
                                int32_t* __restrict__ lrow, int32_t* __restrict__ lcol, V* __restrict__ lval,
                                int32_t* __restrict__ nrow, int64_t* __restrict__ ncol_global, V* __restrict__ nval,
                                uint64_t* __restrict__ keys, int64_t key_mult)
{
    const int64_t i = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x;
    if (i >= nnz) return;
    const bool loc = s_loc[i + 1] != s_loc[i], nl = s_nl[i + 1] != s_nl[i];
    if (!loc && !nl) return;
    const int64_t rr = rp.find_range(rows[i]);
    const int32_t r = rp.to_local(rows[i], rr);
    if (loc) {
        const int64_t cr = cp.find_range(cols[i]);
        const int32_t o = s_loc[i];
        lrow[o] = r;
        lcol[o] = cp.to_local(cols[i], cr);
        lval[o] = vals[i];
    } else {
        const int64_t cr = cp.find_range(cols[i]);
        const int32_t o = s_nl[i];
        nrow[o] = r;
        ncol_global[o] = cols[i];
        nval[o] = vals[i];
        // sort key: (owner part, global column)
        keys[o] = static_cast<uint64_t>(cp.part_ids[cr]) * static_cast<uint64_t>(key_mult) + static_cast<uint64_t>(cols[i]);
    }
}

__global__ void unique_flags(int64_t n, const uint64_t* __restrict__ sorted, int32_t* __restrict__ f)
{
    const int64_t i = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x;
    if (i > n) return;
    f[i] = (i < n && (i == 0 || sorted[i] != sorted[i - 1])) ? 1 : 0;
}

__global__ void compact_unique(int64_t n, const uint64_t* __restrict__ sorted, const int32_t* __restrict__ s,
                               uint64_t* __restrict__ uniq, int64_t key_mult, PartitionView cp,
                               int64_t* __restrict__ nl_to_global, int32_t* __restrict__ gather_idxs)
{
    const int64_t i = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x;
    if (i >= n) return;
    if (s[i + 1] == s[i]) return;
    const int32_t o = s[i];
    const uint64_t key = sorted[i];
    uniq[o] = key;
    const int64_t col = static_cast<int64_t>(key % static_cast<uint64_t>(key_mult));
    nl_to_global[o] = col;
    gather_idxs[o] = cp.to_local(col, cp.find_range(col));
}

// non-local column index = rank of the entry's key among the unique keys
__global__ void renumber_columns(int64_t n_nl, const uint64_t* __restrict__ keys, int64_t n_uniq,
                                 const uint64_t* __restrict__ uniq, int32_t* __restrict__ ncol)
{
    const int64_t i = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x;
    if (i >= n_nl) return;
    const uint64_t key = keys[i];
    int64_t lo = 0, hi = n_uniq;
    while (lo < hi) {
        const int64_t mid = (lo + hi) >> 1;
        if (uniq[mid] < key) lo = mid + 1; else hi = mid;
    }
    ncol[i] = static_cast<int32_t>(lo);
}

// recv_sizes[p] = number of unique keys owned by part p (keys are sorted by part)
__global__ void count_recv_sizes(int32_t num_parts, int64_t n_uniq, const uint64_t* __restrict__ uniq, int64_t key_mult,
                                 int32_t* __restrict__ recv_sizes)
{
    const int32_t p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= num_parts) return;
    auto lower = [&](uint64_t key) {
        int64_t lo = 0, hi = n_uniq;
        while (lo < hi) {
            const int64_t mid = (lo + hi) >> 1;
            if (uniq[mid] < key) lo = mid + 1; else hi = mid;
        }
        return lo;
    };
    recv_sizes[p] = static_cast<int32_t>(lower(static_cast<uint64_t>(p + 1) * key_mult) - lower(static_cast<uint64_t>(p) * key_mult));
}

struct TmpBuf {
    void* p = nullptr;
    ~TmpBuf() { if (p) cudaFree(p); }
    int alloc(size_t bytes) { return static_cast<int>(cudaMalloc(&p, bytes ? bytes : 8)); }
};

template <typename V>
int build_local_nonlocal_impl(cudaStream_t s, int64_t nnz, const int64_t* rows, const int64_t* cols, const V* vals,
                              PartitionView rp, PartitionView cp, int64_t global_cols, int32_t num_parts,
                              int32_t local_part, int32_t* lrow, int32_t* lcol, V* lval, int32_t* nrow, int32_t* ncol,
                              V* nval, int32_t* gather_idxs, int32_t* recv_sizes, int64_t* nl_to_global,
                              int64_t* counts_host)
{
    if (nnz < 0 || num_parts <= 0 || !counts_host || !recv_sizes) return GKOB200_EINVAL;
    counts_host[0] = counts_host[1] = counts_host[2] = 0;
    GKOB200_CUDA(cudaMemsetAsync(recv_sizes, 0, sizeof(int32_t) * num_parts, s));
    if (nnz == 0) {
        GKOB200_CUDA(cudaStreamSynchronize(s));
        return 0;
    }
    if (nnz >= (int64_t(1) << 31) - 2) return GKOB200_EUNSUPPORTED;  // LocalIndexType is int32
    const int64_t key_mult = global_cols > 0 ? global_cols : 1;
    TmpBuf f_loc, f_nl, scanws, keys, keys_sorted, ncol_g, uniq, cubtmp;
    int rc;
    if ((rc = f_loc.alloc((nnz + 1) * sizeof(int32_t)))) return rc;
    if ((rc = f_nl.alloc((nnz + 1) * sizeof(int32_t)))) return rc;
    const size_t scan_bytes = gkob200_prefix_sum_workspace_bytes(nnz + 1);
    if ((rc = scanws.alloc(scan_bytes))) return rc;
    int32_t* sl = static_cast<int32_t*>(f_loc.p);
    int32_t* sn = static_cast<int32_t*>(f_nl.p);
    const unsigned grid = static_cast<unsigned>(ceildiv(nnz + 1, 256));
    classify_entries<<<grid, 256, 0, s>>>(nnz, rows, cols, rp, cp, local_part, sl, sn);
    if ((rc = gkob200_prefix_sum_i32(s, sl, nnz + 1, scanws.p, scan_bytes))) return rc;
    if ((rc = gkob200_prefix_sum_i32(s, sn, nnz + 1, scanws.p, scan_bytes))) return rc;
    int32_t n_loc = 0, n_nl = 0;
    GKOB200_CUDA(cudaMemcpyAsync(&n_loc, sl + nnz, sizeof(int32_t), cudaMemcpyDeviceToHost, s));
    GKOB200_CUDA(cudaMemcpyAsync(&n_nl, sn + nnz, sizeof(int32_t), cudaMemcpyDeviceToHost, s));
    GKOB200_CUDA(cudaStreamSynchronize(s));
    if ((rc = keys.alloc(static_cast<size_t>(n_nl) * sizeof(uint64_t)))) return rc;
    if ((rc = keys_sorted.alloc(static_cast<size_t>(n_nl) * sizeof(uint64_t)))) return rc;
    if ((rc = ncol_g.alloc(static_cast<size_t>(n_nl) * sizeof(int64_t)))) return rc;
    scatter_entries<V><<<grid, 256, 0, s>>>(nnz, rows, cols, vals, rp, cp, sl, sn, lrow, lcol, lval, nrow,
                                            static_cast<int64_t*>(ncol_g.p), nval, static_cast<uint64_t*>(keys.p),
                                            key_mult);
    GKOB200_CHECK_LAUNCH();
    int64_t n_uniq = 0;
    if (n_nl > 0) {
        size_t tmp_bytes = 0;
        uint64_t* k_in = static_cast<uint64_t*>(keys.p);
        uint64_t* k_out = static_cast<uint64_t*>(keys_sorted.p);
        GKOB200_CUDA(cub::DeviceRadixSort::SortKeys(nullptr, tmp_bytes, k_in, k_out, n_nl, 0, 64, s));
        if ((rc = cubtmp.alloc(tmp_bytes))) return rc;
        GKOB200_CUDA(cub::DeviceRadixSort::SortKeys(cubtmp.p, tmp_bytes, k_in, k_out, n_nl, 0, 64, s));
        // unique
        const unsigned g2 = static_cast<unsigned>(ceildiv(static_cast<int64_t>(n_nl) + 1, 256));
        unique_flags<<<g2, 256, 0, s>>>(n_nl, k_out, sl);  // reuse sl as flag/scan array (n_nl <= nnz)
        if ((rc = gkob200_prefix_sum_i32(s, sl, static_cast<int64_t>(n_nl) + 1, scanws.p, scan_bytes))) return rc;
        int32_t nu = 0;
        GKOB200_CUDA(cudaMemcpyAsync(&nu, sl + n_nl, sizeof(int32_t), cudaMemcpyDeviceToHost, s));
        GKOB200_CUDA(cudaStreamSynchronize(s));
        n_uniq = nu;
        if ((rc = uniq.alloc(static_cast<size_t>(n_uniq) * sizeof(uint64_t)))) return rc;
        compact_unique<<<g2, 256, 0, s>>>(n_nl, k_out, sl, static_cast<uint64_t*>(uniq.p), key_mult, cp, nl_to_global,
                                          gather_idxs);
        renumber_columns<<<g2, 256, 0, s>>>(n_nl, k_in, n_uniq, static_cast<uint64_t*>(uniq.p), ncol);
        count_recv_sizes<<<static_cast<unsigned>(ceildiv(num_parts, 128)), 128, 0, s>>>(
            num_parts, n_uniq, static_cast<uint64_t*>(uniq.p), key_mult, recv_sizes);
        GKOB200_CHECK_LAUNCH();
    }
    GKOB200_CUDA(cudaStreamSynchronize(s));
    counts_host[0] = n_loc;
    counts_host[1] = n_nl;
    counts_host[2] = n_uniq;
    return 0;
}

}  // namespace
}  // namespace gkob200

using namespace gkob200;

extern "C" {

/* ranges[i] = ranges[i-1] + size/P + ((i-1) < size % P)   [partition_kernels.cpp:95-110] */
int gkob200_partition_build_ranges_from_global_size_i64(void* stream, int32_t num_parts, int64_t global_size,
                                                        int64_t* ranges)
{
    if (num_parts <= 0 || global_size < 0 || !ranges) return GKOB200_EINVAL;
    return launch_2d(as_stream(stream), static_cast<int64_t>(num_parts) + 1, 1, [=] __device__(int64_t i, int64_t) {
        const int64_t per = global_size / num_parts, rest = global_size - num_parts * per;
        ranges[i] = i * per + (i < rest ? i : rest);
    });
}

/* range_bounds = ranges, part_ids[i] = i   [partition_kernels.cpp:55-67] */
int gkob200_partition_build_from_contiguous_i64(void* stream, int32_t num_parts, const int64_t* ranges,
                                                int64_t* range_bounds, int32_t* part_ids)
{
    if (num_parts < 0 || !ranges || !range_bounds) return GKOB200_EINVAL;
    return launch_2d(as_stream(stream), static_cast<int64_t>(num_parts) + 1, 1, [=] __device__(int64_t i, int64_t) {
        range_bounds[i] = i == 0 ? 0 : ranges[i];
        if (i < num_parts) part_ids[i] = static_cast<int32_t>(i);
    });
}

/* [partition_kernels.cpp:70-90] one range per change of part id in `mapping` */
int gkob200_partition_build_from_mapping_i64(void* stream, int64_t n, const int32_t* mapping, int64_t* range_bounds,
                                             int32_t* part_ids, int64_t* num_ranges_dev, void* ws, size_t ws_bytes)
{
    if (n < 0 || !range_bounds || !num_ranges_dev) return GKOB200_EINVAL;
    const size_t need = static_cast<size_t>(n + 2) * sizeof(int32_t) + gkob200_prefix_sum_workspace_bytes(n + 1);
    if (!ws || ws_bytes < need) return GKOB200_EWORKSPACE;
    int32_t* f = static_cast<int32_t*>(ws);
    void* sws = f + n + 2;
    int rc = launch_2d(as_stream(stream), n + 1, 1, [=] __device__(int64_t i, int64_t) {
        f[i] = (i < n && (i == 0 || mapping[i] != mapping[i - 1])) ? 1 : 0;
    });
    if (rc) return rc;
    if ((rc = gkob200_prefix_sum_i32(stream, f, n + 1, sws, ws_bytes - static_cast<size_t>(n + 2) * sizeof(int32_t)))) return rc;
    return launch_2d(as_stream(stream), n + 1, 1, [=] __device__(int64_t i, int64_t) {
        if (i == n) {
            range_bounds[f[n]] = n;
            *num_ranges_dev = f[n];
        } else if (f[i + 1] != f[i]) {
            range_bounds[f[i]] = i;
            part_ids[f[i]] = mapping[i];
        }
    });
}

/* [partition_kernels.cpp:113-133] sequential over the (few) ranges */
int gkob200_partition_build_starting_indices_i32_i64(void* stream, const int64_t* range_offsets,
                                                     const int32_t* range_parts, int64_t num_ranges, int32_t num_parts,
                                                     int32_t* num_empty_parts, int32_t* ranks, int32_t* sizes)
{
    if (num_ranges < 0 || num_parts < 0 || !num_empty_parts) return GKOB200_EINVAL;
    return launch_2d(as_stream(stream), 1, 1, [=] __device__(int64_t, int64_t) {
        for (int32_t p = 0; p < num_parts; ++p) sizes[p] = 0;
        for (int64_t r = 0; r < num_ranges; ++r) {
            const int32_t part = range_parts[r];
            ranks[r] = sizes[part];
            sizes[part] += static_cast<int32_t>(range_offsets[r + 1] - range_offsets[r]);
        }
        int32_t e = 0;
        for (int32_t p = 0; p < num_parts; ++p) e += sizes[p] == 0;
        *num_empty_parts = e;
    });
}

#define GKOB200_DEF_DIST(V, VT)                                                                                   \
    int gkob200_dist_build_local_nonlocal_##V(                                                                    \
        void* stream, int64_t nnz, const int64_t* rows, const int64_t* cols, const VT* vals, int64_t row_num_ranges, \
        const int64_t* row_range_bounds, const int32_t* row_part_ids, const int32_t* row_range_starts,             \
        int64_t col_num_ranges, const int64_t* col_range_bounds, const int32_t* col_part_ids,                      \
        const int32_t* col_range_starts, int64_t global_cols, int32_t num_parts, int32_t local_part,               \
        int32_t* local_row_idxs, int32_t* local_col_idxs, VT* local_values, int32_t* non_local_row_idxs,           \
        int32_t* non_local_col_idxs, VT* non_local_values, int32_t* local_gather_idxs, int32_t* recv_sizes,        \
        int64_t* non_local_to_global, int64_t* counts_host)                                                       \
    {                                                                                                             \
        PartitionView rp{row_num_ranges, row_range_bounds, row_part_ids, row_range_starts};                        \
        PartitionView cp{col_num_ranges, col_range_bounds, col_part_ids, col_range_starts};                        \
        return build_local_nonlocal_impl<VT>(as_stream(stream), nnz, rows, cols, vals, rp, cp, global_cols,        \
                                             num_parts, local_part, local_row_idxs, local_col_idxs, local_values,  \
                                             non_local_row_idxs, non_local_col_idxs, non_local_values,             \
                                             local_gather_idxs, recv_sizes, non_local_to_global, counts_host);     \
    }                                                                                                             \
    /* distributed_vector::build_local [reference/distributed/vector_kernels.cpp:45-95] */                        \
    int gkob200_dist_vector_build_local_##V(void* stream, int64_t nnz, const int64_t* rows, const int64_t* cols,   \
                                            const VT* vals, int64_t num_ranges, const int64_t* range_bounds,       \
                                            const int32_t* part_ids, const int32_t* range_starts,                  \
                                            int32_t local_part, VT* local, int64_t local_stride)                   \
    {                                                                                                             \
        if (nnz < 0) return GKOB200_EINVAL;                                                                       \
        PartitionView p{num_ranges, range_bounds, part_ids, range_starts};                                         \
        return launch_2d(as_stream(stream), nnz, 1, [=] __device__(int64_t i, int64_t) {                           \
            const int64_t r = p.find_range(rows[i]);                                                              \
            if (p.part_ids[r] == local_part) local[p.to_local(rows[i], r) * local_stride + cols[i]] = vals[i];     \
        });                                                                                                       \
    }
GKOB200_DEF_DIST(f64, double)
GKOB200_DEF_DIST(f32, float)

}  // extern "C"
