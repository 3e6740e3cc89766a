// setup_kernels.cu — matrix assembly on the device (the steps every run
// executes before the hot loop).  Replaces
//   components::sort_row_major / sum_duplicates / remove_zeros
//       (reference/base/device_matrix_data_kernels.cpp:82-172; cuda: thrust sort / reduce_by_key)
//   csr::sort_by_column_index, csr::transpose
//       (reference/matrix/csr_kernels.cpp:969-987, :551-587; cuda: cusparse csrsort / csr2csc)
// All of them are one stable radix sort of 64-bit keys (cub::DeviceRadixSort, CUDA toolkit —
// setup-time only) plus gathers, an exclusive scan and per-entry kernels:
//   * sort_row_major: key = row * n_cols + col.  The reference's std::sort leaves the order of
//     duplicates unspecified; the stable sort keeps their input order, which is one of the orders
//     the reference may produce.  Rows / columns are decoded from the sorted keys.
//   * sum_duplicates: head flags -> exclusive scan -> one thread per unique entry adds its
//     duplicates one after the other starting from zero: the same bits as the reference loop.
//   * remove_zeros: stable compaction (flags -> scan -> scatter).
//   * transpose: stable sort of the row-major entries by column = CSR of A^T with rows ascending
//     inside every column, exactly the reference's counting sort.
//   * sort_by_column_index: key = row * n_cols + col over the expanded row indices.
// Integer outputs are bit-exact, values bit-exact (sums in the reference's order).
#include <cub/cub.cuh>

#include "common.cuh"

using namespace gkob200;

extern "C" {
size_t gkob200_prefix_sum_workspace_bytes(int64_t n);
int gkob200_prefix_sum_i64(void* stream, int64_t* data, int64_t n, void* ws, size_t ws_bytes);
int gkob200_convert_ptrs_to_idxs_i32(void* stream, const int32_t* ptrs, int64_t n, int32_t* idxs);
int gkob200_convert_ptrs_to_idxs_i64(void* stream, const int64_t* ptrs, int64_t n, int64_t* idxs);
int gkob200_convert_idxs_to_ptrs_i32(void* stream, const int32_t* idxs, int64_t num_idxs, int64_t n, int32_t* ptrs);
int gkob200_convert_idxs_to_ptrs_i64(void* stream, const int64_t* idxs, int64_t num_idxs, int64_t n, int64_t* ptrs);
}

namespace gkob200 {
namespace {

inline size_t up256(size_t b) { return (b + 255) / 256 * 256; }

// row * n_cols + col must fit 64 bits
inline bool key_overflows(int64_t n_rows, int64_t n_cols)
{
    return n_cols > 0 && static_cast<uint64_t>(n_rows) > ~uint64_t(0) / static_cast<uint64_t>(n_cols);
}

inline int bits_for(uint64_t max_value)
{
    int b = 1;
    while (b < 64 && (max_value >> b) != 0) ++b;
    return b;
}

// carve consecutive 256-byte aligned arrays out of a workspace
struct Carver {
    unsigned char* p;
    size_t left;
    template <typename T>
    T* take(size_t count)
    {
        const size_t b = up256(count * sizeof(T));
        if (b > left) return nullptr;
        T* r = reinterpret_cast<T*>(p);
        p += b;
        left -= b;
        return r;
    }
};

size_t cub_sort_bytes(int64_t n)
{
    size_t bytes = 0;
    cub::DeviceRadixSort::SortPairs(nullptr, bytes, static_cast<const uint64_t*>(nullptr),
                                    static_cast<uint64_t*>(nullptr), static_cast<const int64_t*>(nullptr),
                                    static_cast<int64_t*>(nullptr), n, 0, 64, static_cast<cudaStream_t>(nullptr));
    return bytes + 256;
}

// everything the sort-based kernels need: keys in/out, permutation in/out, one index and one
// value array of scratch, the cub temporary
size_t sort_ws_bytes(int64_t nnz, size_t vb, size_t ib)
{
    const size_t n = static_cast<size_t>(nnz > 0 ? nnz : 1);
    return 4 * up256(n * 8) + up256(n * ib) + up256(n * vb) + up256(cub_sort_bytes(nnz)) + 1024;
}

template <typename I>
__global__ void __launch_bounds__(256)
    make_keys(int64_t nnz, const I* __restrict__ rows, const I* __restrict__ cols, uint64_t n_cols, bool col_only,
              uint64_t* __restrict__ keys, int64_t* __restrict__ perm)
{
    const int64_t i = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x;
    if (i >= nnz) return;
    keys[i] = col_only ? static_cast<uint64_t>(cols[i])
                       : static_cast<uint64_t>(rows[i]) * n_cols + static_cast<uint64_t>(cols[i]);
    perm[i] = i;
}

template <typename I>
__global__ void __launch_bounds__(256)
    decode_keys(int64_t nnz, const uint64_t* __restrict__ keys, uint64_t n_cols, I* __restrict__ rows, I* __restrict__ cols)
{
    const int64_t i = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x;
    if (i >= nnz) return;
    const uint64_t k = keys[i];
    if (rows) rows[i] = static_cast<I>(k / n_cols);
    cols[i] = static_cast<I>(k % n_cols);
}

template <typename T>
__global__ void __launch_bounds__(256) gather(int64_t n, const int64_t* __restrict__ perm, const T* __restrict__ in, T* __restrict__ out)
{
    const int64_t i = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x;
    if (i < n) out[i] = in[perm[i]];
}

inline unsigned blocks(int64_t n) { return static_cast<unsigned>((n + 255) / 256); }

struct SortBuffers {
    uint64_t *keys_in, *keys_out;
    int64_t *perm_in, *perm_out;
    void* idx_tmp;
    void* val_tmp;
    void* cub_tmp;
    size_t cub_bytes;
};

int carve_sort(void* ws, size_t ws_bytes, int64_t nnz, size_t vb, size_t ib, SortBuffers& b)
{
    if (!ws || ws_bytes < sort_ws_bytes(nnz, vb, ib)) return GKOB200_EWORKSPACE;
    Carver c{static_cast<unsigned char*>(ws), ws_bytes};
    const size_t n = static_cast<size_t>(nnz > 0 ? nnz : 1);
    b.keys_in = c.take<uint64_t>(n);
    b.keys_out = c.take<uint64_t>(n);
    b.perm_in = c.take<int64_t>(n);
    b.perm_out = c.take<int64_t>(n);
    b.idx_tmp = c.take<unsigned char>(n * ib);
    b.val_tmp = c.take<unsigned char>(n * vb);
    b.cub_bytes = cub_sort_bytes(nnz);
    b.cub_tmp = c.take<unsigned char>(b.cub_bytes);
    return b.cub_tmp ? 0 : GKOB200_EWORKSPACE;
}

template <typename V, typename I>
int sort_row_major(cudaStream_t s, int64_t n_rows, int64_t n_cols, int64_t nnz, I* rows, I* cols, V* vals, void* ws,
                   size_t ws_bytes)
{
    if (n_rows < 0 || n_cols < 0 || nnz < 0) return GKOB200_EINVAL;
    if (nnz == 0) return 0;
    if (!rows || !cols || !vals || n_cols == 0) return GKOB200_EINVAL;
    if (key_overflows(n_rows, n_cols)) return GKOB200_EUNSUPPORTED;
    SortBuffers b;
    int rc = carve_sort(ws, ws_bytes, nnz, sizeof(V), sizeof(I), b);
    if (rc) return rc;
    make_keys<I><<<blocks(nnz), 256, 0, s>>>(nnz, rows, cols, static_cast<uint64_t>(n_cols), false, b.keys_in, b.perm_in);
    GKOB200_CHECK_LAUNCH();
    const int end_bit = bits_for(static_cast<uint64_t>(n_rows) * static_cast<uint64_t>(n_cols));
    GKOB200_CUDA(cub::DeviceRadixSort::SortPairs(b.cub_tmp, b.cub_bytes, b.keys_in, b.keys_out, b.perm_in, b.perm_out, nnz,
                                                 0, end_bit, s));
    decode_keys<I><<<blocks(nnz), 256, 0, s>>>(nnz, b.keys_out, static_cast<uint64_t>(n_cols), rows, cols);
    GKOB200_CHECK_LAUNCH();
    V* tmp = static_cast<V*>(b.val_tmp);
    gather<V><<<blocks(nnz), 256, 0, s>>>(nnz, b.perm_out, vals, tmp);
    GKOB200_CHECK_LAUNCH();
    GKOB200_CUDA(cudaMemcpyAsync(vals, tmp, static_cast<size_t>(nnz) * sizeof(V), cudaMemcpyDeviceToDevice, s));
    return 0;
}

// ---- compaction-type kernels: flags -> exclusive scan -> scatter -----------------------------
size_t scan_ws_bytes(int64_t nnz)
{
    const size_t n = static_cast<size_t>(nnz) + 2;
    return 2 * up256(n * 8) + up256(gkob200_prefix_sum_workspace_bytes(nnz + 1)) + 1024;
}

template <typename V, typename I, bool Duplicates>
__global__ void __launch_bounds__(256)
    mark(int64_t nnz, const I* __restrict__ rows, const I* __restrict__ cols, const V* __restrict__ vals, int64_t* __restrict__ pos)
{
    const int64_t i = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x;
    if (i > nnz) return;
    if (i == nnz) {
        pos[i] = 0;   // receives the total
        return;
    }
    if (Duplicates)
        pos[i] = (i == 0 || rows[i] != rows[i - 1] || cols[i] != cols[i - 1]) ? 1 : 0;
    else
        pos[i] = vals[i] != V(0) ? 1 : 0;
}

// Duplicates: entry i is a head iff pos[i+1] != pos[i]; unique entry u = pos[i] starts at i
template <typename V, typename I>
__global__ void __launch_bounds__(256)
    scatter_heads(int64_t nnz, const I* __restrict__ rows, const I* __restrict__ cols, const int64_t* __restrict__ pos,
                  I* __restrict__ out_rows, I* __restrict__ out_cols, int64_t* __restrict__ starts, int64_t* __restrict__ out_nnz)
{
    const int64_t i = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x;
    if (i > nnz) return;
    if (i == nnz) {
        starts[pos[nnz]] = nnz;
        *out_nnz = pos[nnz];
        return;
    }
    if (pos[i + 1] != pos[i]) {
        const int64_t u = pos[i];
        out_rows[u] = rows[i];
        out_cols[u] = cols[i];
        starts[u] = i;
    }
}

template <typename V>
__global__ void __launch_bounds__(256)
    sum_runs(int64_t nnz, const int64_t* __restrict__ pos, const int64_t* __restrict__ starts, const V* __restrict__ vals,
             V* __restrict__ out_vals)
{
    const int64_t u = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x;
    if (u >= pos[nnz]) return;
    V acc = V(0);   // the reference starts every unique entry at zero and adds in order
    for (int64_t k = starts[u]; k < starts[u + 1]; ++k) acc = add_rn(acc, vals[k]);
    out_vals[u] = acc;
}

template <typename V, typename I>
__global__ void __launch_bounds__(256)
    scatter_nonzeros(int64_t nnz, const I* __restrict__ rows, const I* __restrict__ cols, const V* __restrict__ vals,
                     const int64_t* __restrict__ pos, I* __restrict__ out_rows, I* __restrict__ out_cols,
                     V* __restrict__ out_vals, int64_t* __restrict__ out_nnz)
{
    const int64_t i = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x;
    if (i > nnz) return;
    if (i == nnz) {
        *out_nnz = pos[nnz];
        return;
    }
    if (pos[i + 1] != pos[i]) {
        const int64_t u = pos[i];
        out_rows[u] = rows[i];
        out_cols[u] = cols[i];
        out_vals[u] = vals[i];
    }
}

template <typename V, typename I, bool Duplicates>
int compact(cudaStream_t s, int64_t nnz, const I* rows, const I* cols, const V* vals, I* out_rows, I* out_cols,
            V* out_vals, int64_t* out_nnz, void* ws, size_t ws_bytes)
{
    if (nnz < 0 || !out_nnz) return GKOB200_EINVAL;
    if (nnz == 0) {
        GKOB200_CUDA(cudaMemsetAsync(out_nnz, 0, sizeof(int64_t), s));
        return 0;
    }
    if (!rows || !cols || !vals || !out_rows || !out_cols || !out_vals) return GKOB200_EINVAL;
    if (!ws || ws_bytes < scan_ws_bytes(nnz)) return GKOB200_EWORKSPACE;
    Carver c{static_cast<unsigned char*>(ws), ws_bytes};
    int64_t* pos = c.take<int64_t>(static_cast<size_t>(nnz) + 2);
    int64_t* starts = c.take<int64_t>(static_cast<size_t>(nnz) + 2);
    const size_t scan_bytes = gkob200_prefix_sum_workspace_bytes(nnz + 1);
    void* scan_ws = c.take<unsigned char>(scan_bytes);
    if (!scan_ws) return GKOB200_EWORKSPACE;
    mark<V, I, Duplicates><<<blocks(nnz + 1), 256, 0, s>>>(nnz, rows, cols, vals, pos);
    GKOB200_CHECK_LAUNCH();
    int rc = gkob200_prefix_sum_i64(s, pos, nnz + 1, scan_ws, scan_bytes);
    if (rc) return rc;
    if (Duplicates) {
        scatter_heads<V, I><<<blocks(nnz + 1), 256, 0, s>>>(nnz, rows, cols, pos, out_rows, out_cols, starts, out_nnz);
        GKOB200_CHECK_LAUNCH();
        sum_runs<V><<<blocks(nnz), 256, 0, s>>>(nnz, pos, starts, vals, out_vals);
    } else {
        scatter_nonzeros<V, I><<<blocks(nnz + 1), 256, 0, s>>>(nnz, rows, cols, vals, pos, out_rows, out_cols, out_vals,
                                                              out_nnz);
    }
    GKOB200_CHECK_LAUNCH();
    return 0;
}

// ---- CSR transpose / sort_by_column_index -----------------------------------------------------
int ptrs_to_idxs(cudaStream_t s, const int32_t* p, int64_t n, int32_t* idx) { return gkob200_convert_ptrs_to_idxs_i32(s, p, n, idx); }
int ptrs_to_idxs(cudaStream_t s, const int64_t* p, int64_t n, int64_t* idx) { return gkob200_convert_ptrs_to_idxs_i64(s, p, n, idx); }
int idxs_to_ptrs(cudaStream_t s, const int32_t* idx, int64_t m, int64_t n, int32_t* p) { return gkob200_convert_idxs_to_ptrs_i32(s, idx, m, n, p); }
int idxs_to_ptrs(cudaStream_t s, const int64_t* idx, int64_t m, int64_t n, int64_t* p) { return gkob200_convert_idxs_to_ptrs_i64(s, idx, m, n, p); }

template <typename V, typename I>
int csr_transpose(cudaStream_t s, int64_t n_rows, int64_t n_cols, int64_t nnz, const I* rp, const I* ci, const V* va,
                  I* out_rp, I* out_ci, V* out_va, void* ws, size_t ws_bytes)
{
    if (n_rows < 0 || n_cols < 0 || nnz < 0 || !out_rp) return GKOB200_EINVAL;
    if (nnz == 0) {
        GKOB200_CUDA(cudaMemsetAsync(out_rp, 0, static_cast<size_t>(n_cols + 1) * sizeof(I), s));
        return 0;
    }
    if (!rp || !ci || !va || !out_ci || !out_va) return GKOB200_EINVAL;
    SortBuffers b;
    int rc = carve_sort(ws, ws_bytes, nnz, sizeof(V), sizeof(I), b);
    if (rc) return rc;
    I* row_idx = static_cast<I*>(b.idx_tmp);
    if ((rc = ptrs_to_idxs(s, rp, n_rows, row_idx))) return rc;
    make_keys<I><<<blocks(nnz), 256, 0, s>>>(nnz, row_idx, ci, 0, true, b.keys_in, b.perm_in);
    GKOB200_CHECK_LAUNCH();
    GKOB200_CUDA(cub::DeviceRadixSort::SortPairs(b.cub_tmp, b.cub_bytes, b.keys_in, b.keys_out, b.perm_in, b.perm_out, nnz,
                                                 0, bits_for(static_cast<uint64_t>(n_cols)), s));
    // new column indices = old rows, in sorted order; values likewise
    gather<I><<<blocks(nnz), 256, 0, s>>>(nnz, b.perm_out, row_idx, out_ci);
    GKOB200_CHECK_LAUNCH();
    gather<V><<<blocks(nnz), 256, 0, s>>>(nnz, b.perm_out, va, out_va);
    GKOB200_CHECK_LAUNCH();
    // new row pointers from the sorted old columns (decoded into the scratch index array)
    decode_keys<I><<<blocks(nnz), 256, 0, s>>>(nnz, b.keys_out, ~uint64_t(0), static_cast<I*>(nullptr), row_idx);
    GKOB200_CHECK_LAUNCH();
    return idxs_to_ptrs(s, row_idx, nnz, n_cols, out_rp);
}

template <typename V, typename I>
int csr_sort_by_column_index(cudaStream_t s, int64_t n_rows, int64_t n_cols, int64_t nnz, const I* rp, I* ci, V* va,
                             void* ws, size_t ws_bytes)
{
    if (n_rows < 0 || n_cols < 0 || nnz < 0) return GKOB200_EINVAL;
    if (nnz == 0) return 0;
    if (!rp || !ci || !va || n_cols == 0) return GKOB200_EINVAL;
    if (key_overflows(n_rows, n_cols)) return GKOB200_EUNSUPPORTED;
    SortBuffers b;
    int rc = carve_sort(ws, ws_bytes, nnz, sizeof(V), sizeof(I), b);
    if (rc) return rc;
    I* row_idx = static_cast<I*>(b.idx_tmp);
    if ((rc = ptrs_to_idxs(s, rp, n_rows, row_idx))) return rc;
    make_keys<I><<<blocks(nnz), 256, 0, s>>>(nnz, row_idx, ci, static_cast<uint64_t>(n_cols), false, b.keys_in, b.perm_in);
    GKOB200_CHECK_LAUNCH();
    const int end_bit = bits_for(static_cast<uint64_t>(n_rows) * static_cast<uint64_t>(n_cols));
    GKOB200_CUDA(cub::DeviceRadixSort::SortPairs(b.cub_tmp, b.cub_bytes, b.keys_in, b.keys_out, b.perm_in, b.perm_out, nnz,
                                                 0, end_bit, s));
    decode_keys<I><<<blocks(nnz), 256, 0, s>>>(nnz, b.keys_out, static_cast<uint64_t>(n_cols), static_cast<I*>(nullptr), ci);
    GKOB200_CHECK_LAUNCH();
    V* tmp = static_cast<V*>(b.val_tmp);
    gather<V><<<blocks(nnz), 256, 0, s>>>(nnz, b.perm_out, va, tmp);
    GKOB200_CHECK_LAUNCH();
    GKOB200_CUDA(cudaMemcpyAsync(va, tmp, static_cast<size_t>(nnz) * sizeof(V), cudaMemcpyDeviceToDevice, s));
    return 0;
}

// matrix_data_entry<V, I> of the reference: { I row; I column; V value; } with natural alignment
// (include/ginkgo/core/base/matrix_data.hpp:89-118)
template <typename V, typename I>
struct Entry {
    I row;
    I column;
    V value;
};
template <typename V, typename I>
__global__ void __launch_bounds__(256) aos_to_soa_kernel(int64_t nnz, const Entry<V, I>* __restrict__ in, I* rows,
                                                         I* cols, V* vals)
{
    for (int64_t i = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; i < nnz;
         i += static_cast<int64_t>(gridDim.x) * blockDim.x) {
        const Entry<V, I> e = in[i];
        rows[i] = e.row;
        cols[i] = e.column;
        vals[i] = e.value;
    }
}
template <typename V, typename I>
__global__ void __launch_bounds__(256) soa_to_aos_kernel(int64_t nnz, const I* __restrict__ rows,
                                                         const I* __restrict__ cols, const V* __restrict__ vals,
                                                         Entry<V, I>* out)
{
    for (int64_t i = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; i < nnz;
         i += static_cast<int64_t>(gridDim.x) * blockDim.x) {
        Entry<V, I> e;
        e.row = rows[i];
        e.column = cols[i];
        e.value = vals[i];
        out[i] = e;
    }
}

}  // namespace
}  // namespace gkob200

extern "C" {

size_t gkob200_setup_sort_workspace_bytes(int64_t nnz, int value_bytes, int index_bytes)
{
    return sort_ws_bytes(nnz, static_cast<size_t>(value_bytes), static_cast<size_t>(index_bytes));
}
size_t gkob200_setup_compact_workspace_bytes(int64_t nnz) { return scan_ws_bytes(nnz); }

#define GKOB200_DEF_SETUP(V, VT, I, IT)                                                                            \
    int gkob200_coo_sort_row_major_##V##_##I(void* st, int64_t n_rows, int64_t n_cols, int64_t nnz, IT* rows,       \
                                             IT* cols, VT* vals, void* ws, size_t wsb)                              \
    { return sort_row_major<VT, IT>(as_stream(st), n_rows, n_cols, nnz, rows, cols, vals, ws, wsb); }              \
    int gkob200_coo_sum_duplicates_##V##_##I(void* st, int64_t nnz, const IT* rows, const IT* cols, const VT* vals, \
                                             IT* out_rows, IT* out_cols, VT* out_vals, int64_t* out_nnz, void* ws,  \
                                             size_t wsb)                                                            \
    { return compact<VT, IT, true>(as_stream(st), nnz, rows, cols, vals, out_rows, out_cols, out_vals, out_nnz, ws, wsb); } \
    int gkob200_coo_remove_zeros_##V##_##I(void* st, int64_t nnz, const IT* rows, const IT* cols, const VT* vals,   \
                                           IT* out_rows, IT* out_cols, VT* out_vals, int64_t* out_nnz, void* ws,    \
                                           size_t wsb)                                                              \
    { return compact<VT, IT, false>(as_stream(st), nnz, rows, cols, vals, out_rows, out_cols, out_vals, out_nnz, ws, wsb); } \
    int gkob200_csr_transpose_##V##_##I(void* st, int64_t n_rows, int64_t n_cols, int64_t nnz, const IT* rp,        \
                                        const IT* ci, const VT* va, IT* out_rp, IT* out_ci, VT* out_va, void* ws,   \
                                        size_t wsb)                                                                 \
    { return csr_transpose<VT, IT>(as_stream(st), n_rows, n_cols, nnz, rp, ci, va, out_rp, out_ci, out_va, ws, wsb); } \
    int gkob200_csr_sort_by_column_index_##V##_##I(void* st, int64_t n_rows, int64_t n_cols, int64_t nnz,           \
                                                   const IT* rp, IT* ci, VT* va, void* ws, size_t wsb)              \
    { return csr_sort_by_column_index<VT, IT>(as_stream(st), n_rows, n_cols, nnz, rp, ci, va, ws, wsb); }

GKOB200_DEF_SETUP(f64, double, i32, int32_t)
GKOB200_DEF_SETUP(f32, float, i32, int32_t)
GKOB200_DEF_SETUP(f64, double, i64, int64_t)
GKOB200_DEF_SETUP(f32, float, i64, int64_t)
#undef GKOB200_DEF_SETUP

#define GKOB200_DEF_AOS(V, VT, I, IT)                                                                              \
    int gkob200_aos_to_soa_##V##_##I(void* st, int64_t nnz, const void* entries, IT* rows, IT* cols, VT* vals)     \
    {                                                                                                              \
        if (nnz < 0 || (nnz > 0 && (!entries || !rows || !cols || !vals))) return GKOB200_EINVAL;                  \
        if (nnz == 0) return 0;                                                                                    \
        aos_to_soa_kernel<VT, IT><<<grid_for(nnz, 256, 8), 256, 0, as_stream(st)>>>(                               \
            nnz, static_cast<const Entry<VT, IT>*>(entries), rows, cols, vals);                                    \
        GKOB200_CHECK_LAUNCH();                                                                                    \
        return 0;                                                                                                  \
    }                                                                                                              \
    int gkob200_soa_to_aos_##V##_##I(void* st, int64_t nnz, const IT* rows, const IT* cols, const VT* vals,        \
                                     void* entries)                                                                \
    {                                                                                                              \
        if (nnz < 0 || (nnz > 0 && (!entries || !rows || !cols || !vals))) return GKOB200_EINVAL;                  \
        if (nnz == 0) return 0;                                                                                    \
        soa_to_aos_kernel<VT, IT><<<grid_for(nnz, 256, 8), 256, 0, as_stream(st)>>>(                               \
            nnz, rows, cols, vals, static_cast<Entry<VT, IT>*>(entries));                                          \
        GKOB200_CHECK_LAUNCH();                                                                                    \
        return 0;                                                                                                  \
    }
GKOB200_DEF_AOS(f64, double, i32, int32_t)
GKOB200_DEF_AOS(f32, float, i32, int32_t)
GKOB200_DEF_AOS(f64, double, i64, int64_t)
GKOB200_DEF_AOS(f32, float, i64, int64_t)
#undef GKOB200_DEF_AOS

}  // extern "C"
