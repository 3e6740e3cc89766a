// shim/kernels_setup.cpp — the set-up side of the hot path in namespace gko::kernels::cuda, bound
// to the C-ABI of libgko_b200.so: what Csr::read / convert_to(Ell|Sellp|Hybrid|Coo), Dense::row_gather,
// distributed::Partition and distributed::Matrix::read_distributed dispatch to on a CudaExecutor
// (SURVEY.md §8 rows a7, a8, a16, a17 and f1).  Same conventions as kernels.cpp: every symbol
// defined here overrides the reference's NotCompiled stub at link time.
#include <gko_b200.h>
#include <gko_b200_solver.h>

#include <cuda_runtime.h>

#include <ginkgo/core/base/array.hpp>
#include <ginkgo/core/base/device_matrix_data.hpp>
#include <ginkgo/core/base/exception_helpers.hpp>
#include <ginkgo/core/base/executor.hpp>
#include <ginkgo/core/distributed/partition.hpp>
#include <ginkgo/core/matrix/coo.hpp>
#include <ginkgo/core/matrix/csr.hpp>
#include <ginkgo/core/matrix/dense.hpp>
#include <ginkgo/core/matrix/ell.hpp>
#include <ginkgo/core/matrix/hybrid.hpp>
#include <ginkgo/core/matrix/sellp.hpp>

#include "core/base/device_matrix_data_kernels.hpp"
#include "core/components/format_conversion_kernels.hpp"
#include "core/components/prefix_sum_kernels.hpp"
#include "core/distributed/matrix_kernels.hpp"
#include "core/distributed/partition_kernels.hpp"
#include "core/distributed/vector_kernels.hpp"
#include "core/matrix/csr_kernels.hpp"
#include "core/matrix/dense_kernels.hpp"
#include "core/matrix/ell_kernels.hpp"
#include "core/matrix/hybrid_kernels.hpp"
#include "core/matrix/sellp_kernels.hpp"

namespace gko {
namespace kernels {
namespace cuda {
namespace {

using Exec = std::shared_ptr<const CudaExecutor>;
template <typename V>
using D = matrix::Dense<V>;

inline void check(int rc, const char* fn)
{
    if (rc > 0) throw CudaError(__FILE__, __LINE__, fn, rc);
    if (rc < 0) throw NotSupported(__FILE__, __LINE__, fn, "libgko_b200 (argument / unsupported)");
}
#define B200(call) check(call, #call)
#define CUDA_OK(call)                                                                  \
    do {                                                                               \
        const cudaError_t e__ = (call);                                                \
        if (e__ != cudaSuccess) throw CudaError(__FILE__, __LINE__, #call, (int)e__);  \
    } while (0)

void* const kStream = nullptr;  // Ginkgo 1.5 launches everything on stream 0

// scratch that lives for one call (the reference's kernels allocate temporaries through
// array<T>(exec, n) the same way, e.g. cuda/matrix/csr_kernels.cu:134-135)
struct Temp {
    array<char> buf;
    Temp(const Exec& exec, size_t bytes) : buf(exec, bytes + 64) {}
    void* get() { return buf.get_data(); }
};

template <typename T>
T to_host(const T* dev)
{
    T v{};
    CUDA_OK(cudaMemcpy(&v, dev, sizeof(T), cudaMemcpyDeviceToHost));
    return v;
}

// ---- typed dispatch -----------------------------------------------------------------------
#define TYPED_VI(name)                                                                                                 \
    template <typename... A> inline int name(double, int32, A... a) { return gkob200_##name##_f64_i32(a...); }          \
    template <typename... A> inline int name(float, int32, A... a) { return gkob200_##name##_f32_i32(a...); }           \
    template <typename... A> inline int name(double, int64, A... a) { return gkob200_##name##_f64_i64(a...); }          \
    template <typename... A> inline int name(float, int64, A... a) { return gkob200_##name##_f32_i64(a...); }
#define TYPED_IDX(name)                                                                            \
    template <typename... A> inline int name(int32, A... a) { return gkob200_##name##_i32(a...); } \
    template <typename... A> inline int name(int64, A... a) { return gkob200_##name##_i64(a...); }
namespace t {
TYPED_VI(csr_convert_to_ell)
TYPED_VI(csr_convert_to_sellp)
TYPED_VI(csr_convert_to_hybrid)
TYPED_VI(aos_to_soa)
TYPED_VI(soa_to_aos)
TYPED_VI(coo_sort_row_major)
TYPED_VI(coo_sum_duplicates)
TYPED_VI(coo_remove_zeros)
TYPED_VI(dense_row_gather)
TYPED_IDX(convert_ptrs_to_idxs)
TYPED_IDX(convert_idxs_to_ptrs)
TYPED_IDX(convert_ptrs_to_sizes)
TYPED_IDX(compute_max_row_nnz)
TYPED_IDX(sellp_compute_slice_sets)
template <typename... A> inline int prefix_sum(int32, A... a) { return gkob200_prefix_sum_i32(a...); }
template <typename... A> inline int prefix_sum(int64, A... a) { return gkob200_prefix_sum_i64(a...); }
template <typename... A> inline int prefix_sum(size_type, A... a)
{
    return gkob200_prefix_sum_u64(a...);
}
inline int build_local_nonlocal(double, const double* v, double* lv, double* nv, void* s, int64_t nnz, const int64_t* r,
                                const int64_t* c, int64_t rnr, const int64_t* rrb, const int32_t* rpi, const int32_t* rrs,
                                int64_t cnr, const int64_t* crb, const int32_t* cpi, const int32_t* crs, int64_t gcols,
                                int32_t nparts, int32_t lp, int32_t* lr, int32_t* lc, int32_t* nr, int32_t* nc, int32_t* g,
                                int32_t* rs, int64_t* n2g, int64_t* counts)
{
    return gkob200_dist_build_local_nonlocal_f64(s, nnz, r, c, v, rnr, rrb, rpi, rrs, cnr, crb, cpi, crs, gcols, nparts, lp,
                                                 lr, lc, lv, nr, nc, nv, g, rs, n2g, counts);
}
inline int build_local_nonlocal(float, const float* v, float* lv, float* nv, void* s, int64_t nnz, const int64_t* r,
                                const int64_t* c, int64_t rnr, const int64_t* rrb, const int32_t* rpi, const int32_t* rrs,
                                int64_t cnr, const int64_t* crb, const int32_t* cpi, const int32_t* crs, int64_t gcols,
                                int32_t nparts, int32_t lp, int32_t* lr, int32_t* lc, int32_t* nr, int32_t* nc, int32_t* g,
                                int32_t* rs, int64_t* n2g, int64_t* counts)
{
    return gkob200_dist_build_local_nonlocal_f32(s, nnz, r, c, v, rnr, rrb, rpi, rrs, cnr, crb, cpi, crs, gcols, nparts, lp,
                                                 lr, lc, lv, nr, nc, nv, g, rs, n2g, counts);
}
}  // namespace t

}  // namespace

// ===================================== components ====================================
namespace components {

// [core/components/prefix_sum_kernels.hpp:67; replaces common/cuda_hip/components/prefix_sum_kernels.hpp.inc]
template <typename IndexType>
void prefix_sum(Exec exec, IndexType* counts, size_type num_entries)
{
    if (num_entries == 0) return;
    const size_t wsb = gkob200_prefix_sum_workspace_bytes((int64_t)num_entries);
    Temp ws(exec, wsb);
    B200(t::prefix_sum(IndexType{}, kStream, counts, (int64_t)num_entries, ws.get(), wsb));
}
template void prefix_sum<int32>(Exec, int32*, size_type);
template void prefix_sum<int64>(Exec, int64*, size_type);
template void prefix_sum<size_type>(Exec, size_type*, size_type);

// [core/components/format_conversion_kernels.hpp:52-73]; index type == pointer type only
// (the combinations Csr / Coo / Hybrid / Sellp conversions use)
template <typename IndexType, typename RowPtrType>
void convert_ptrs_to_idxs(Exec, const RowPtrType* ptrs, size_type num_blocks, IndexType* idxs);
template <typename IndexType, typename RowPtrType>
void convert_idxs_to_ptrs(Exec, const IndexType* idxs, size_type num_idxs, size_type num_blocks, RowPtrType* ptrs);
#define DEF_CONV(I)                                                                                          \
    template <>                                                                                              \
    void convert_ptrs_to_idxs<I, I>(Exec, const I* ptrs, size_type num_blocks, I* idxs)                      \
    {                                                                                                        \
        B200(t::convert_ptrs_to_idxs(I{}, kStream, ptrs, (int64_t)num_blocks, idxs));                        \
    }                                                                                                        \
    template <>                                                                                              \
    void convert_idxs_to_ptrs<I, I>(Exec, const I* idxs, size_type num_idxs, size_type num_blocks, I* ptrs)  \
    {                                                                                                        \
        B200(t::convert_idxs_to_ptrs(I{}, kStream, idxs, (int64_t)num_idxs, (int64_t)num_blocks, ptrs));     \
    }
DEF_CONV(int32)
DEF_CONV(int64)
template <typename RowPtrType>
void convert_ptrs_to_sizes(Exec, const RowPtrType* ptrs, size_type num_blocks, size_type* sizes)
{
    B200(t::convert_ptrs_to_sizes(RowPtrType{}, kStream, ptrs, (int64_t)num_blocks, reinterpret_cast<uint64_t*>(sizes)));
}
template void convert_ptrs_to_sizes<int32>(Exec, const int32*, size_type, size_type*);
template void convert_ptrs_to_sizes<int64>(Exec, const int64*, size_type, size_type*);

// device_matrix_data [core/base/device_matrix_data_kernels.hpp:54-80]
template <typename V, typename I>
void aos_to_soa(Exec, const array<matrix_data_entry<V, I>>& in, device_matrix_data<V, I>& out)
{
    B200(t::aos_to_soa(V{}, I{}, kStream, (int64_t)in.get_num_elems(), static_cast<const void*>(in.get_const_data()),
                       out.get_row_idxs(), out.get_col_idxs(), out.get_values()));
}
template <typename V, typename I>
void soa_to_aos(Exec, const device_matrix_data<V, I>& in, array<matrix_data_entry<V, I>>& out)
{
    B200(t::soa_to_aos(V{}, I{}, kStream, (int64_t)in.get_num_elems(), in.get_const_row_idxs(), in.get_const_col_idxs(),
                       in.get_const_values(), static_cast<void*>(out.get_data())));
}
template <typename V, typename I>
void sort_row_major(Exec exec, device_matrix_data<V, I>& data)
{
    const int64_t nnz = (int64_t)data.get_num_elems();
    const size_t wsb = gkob200_setup_sort_workspace_bytes(nnz, (int)sizeof(V), (int)sizeof(I));
    Temp ws(exec, wsb);
    B200(t::coo_sort_row_major(V{}, I{}, kStream, (int64_t)data.get_size()[0], (int64_t)data.get_size()[1], nnz,
                               data.get_row_idxs(), data.get_col_idxs(), data.get_values(), ws.get(), wsb));
}
// compacting kernels: results go to fresh arrays of the final size (like the reference's
// thrust implementation, cuda/base/device_matrix_data_kernels.cu)
template <typename V, typename I, typename Fn>
void compact_into(Exec exec, array<V>& values, array<I>& row_idxs, array<I>& col_idxs, Fn fn)
{
    const int64_t nnz = (int64_t)values.get_num_elems();
    if (nnz == 0) return;
    const size_t wsb = gkob200_setup_compact_workspace_bytes(nnz);
    Temp ws(exec, wsb);
    array<V> nv(exec, nnz);
    array<I> nr(exec, nnz), nc(exec, nnz);
    array<int64> count(exec, 1);
    B200(fn(nnz, nr.get_data(), nc.get_data(), nv.get_data(), count.get_data(), ws.get(), wsb));
    const auto m = static_cast<size_type>(to_host(count.get_const_data()));
    if (m == static_cast<size_type>(nnz)) {
        values = std::move(nv);
        row_idxs = std::move(nr);
        col_idxs = std::move(nc);
        return;
    }
    array<V> fv(exec, m);
    array<I> fr(exec, m), fc(exec, m);
    exec->copy(m, nv.get_const_data(), fv.get_data());
    exec->copy(m, nr.get_const_data(), fr.get_data());
    exec->copy(m, nc.get_const_data(), fc.get_data());
    values = std::move(fv);
    row_idxs = std::move(fr);
    col_idxs = std::move(fc);
}
template <typename V, typename I>
void remove_zeros(Exec exec, array<V>& values, array<I>& row_idxs, array<I>& col_idxs)
{
    const I* r = row_idxs.get_const_data();
    const I* c = col_idxs.get_const_data();
    const V* v = values.get_const_data();
    compact_into<V, I>(exec, values, row_idxs, col_idxs,
                       [&](int64_t nnz, I* nr, I* nc, V* nv, int64* cnt, void* ws, size_t wsb) {
                           return t::coo_remove_zeros(V{}, I{}, kStream, nnz, r, c, v, nr, nc, nv, cnt, ws, wsb);
                       });
}
template <typename V, typename I>
void sum_duplicates(Exec exec, size_type, array<V>& values, array<I>& row_idxs, array<I>& col_idxs)
{
    const I* r = row_idxs.get_const_data();
    const I* c = col_idxs.get_const_data();
    const V* v = values.get_const_data();
    compact_into<V, I>(exec, values, row_idxs, col_idxs,
                       [&](int64_t nnz, I* nr, I* nc, V* nv, int64* cnt, void* ws, size_t wsb) {
                           return t::coo_sum_duplicates(V{}, I{}, kStream, nnz, r, c, v, nr, nc, nv, cnt, ws, wsb);
                       });
}
#define INST_DMD(V, I)                                                                                         \
    template void aos_to_soa<V, I>(Exec, const array<matrix_data_entry<V, I>>&, device_matrix_data<V, I>&);     \
    template void soa_to_aos<V, I>(Exec, const device_matrix_data<V, I>&, array<matrix_data_entry<V, I>>&);     \
    template void sort_row_major<V, I>(Exec, device_matrix_data<V, I>&);                                        \
    template void remove_zeros<V, I>(Exec, array<V>&, array<I>&, array<I>&);                                    \
    template void sum_duplicates<V, I>(Exec, size_type, array<V>&, array<I>&, array<I>&);
INST_DMD(double, int32)
INST_DMD(float, int32)
INST_DMD(double, int64)
INST_DMD(float, int64)

}  // namespace components

// ===================================== csr conversions ================================
namespace csr {

// [core/matrix/csr_kernels.hpp:101-122; replaces common/unified/matrix/csr_kernels.cpp:137-243]
template <typename V, typename I>
void convert_to_ell(Exec, const matrix::Csr<V, I>* source, matrix::Ell<V, I>* result)
{
    B200(t::csr_convert_to_ell(V{}, I{}, kStream, (int64_t)source->get_size()[0], source->get_const_row_ptrs(),
                               source->get_const_col_idxs(), source->get_const_values(),
                               (int64_t)result->get_num_stored_elements_per_row(), (int64_t)result->get_stride(),
                               result->get_col_idxs(), result->get_values()));
}
template <typename V, typename I>
void convert_to_sellp(Exec, const matrix::Csr<V, I>* source, matrix::Sellp<V, I>* result)
{
    B200(t::csr_convert_to_sellp(V{}, I{}, kStream, (int64_t)source->get_size()[0], source->get_const_row_ptrs(),
                                 source->get_const_col_idxs(), source->get_const_values(),
                                 (int64_t)result->get_slice_size(),
                                 reinterpret_cast<const uint64_t*>(result->get_const_slice_sets()),
                                 result->get_col_idxs(), result->get_values()));
}
template <typename V, typename I>
void convert_to_hybrid(Exec, const matrix::Csr<V, I>* source, const int64* coo_row_ptrs, matrix::Hybrid<V, I>* result)
{
    B200(t::csr_convert_to_hybrid(V{}, I{}, kStream, (int64_t)source->get_size()[0], source->get_const_row_ptrs(),
                                  source->get_const_col_idxs(), source->get_const_values(), coo_row_ptrs,
                                  (int64_t)result->get_ell_stride(), (int64_t)result->get_ell_num_stored_elements_per_row(),
                                  result->get_ell_col_idxs(), result->get_ell_values(), result->get_coo_row_idxs(),
                                  result->get_coo_col_idxs(), result->get_coo_values()));
}
#define INST_CSR_CONV(V, I)                                                                        \
    template void convert_to_ell<V, I>(Exec, const matrix::Csr<V, I>*, matrix::Ell<V, I>*);         \
    template void convert_to_sellp<V, I>(Exec, const matrix::Csr<V, I>*, matrix::Sellp<V, I>*);     \
    template void convert_to_hybrid<V, I>(Exec, const matrix::Csr<V, I>*, const int64*, matrix::Hybrid<V, I>*);
INST_CSR_CONV(double, int32)
INST_CSR_CONV(float, int32)
INST_CSR_CONV(double, int64)
INST_CSR_CONV(float, int64)

}  // namespace csr

namespace ell {
// [core/matrix/ell_kernels.hpp:68-71; oracle reference/matrix/ell_kernels.cpp:159-168]
template <typename IndexType>
void compute_max_row_nnz(Exec exec, const array<IndexType>& row_ptrs, size_type& max_nnz)
{
    array<size_type> out(exec, 1);
    B200(t::compute_max_row_nnz(IndexType{}, kStream, row_ptrs.get_const_data(), (int64_t)row_ptrs.get_num_elems() - 1,
                                reinterpret_cast<uint64_t*>(out.get_data())));
    max_nnz = to_host(out.get_const_data());
}
template void compute_max_row_nnz<int32>(Exec, const array<int32>&, size_type&);
// [core/matrix/ell_kernels.hpp:84-87; oracle reference/matrix/ell_kernels.cpp:222-234] — what
// Ell::operator= runs on the SOURCE's executor: the stored columns, stride to stride
template <typename V, typename I>
void copy(Exec, const matrix::Ell<V, I>* source, matrix::Ell<V, I>* result)
{
    const size_t rows = source->get_size()[0], width = source->get_num_stored_elements_per_row();
    if (rows == 0 || width == 0) return;
    CUDA_OK(cudaMemcpy2DAsync(result->get_values(), result->get_stride() * sizeof(V), source->get_const_values(),
                              source->get_stride() * sizeof(V), rows * sizeof(V), width, cudaMemcpyDeviceToDevice, nullptr));
    CUDA_OK(cudaMemcpy2DAsync(result->get_col_idxs(), result->get_stride() * sizeof(I), source->get_const_col_idxs(),
                              source->get_stride() * sizeof(I), rows * sizeof(I), width, cudaMemcpyDeviceToDevice, nullptr));
}
template void copy<double, int32>(Exec, const matrix::Ell<double, int32>*, matrix::Ell<double, int32>*);
template void copy<float, int32>(Exec, const matrix::Ell<float, int32>*, matrix::Ell<float, int32>*);
template void copy<double, int64>(Exec, const matrix::Ell<double, int64>*, matrix::Ell<double, int64>*);
template void copy<float, int64>(Exec, const matrix::Ell<float, int64>*, matrix::Ell<float, int64>*);

template void compute_max_row_nnz<int64>(Exec, const array<int64>&, size_type&);
}  // namespace ell

namespace sellp {
// [core/matrix/sellp_kernels.hpp:71-75; oracle reference/matrix/sellp_kernels.cpp:134-160]
template <typename IndexType>
void compute_slice_sets(Exec exec, const array<IndexType>& row_ptrs, size_type slice_size, size_type stride_factor,
                        size_type* slice_sets, size_type* slice_lengths)
{
    const int64_t n = (int64_t)row_ptrs.get_num_elems() - 1;
    const size_t wsb = gkob200_prefix_sum_workspace_bytes(n / (int64_t)slice_size + 2);
    Temp ws(exec, wsb);
    B200(t::sellp_compute_slice_sets(IndexType{}, kStream, row_ptrs.get_const_data(), n, (int64_t)slice_size,
                                     (int64_t)stride_factor, reinterpret_cast<uint64_t*>(slice_sets),
                                     reinterpret_cast<uint64_t*>(slice_lengths), ws.get(), wsb));
}
template void compute_slice_sets<int32>(Exec, const array<int32>&, size_type, size_type, size_type*, size_type*);
template void compute_slice_sets<int64>(Exec, const array<int64>&, size_type, size_type, size_type*, size_type*);
}  // namespace sellp

namespace hybrid {
// [core/matrix/hybrid_kernels.hpp:50-57; common/unified/matrix/hybrid_kernels.cpp:51-76]
void compute_row_nnz(Exec, const array<int64>& row_ptrs, size_type* row_nnzs)
{
    B200(gkob200_convert_ptrs_to_sizes_i64(kStream, row_ptrs.get_const_data(), (int64_t)row_ptrs.get_num_elems() - 1,
                                           reinterpret_cast<uint64_t*>(row_nnzs)));
}
void compute_coo_row_ptrs(Exec exec, const array<size_type>& row_nnz, size_type ell_lim, int64* coo_row_ptrs)
{
    const int64_t n = (int64_t)row_nnz.get_num_elems();
    const size_t wsb = gkob200_prefix_sum_workspace_bytes(n + 1);
    Temp ws(exec, wsb);
    B200(gkob200_hybrid_compute_coo_row_ptrs(kStream, reinterpret_cast<const uint64_t*>(row_nnz.get_const_data()), n,
                                             (uint64_t)ell_lim, coo_row_ptrs, ws.get(), wsb));
}
}  // namespace hybrid

// ===================================== dense =========================================
namespace dense {
// [core/matrix/dense_kernels.hpp:247-251; oracle reference/matrix/dense_kernels.cpp row_gather]
template <typename V, typename O, typename I>
void row_gather(Exec, const array<I>* gather_indices, const D<V>* orig, D<O>* row_collection);
#define DEF_GATHER(V, I)                                                                                        \
    template <>                                                                                                 \
    void row_gather<V, V, I>(Exec, const array<I>* idx, const D<V>* orig, D<V>* out)                            \
    {                                                                                                           \
        B200(t::dense_row_gather(V{}, I{}, kStream, (int64_t)out->get_size()[0], (int64_t)out->get_size()[1],   \
                                 idx->get_const_data(), orig->get_const_values(), (int64_t)orig->get_stride(),  \
                                 out->get_values(), (int64_t)out->get_stride()));                               \
    }
DEF_GATHER(double, int32)
DEF_GATHER(float, int32)
DEF_GATHER(double, int64)
DEF_GATHER(float, int64)
}  // namespace dense

// ===================================== distributed ===================================
namespace partition {

using comm_index_type = experimental::distributed::comm_index_type;

// [core/distributed/partition_kernels.hpp:47-83; oracle reference/distributed/partition_kernels.cpp:42-160]
void count_ranges(Exec exec, const array<comm_index_type>& mapping, size_type& num_ranges)
{
    // number of maximal runs of equal part ids: build_from_mapping computes it on the device
    const int64_t n = (int64_t)mapping.get_num_elems();
    if (n == 0) {
        num_ranges = 0;
        return;
    }
    const size_t wsb = (size_t)(n + 2) * 4 + gkob200_prefix_sum_workspace_bytes(n + 1) + 64;
    Temp ws(exec, wsb);
    array<int64> bounds(exec, n + 1), nr(exec, 1);
    array<comm_index_type> ids(exec, n);
    B200(gkob200_partition_build_from_mapping_i64(kStream, n, mapping.get_const_data(), bounds.get_data(), ids.get_data(),
                                                  nr.get_data(), ws.get(), wsb));
    num_ranges = static_cast<size_type>(to_host(nr.get_const_data()));
}
template <typename GlobalIndexType>
void build_from_contiguous(Exec, const array<GlobalIndexType>& ranges, GlobalIndexType* range_bounds,
                           comm_index_type* part_ids);
template <>
void build_from_contiguous<int64>(Exec, const array<int64>& ranges, int64* range_bounds, comm_index_type* part_ids)
{
    B200(gkob200_partition_build_from_contiguous_i64(kStream, (int32_t)ranges.get_num_elems() - 1, ranges.get_const_data(),
                                                     range_bounds, part_ids));
}
template <typename GlobalIndexType>
void build_from_mapping(Exec, const array<comm_index_type>& mapping, GlobalIndexType* range_bounds,
                        comm_index_type* part_ids);
template <>
void build_from_mapping<int64>(Exec exec, const array<comm_index_type>& mapping, int64* range_bounds,
                               comm_index_type* part_ids)
{
    const int64_t n = (int64_t)mapping.get_num_elems();
    if (n == 0) return;
    const size_t wsb = (size_t)(n + 2) * 4 + gkob200_prefix_sum_workspace_bytes(n + 1) + 64;
    Temp ws(exec, wsb);
    // the C-ABI writes up to n + 1 bounds / n ids before it knows the range count; the caller's
    // arrays are sized for exactly num_ranges (count_ranges): go through full-size temporaries
    array<int64> bounds(exec, n + 1), nr(exec, 1);
    array<comm_index_type> ids(exec, n);
    B200(gkob200_partition_build_from_mapping_i64(kStream, n, mapping.get_const_data(), bounds.get_data(), ids.get_data(),
                                                  nr.get_data(), ws.get(), wsb));
    const auto k = static_cast<size_type>(to_host(nr.get_const_data()));
    exec->copy(k + 1, bounds.get_const_data(), range_bounds);
    exec->copy(k, ids.get_const_data(), part_ids);
}
template <typename GlobalIndexType>
void build_ranges_from_global_size(Exec, comm_index_type num_parts, GlobalIndexType global_size,
                                   array<GlobalIndexType>& ranges);
template <>
void build_ranges_from_global_size<int64>(Exec, comm_index_type num_parts, int64 global_size, array<int64>& ranges)
{
    B200(gkob200_partition_build_ranges_from_global_size_i64(kStream, num_parts, global_size, ranges.get_data()));
}
template <typename LocalIndexType, typename GlobalIndexType>
void build_starting_indices(Exec, const GlobalIndexType* range_offsets, const int* range_parts, size_type num_ranges,
                            comm_index_type num_parts, comm_index_type& num_empty_parts, LocalIndexType* ranks,
                            LocalIndexType* sizes);
template <>
void build_starting_indices<int32, int64>(Exec exec, const int64* range_offsets, const int* range_parts,
                                          size_type num_ranges, comm_index_type num_parts,
                                          comm_index_type& num_empty_parts, int32* ranks, int32* sizes)
{
    array<int32> ne(exec, 1);
    B200(gkob200_partition_build_starting_indices_i32_i64(kStream, range_offsets, range_parts, (int64_t)num_ranges,
                                                          num_parts, ne.get_data(), ranks, sizes));
    num_empty_parts = to_host(ne.get_const_data());
}

}  // namespace partition

namespace distributed_matrix {

using comm_index_type = experimental::distributed::comm_index_type;

// [core/distributed/matrix_kernels.hpp:51-67; oracle reference/distributed/matrix_kernels.cpp:49-236]
template <typename V, typename LI, typename GI>
void build_local_nonlocal(Exec exec, const device_matrix_data<V, GI>& input,
                          const experimental::distributed::Partition<LI, GI>* row_partition,
                          const experimental::distributed::Partition<LI, GI>* col_partition,
                          comm_index_type local_part, array<LI>& local_row_idxs, array<LI>& local_col_idxs,
                          array<V>& local_values, array<LI>& non_local_row_idxs, array<LI>& non_local_col_idxs,
                          array<V>& non_local_values, array<LI>& local_gather_idxs,
                          array<comm_index_type>& recv_sizes, array<GI>& non_local_to_global);
template <typename V>
void build_impl(Exec exec, const device_matrix_data<V, int64>& input,
                const experimental::distributed::Partition<int32, int64>* rp,
                const experimental::distributed::Partition<int32, int64>* cp, comm_index_type local_part,
                array<int32>& lr, array<int32>& lc, array<V>& lv, array<int32>& nr, array<int32>& nc, array<V>& nv,
                array<int32>& gather, array<comm_index_type>& recv_sizes, array<int64>& n2g)
{
    const int64_t nnz = (int64_t)input.get_num_elems();
    const size_type cap = static_cast<size_type>(nnz > 0 ? nnz : 1);
    array<int32> tlr(exec, cap), tlc(exec, cap), tnr(exec, cap), tnc(exec, cap), tg(exec, cap);
    array<V> tlv(exec, cap), tnv(exec, cap);
    array<int64> tn2g(exec, cap);
    recv_sizes.resize_and_reset(static_cast<size_type>(rp->get_num_parts()));
    int64_t counts[3] = {0, 0, 0};
    B200(t::build_local_nonlocal(V{}, input.get_const_values(), tlv.get_data(), tnv.get_data(), kStream, nnz,
                                 input.get_const_row_idxs(), input.get_const_col_idxs(), (int64_t)rp->get_num_ranges(),
                                 rp->get_range_bounds(), rp->get_part_ids(), rp->get_range_starting_indices(),
                                 (int64_t)cp->get_num_ranges(), cp->get_range_bounds(), cp->get_part_ids(),
                                 cp->get_range_starting_indices(), (int64_t)cp->get_size(), (int32_t)rp->get_num_parts(),
                                 (int32_t)local_part, tlr.get_data(), tlc.get_data(), tnr.get_data(), tnc.get_data(),
                                 tg.get_data(), recv_sizes.get_data(), tn2g.get_data(), counts));
    auto take = [&](auto& dst, const auto& src, int64_t m) {
        dst.resize_and_reset(static_cast<size_type>(m));
        if (m > 0) exec->copy(static_cast<size_type>(m), src.get_const_data(), dst.get_data());
    };
    take(lr, tlr, counts[0]);
    take(lc, tlc, counts[0]);
    take(lv, tlv, counts[0]);
    take(nr, tnr, counts[1]);
    take(nc, tnc, counts[1]);
    take(nv, tnv, counts[1]);
    take(gather, tg, counts[2]);
    take(n2g, tn2g, counts[2]);
}
template <>
void build_local_nonlocal<double, int32, int64>(
    Exec exec, const device_matrix_data<double, int64>& input,
    const experimental::distributed::Partition<int32, int64>* rp,
    const experimental::distributed::Partition<int32, int64>* cp, comm_index_type local_part, array<int32>& lr,
    array<int32>& lc, array<double>& lv, array<int32>& nr, array<int32>& nc, array<double>& nv, array<int32>& gather,
    array<comm_index_type>& recv_sizes, array<int64>& n2g)
{
    build_impl<double>(exec, input, rp, cp, local_part, lr, lc, lv, nr, nc, nv, gather, recv_sizes, n2g);
}
template <>
void build_local_nonlocal<float, int32, int64>(
    Exec exec, const device_matrix_data<float, int64>& input,
    const experimental::distributed::Partition<int32, int64>* rp,
    const experimental::distributed::Partition<int32, int64>* cp, comm_index_type local_part, array<int32>& lr,
    array<int32>& lc, array<float>& lv, array<int32>& nr, array<int32>& nc, array<float>& nv, array<int32>& gather,
    array<comm_index_type>& recv_sizes, array<int64>& n2g)
{
    build_impl<float>(exec, input, rp, cp, local_part, lr, lc, lv, nr, nc, nv, gather, recv_sizes, n2g);
}

}  // namespace distributed_matrix

namespace distributed_vector {

using comm_index_type = experimental::distributed::comm_index_type;

// [core/distributed/vector_kernels.hpp:52-59; oracle reference/distributed/vector_kernels.cpp]
template <typename V, typename LI, typename GI>
void build_local(Exec, const device_matrix_data<V, GI>& input, const experimental::distributed::Partition<LI, GI>* partition,
                 comm_index_type local_part, matrix::Dense<V>* local_mtx);
template <>
void build_local<double, int32, int64>(Exec, const device_matrix_data<double, int64>& input,
                                       const experimental::distributed::Partition<int32, int64>* p,
                                       comm_index_type local_part, matrix::Dense<double>* local_mtx)
{
    B200(gkob200_dist_vector_build_local_f64(kStream, (int64_t)input.get_num_elems(), input.get_const_row_idxs(),
                                             input.get_const_col_idxs(), input.get_const_values(),
                                             (int64_t)p->get_num_ranges(), p->get_range_bounds(), p->get_part_ids(),
                                             p->get_range_starting_indices(), (int32_t)local_part,
                                             local_mtx->get_values(), (int64_t)local_mtx->get_stride()));
}
template <>
void build_local<float, int32, int64>(Exec, const device_matrix_data<float, int64>& input,
                                      const experimental::distributed::Partition<int32, int64>* p,
                                      comm_index_type local_part, matrix::Dense<float>* local_mtx)
{
    B200(gkob200_dist_vector_build_local_f32(kStream, (int64_t)input.get_num_elems(), input.get_const_row_idxs(),
                                             input.get_const_col_idxs(), input.get_const_values(),
                                             (int64_t)p->get_num_ranges(), p->get_range_bounds(), p->get_part_ids(),
                                             p->get_range_starting_indices(), (int32_t)local_part,
                                             local_mtx->get_values(), (int64_t)local_mtx->get_stride()));
}

}  // namespace distributed_vector

}  // namespace cuda
}  // namespace kernels
}  // namespace gko
