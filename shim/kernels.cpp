// shim/kernels.cpp — the hot-path kernels of namespace gko::kernels::cuda, bound 1:1 to the
// C-ABI of libgko_b200.so (include/gko_b200.h).  Compiled against the Ginkgo 1.5.0 headers
// with the same host compiler; linked together with the reference's own NotCompiled stub
// object (core/device_hooks/cuda_hooks.cpp, symbols weakened) into a drop-in
// libginkgo_cuda.so: every symbol defined here overrides its stub, everything else keeps
// throwing NotCompiled.  See INTEGRATION.md.
//
// Real value types (double, float) x (int32, int64); complex stays on the stubs.
#include <gko_b200.h>

#include <cuda_runtime.h>

#include <ginkgo/core/base/array.hpp>
#include <ginkgo/core/base/exception_helpers.hpp>
#include <ginkgo/core/base/executor.hpp>
#include <ginkgo/core/matrix/coo.hpp>
#include <ginkgo/core/matrix/csr.hpp>
#include <ginkgo/core/matrix/dense.hpp>
#include <ginkgo/core/matrix/diagonal.hpp>
#include <ginkgo/core/matrix/ell.hpp>
#include <ginkgo/core/matrix/sellp.hpp>
#include <ginkgo/core/preconditioner/jacobi.hpp>
#include <ginkgo/core/stop/stopping_status.hpp>

#include "core/components/fill_array_kernels.hpp"
#include "core/components/prefix_sum_kernels.hpp"
#include "core/matrix/coo_kernels.hpp"
#include "core/matrix/csr_kernels.hpp"
#include "core/matrix/dense_kernels.hpp"
#include "core/matrix/ell_kernels.hpp"
#include "core/matrix/sellp_kernels.hpp"
#include "core/preconditioner/jacobi_kernels.hpp"
#include "core/solver/bicg_kernels.hpp"
#include "core/solver/bicgstab_kernels.hpp"
#include "core/solver/ir_kernels.hpp"
#include "core/solver/cg_kernels.hpp"
#include "core/solver/cgs_kernels.hpp"
#include "core/solver/fcg_kernels.hpp"
#include "core/solver/common_gmres_kernels.hpp"
#include "core/solver/gmres_kernels.hpp"
#include "core/stop/criterion_kernels.hpp"
#include "core/stop/residual_norm_kernels.hpp"

namespace gko {
namespace kernels {
namespace cuda {
namespace {

using Exec = std::shared_ptr<const CudaExecutor>;

inline void check(int rc, const char* fn)
{
    if (rc > 0) throw CudaError(__FILE__, __LINE__, fn, rc);
    if (rc < 0) throw NotSupported(__FILE__, __LINE__, fn, "libgko_b200 (argument / unsupported)");
}
#define B200(call) check(call, #call)

void* const kStream = nullptr;  // Ginkgo 1.5 launches everything on stream 0

inline uint8* status_ptr(array<stopping_status>* a) { return reinterpret_cast<uint8*>(a->get_data()); }
inline const uint8* status_ptr(const array<stopping_status>* a) { return reinterpret_cast<const uint8*>(a->get_const_data()); }
inline uint8* status_ptr(stopping_status* a) { return reinterpret_cast<uint8*>(a); }
inline const uint8* status_ptr(const stopping_status* a) { return reinterpret_cast<const uint8*>(a); }

// reduction scratch: the solver's `array<char>& tmp` (GKO_SOLVER_STOP_REDUCTION_ARRAYS),
// grown once and zero-initialised, then reused by every reduction of the solve
inline void* reduce_ws(const Exec& exec, array<char>& tmp)
{
    if (tmp.get_num_elems() < GKOB200_REDUCE_WS_BYTES) {
        tmp.resize_and_reset(GKOB200_REDUCE_WS_BYTES);
        B200(gkob200_reduce_ws_init(kStream, tmp.get_data()));
    }
    return tmp.get_data();
}

// per-matrix scratch of the merge-path CSR kernel, cached per executor
struct Scratch {
    void* p = nullptr;
    size_t bytes = 0;
    void* get(size_t need)
    {
        if (need > bytes) {
            if (p) cudaFree(p);
            if (cudaMalloc(&p, need) != cudaSuccess) throw AllocationError(__FILE__, __LINE__, "cuda", need);
            bytes = need;
        }
        return p;
    }
};
Scratch& scratch()
{
    static thread_local Scratch s;
    return s;
}

inline int strategy_of(const std::string& s)
{
    if (s == "classical") return GKOB200_CSR_CLASSICAL;
    if (s == "merge_path" || s == "load_balance") return GKOB200_CSR_MERGE_PATH;
    return GKOB200_CSR_MERGE_PATH;  // automatical / sparselib without statistics: always-correct kernel
}

// ---- typed dispatch -----------------------------------------------------------------
#define TYPED(name)                                                                     \
    template <typename... A> inline int name(double, A... a) { return gkob200_##name##_f64(a...); } \
    template <typename... A> inline int name(float, A... a) { return gkob200_##name##_f32(a...); }
#define TYPED_I(name)                                                                                  \
    template <typename... A> inline int name(double, int32, A... a) { return gkob200_##name##_f64_i32(a...); } \
    template <typename... A> inline int name(float, int32, A... a) { return gkob200_##name##_f32_i32(a...); }  \
    template <typename... A> inline int name(double, int64, A... a) { return gkob200_##name##_f64_i64(a...); } \
    template <typename... A> inline int name(float, int64, A... a) { return gkob200_##name##_f32_i64(a...); }
namespace t {
TYPED_I(csr_spmv)
TYPED_I(ell_spmv)
TYPED_I(sellp_spmv)
TYPED_I(coo_spmv)
TYPED_I(coo_spmv2)
TYPED(dense_fill)
TYPED(dense_scale)
TYPED(dense_inv_scale)
TYPED(dense_add_scaled)
TYPED(dense_sub_scaled)
TYPED(dense_compute_dot)
TYPED(dense_compute_norm2)
TYPED(dense_compute_squared_norm2)
TYPED(dense_compute_norm1)
TYPED(dense_compute_sqrt)
TYPED(cg_initialize)
TYPED(cg_step_1)
TYPED(cg_step_2)
TYPED(bicg_initialize)
TYPED(bicg_step_1)
TYPED(bicg_step_2)
TYPED(fcg_initialize)
TYPED(fcg_step_1)
TYPED(fcg_step_2)
TYPED(cgs_initialize)
TYPED(cgs_step_1)
TYPED(cgs_step_2)
TYPED(cgs_step_3)
TYPED_I(csr_transpose)
TYPED_I(csr_sort_by_column_index)
TYPED(bicgstab_initialize)
TYPED(bicgstab_step_1)
TYPED(bicgstab_step_2)
TYPED(bicgstab_step_3)
TYPED(bicgstab_finalize)
TYPED(gmres_initialize)
TYPED(gmres_restart)
TYPED(gmres_multi_axpy)
TYPED(gmres_hessenberg_qr)
TYPED(gmres_solve_krylov)
TYPED(residual_norm)
TYPED(implicit_residual_norm)
TYPED(jacobi_invert_diagonal)
TYPED(jacobi_simple_scalar_apply)
TYPED(jacobi_scalar_apply)
TYPED(jacobi_block_generate)
TYPED(jacobi_block_simple_apply)
TYPED(jacobi_block_apply)
TYPED(jacobi_block_transpose)
template <typename... A> inline int csr_extract_diagonal(double, A... a) { return gkob200_csr_extract_diagonal_f64_i32(a...); }
template <typename... A> inline int csr_extract_diagonal(float, A... a) { return gkob200_csr_extract_diagonal_f32_i32(a...); }
}  // namespace t

template <typename V>
using D = matrix::Dense<V>;

}  // namespace

// ===================================== csr ===========================================
namespace csr {

template <typename V, typename I>
void spmv(Exec exec, const matrix::Csr<V, I>* a, const D<V>* b, D<V>* c)
{
    const int strat = strategy_of(a->get_strategy()->get_name());
    const size_t wsb = gkob200_csr_spmv_workspace_bytes(a->get_size()[0], a->get_num_stored_elements(), 1, sizeof(V));
    B200(t::csr_spmv(V{}, I{}, kStream, (int64_t)a->get_size()[0], (int64_t)a->get_size()[1],
                     (int64_t)a->get_num_stored_elements(), a->get_const_row_ptrs(), a->get_const_col_idxs(),
                     a->get_const_values(), b->get_const_values(), (int64_t)b->get_stride(), (int64_t)b->get_size()[1],
                     (const V*)nullptr, (const V*)nullptr, c->get_values(), (int64_t)c->get_stride(), strat, (int64_t)0,
                     strat == GKOB200_CSR_MERGE_PATH ? scratch().get(wsb) : nullptr, wsb));
}

template <typename V, typename I>
void advanced_spmv(Exec exec, const D<V>* alpha, const matrix::Csr<V, I>* a, const D<V>* b, const D<V>* beta, D<V>* c)
{
    const int strat = strategy_of(a->get_strategy()->get_name());
    const size_t wsb = gkob200_csr_spmv_workspace_bytes(a->get_size()[0], a->get_num_stored_elements(), 1, sizeof(V));
    B200(t::csr_spmv(V{}, I{}, kStream, (int64_t)a->get_size()[0], (int64_t)a->get_size()[1],
                     (int64_t)a->get_num_stored_elements(), a->get_const_row_ptrs(), a->get_const_col_idxs(),
                     a->get_const_values(), b->get_const_values(), (int64_t)b->get_stride(), (int64_t)b->get_size()[1],
                     alpha->get_const_values(), beta->get_const_values(), c->get_values(), (int64_t)c->get_stride(),
                     strat, (int64_t)0, strat == GKOB200_CSR_MERGE_PATH ? scratch().get(wsb) : nullptr, wsb));
}

template <typename V, typename I>
void extract_diagonal(Exec exec, const matrix::Csr<V, I>* orig, matrix::Diagonal<V>* diag)
{
    B200(t::csr_extract_diagonal(V{}, kStream, (int64_t)orig->get_size()[0], (int64_t)orig->get_size()[1],
                                 orig->get_const_row_ptrs(), orig->get_const_col_idxs(), orig->get_const_values(),
                                 diag->get_values()));
}

// [cuda/matrix/csr_kernels.cu transpose (cusparse csr2csc) / sort_by_column_index (cusparse csrsort)]
template <typename V, typename I>
void transpose(Exec, const matrix::Csr<V, I>* orig, matrix::Csr<V, I>* trans)
{
    const int64_t n = (int64_t)orig->get_size()[0], m = (int64_t)orig->get_size()[1];
    const int64_t nnz = (int64_t)orig->get_num_stored_elements();
    const size_t wsb = gkob200_setup_sort_workspace_bytes(nnz, (int)sizeof(V), (int)sizeof(I));
    B200(t::csr_transpose(V{}, I{}, kStream, n, m, nnz, orig->get_const_row_ptrs(), orig->get_const_col_idxs(),
                          orig->get_const_values(), trans->get_row_ptrs(), trans->get_col_idxs(), trans->get_values(),
                          scratch().get(wsb), wsb));
}
template <typename V, typename I>
void sort_by_column_index(Exec, matrix::Csr<V, I>* to_sort)
{
    const int64_t n = (int64_t)to_sort->get_size()[0], m = (int64_t)to_sort->get_size()[1];
    const int64_t nnz = (int64_t)to_sort->get_num_stored_elements();
    const size_t wsb = gkob200_setup_sort_workspace_bytes(nnz, (int)sizeof(V), (int)sizeof(I));
    B200(t::csr_sort_by_column_index(V{}, I{}, kStream, n, m, nnz, to_sort->get_const_row_ptrs(), to_sort->get_col_idxs(),
                                     to_sort->get_values(), scratch().get(wsb), wsb));
}
// real value types: the conjugate transpose is the transpose (what solver::Bicg asks for)
template <typename V, typename I>
void conj_transpose(Exec exec, const matrix::Csr<V, I>* orig, matrix::Csr<V, I>* trans)
{
    transpose<V, I>(exec, orig, trans);
}
template void conj_transpose<double, int32>(Exec, const matrix::Csr<double, int32>*, matrix::Csr<double, int32>*);
template void conj_transpose<float, int32>(Exec, const matrix::Csr<float, int32>*, matrix::Csr<float, int32>*);
template void conj_transpose<double, int64>(Exec, const matrix::Csr<double, int64>*, matrix::Csr<double, int64>*);
template void conj_transpose<float, int64>(Exec, const matrix::Csr<float, int64>*, matrix::Csr<float, int64>*);
template void transpose<double, int32>(Exec, const matrix::Csr<double, int32>*, matrix::Csr<double, int32>*);
template void transpose<float, int32>(Exec, const matrix::Csr<float, int32>*, matrix::Csr<float, int32>*);
template void transpose<double, int64>(Exec, const matrix::Csr<double, int64>*, matrix::Csr<double, int64>*);
template void transpose<float, int64>(Exec, const matrix::Csr<float, int64>*, matrix::Csr<float, int64>*);
template void sort_by_column_index<double, int32>(Exec, matrix::Csr<double, int32>*);
template void sort_by_column_index<float, int32>(Exec, matrix::Csr<float, int32>*);
template void sort_by_column_index<double, int64>(Exec, matrix::Csr<double, int64>*);
template void sort_by_column_index<float, int64>(Exec, matrix::Csr<float, int64>*);

#define INST_CSR(V, I)                                                                                     \
    template void spmv<V, I>(Exec, const matrix::Csr<V, I>*, const D<V>*, D<V>*);                           \
    template void advanced_spmv<V, I>(Exec, const D<V>*, const matrix::Csr<V, I>*, const D<V>*, const D<V>*, D<V>*);
INST_CSR(double, int32)
INST_CSR(float, int32)
INST_CSR(double, int64)
INST_CSR(float, int64)
template void extract_diagonal<double, int32>(Exec, const matrix::Csr<double, int32>*, matrix::Diagonal<double>*);
template void extract_diagonal<float, int32>(Exec, const matrix::Csr<float, int32>*, matrix::Diagonal<float>*);

}  // namespace csr

// ===================================== ell / sellp / coo ==============================
namespace ell {

template <typename IV, typename MV, typename OV, typename I>
void spmv(Exec exec, const matrix::Ell<MV, I>* a, const D<IV>* b, D<OV>* c)
{
    B200(t::ell_spmv(MV{}, I{}, kStream, (int64_t)a->get_size()[0], (int64_t)a->get_size()[1], (int64_t)a->get_stride(),
                     (int64_t)a->get_num_stored_elements_per_row(), a->get_const_col_idxs(), a->get_const_values(),
                     b->get_const_values(), (int64_t)b->get_stride(), (int64_t)b->get_size()[1], (const MV*)nullptr,
                     (const MV*)nullptr, c->get_values(), (int64_t)c->get_stride()));
}
template <typename IV, typename MV, typename OV, typename I>
void advanced_spmv(Exec exec, const D<MV>* alpha, const matrix::Ell<MV, I>* a, const D<IV>* b, const D<OV>* beta, D<OV>* c)
{
    B200(t::ell_spmv(MV{}, I{}, kStream, (int64_t)a->get_size()[0], (int64_t)a->get_size()[1], (int64_t)a->get_stride(),
                     (int64_t)a->get_num_stored_elements_per_row(), a->get_const_col_idxs(), a->get_const_values(),
                     b->get_const_values(), (int64_t)b->get_stride(), (int64_t)b->get_size()[1], alpha->get_const_values(),
                     beta->get_const_values(), c->get_values(), (int64_t)c->get_stride()));
}
#define INST_ELL(V, I)                                                                                      \
    template void spmv<V, V, V, I>(Exec, const matrix::Ell<V, I>*, const D<V>*, D<V>*);                      \
    template void advanced_spmv<V, V, V, I>(Exec, const D<V>*, const matrix::Ell<V, I>*, const D<V>*, const D<V>*, D<V>*);
INST_ELL(double, int32)
INST_ELL(float, int32)
INST_ELL(double, int64)
INST_ELL(float, int64)

}  // namespace ell

namespace sellp {

template <typename V, typename I>
void spmv(Exec exec, const matrix::Sellp<V, I>* a, const D<V>* b, D<V>* c)
{
    B200(t::sellp_spmv(V{}, I{}, kStream, (int64_t)a->get_size()[0], (int64_t)a->get_size()[1],
                       (int64_t)a->get_slice_size(), reinterpret_cast<const uint64_t*>(a->get_const_slice_sets()),
                       reinterpret_cast<const uint64_t*>(a->get_const_slice_lengths()), a->get_const_col_idxs(),
                       a->get_const_values(), b->get_const_values(), (int64_t)b->get_stride(), (int64_t)b->get_size()[1],
                       (const V*)nullptr, (const V*)nullptr, c->get_values(), (int64_t)c->get_stride()));
}
template <typename V, typename I>
void advanced_spmv(Exec exec, const D<V>* alpha, const matrix::Sellp<V, I>* a, const D<V>* b, const D<V>* beta, D<V>* c)
{
    B200(t::sellp_spmv(V{}, I{}, kStream, (int64_t)a->get_size()[0], (int64_t)a->get_size()[1],
                       (int64_t)a->get_slice_size(), reinterpret_cast<const uint64_t*>(a->get_const_slice_sets()),
                       reinterpret_cast<const uint64_t*>(a->get_const_slice_lengths()), a->get_const_col_idxs(),
                       a->get_const_values(), b->get_const_values(), (int64_t)b->get_stride(), (int64_t)b->get_size()[1],
                       alpha->get_const_values(), beta->get_const_values(), c->get_values(), (int64_t)c->get_stride()));
}
#define INST_SELLP(V, I)                                                                  \
    template void spmv<V, I>(Exec, const matrix::Sellp<V, I>*, const D<V>*, D<V>*);        \
    template void advanced_spmv<V, I>(Exec, const D<V>*, const matrix::Sellp<V, I>*, const D<V>*, const D<V>*, D<V>*);
INST_SELLP(double, int32)
INST_SELLP(float, int32)
INST_SELLP(double, int64)
INST_SELLP(float, int64)

}  // namespace sellp

namespace coo {

template <typename V, typename I>
void coo_call(const matrix::Coo<V, I>* a, const D<V>* b, const V* alpha, const V* beta, bool accumulate, D<V>* c)
{
    const size_t wsb = gkob200_coo_spmv_workspace_bytes(a->get_num_stored_elements(), sizeof(V));
    void* ws = scratch().get(wsb);
    if (accumulate)
        B200(t::coo_spmv2(V{}, I{}, kStream, (int64_t)a->get_size()[0], (int64_t)a->get_size()[1],
                          (int64_t)a->get_num_stored_elements(), a->get_const_row_idxs(), a->get_const_col_idxs(),
                          a->get_const_values(), b->get_const_values(), (int64_t)b->get_stride(),
                          (int64_t)b->get_size()[1], alpha, c->get_values(), (int64_t)c->get_stride(), ws, wsb));
    else
        B200(t::coo_spmv(V{}, I{}, kStream, (int64_t)a->get_size()[0], (int64_t)a->get_size()[1],
                         (int64_t)a->get_num_stored_elements(), a->get_const_row_idxs(), a->get_const_col_idxs(),
                         a->get_const_values(), b->get_const_values(), (int64_t)b->get_stride(),
                         (int64_t)b->get_size()[1], alpha, beta, c->get_values(), (int64_t)c->get_stride(), ws, wsb));
}
template <typename V, typename I>
void spmv(Exec, const matrix::Coo<V, I>* a, const D<V>* b, D<V>* c) { coo_call<V, I>(a, b, nullptr, nullptr, false, c); }
template <typename V, typename I>
void advanced_spmv(Exec, const D<V>* alpha, const matrix::Coo<V, I>* a, const D<V>* b, const D<V>* beta, D<V>* c)
{
    coo_call<V, I>(a, b, alpha->get_const_values(), beta->get_const_values(), false, c);
}
template <typename V, typename I>
void spmv2(Exec, const matrix::Coo<V, I>* a, const D<V>* b, D<V>* c) { coo_call<V, I>(a, b, nullptr, nullptr, true, c); }
template <typename V, typename I>
void advanced_spmv2(Exec, const D<V>* alpha, const matrix::Coo<V, I>* a, const D<V>* b, D<V>* c)
{
    coo_call<V, I>(a, b, alpha->get_const_values(), nullptr, true, c);
}
#define INST_COO(V, I)                                                                                     \
    template void spmv<V, I>(Exec, const matrix::Coo<V, I>*, const D<V>*, D<V>*);                           \
    template void advanced_spmv<V, I>(Exec, const D<V>*, const matrix::Coo<V, I>*, const D<V>*, const D<V>*, D<V>*); \
    template void spmv2<V, I>(Exec, const matrix::Coo<V, I>*, const D<V>*, D<V>*);                          \
    template void advanced_spmv2<V, I>(Exec, const D<V>*, const matrix::Coo<V, I>*, const D<V>*, D<V>*);
INST_COO(double, int32)
INST_COO(float, int32)
INST_COO(double, int64)
INST_COO(float, int64)

}  // namespace coo

// ===================================== components / dense =============================
namespace components {

template <typename T>
void fill_array(Exec, T* data, size_type n, T val)
{
    B200(gkob200_fill_array(kStream, data, (int64_t)n, (int)sizeof(T), &val));
}
template void fill_array<double>(Exec, double*, size_type, double);
template void fill_array<float>(Exec, float*, size_type, float);
template void fill_array<int32>(Exec, int32*, size_type, int32);
template void fill_array<int64>(Exec, int64*, size_type, int64);
template void fill_array<size_type>(Exec, size_type*, size_type, size_type);

}  // namespace components

namespace dense {

template <typename V>
void fill(Exec, D<V>* m, V value)
{
    B200(t::dense_fill(V{}, kStream, (int64_t)m->get_size()[0], (int64_t)m->get_size()[1], m->get_values(),
                       (int64_t)m->get_stride(), value));
}
template <typename V, typename S>
void scale(Exec, const D<S>* alpha, D<V>* x)
{
    B200(t::dense_scale(V{}, kStream, (int64_t)x->get_size()[0], (int64_t)x->get_size()[1], alpha->get_const_values(),
                        (int64_t)alpha->get_size()[1], x->get_values(), (int64_t)x->get_stride()));
}
template <typename V, typename S>
void inv_scale(Exec, const D<S>* alpha, D<V>* x)
{
    B200(t::dense_inv_scale(V{}, kStream, (int64_t)x->get_size()[0], (int64_t)x->get_size()[1], alpha->get_const_values(),
                            (int64_t)alpha->get_size()[1], x->get_values(), (int64_t)x->get_stride()));
}
template <typename V, typename S>
void add_scaled(Exec, const D<S>* alpha, const D<V>* x, D<V>* y)
{
    B200(t::dense_add_scaled(V{}, kStream, (int64_t)x->get_size()[0], (int64_t)x->get_size()[1], alpha->get_const_values(),
                             (int64_t)alpha->get_size()[1], x->get_const_values(), (int64_t)x->get_stride(),
                             y->get_values(), (int64_t)y->get_stride()));
}
template <typename V, typename S>
void sub_scaled(Exec, const D<S>* alpha, const D<V>* x, D<V>* y)
{
    B200(t::dense_sub_scaled(V{}, kStream, (int64_t)x->get_size()[0], (int64_t)x->get_size()[1], alpha->get_const_values(),
                             (int64_t)alpha->get_size()[1], x->get_const_values(), (int64_t)x->get_stride(),
                             y->get_values(), (int64_t)y->get_stride()));
}
template <typename V>
void compute_dot(Exec exec, const D<V>* x, const D<V>* y, D<V>* result, array<char>& tmp)
{
    B200(t::dense_compute_dot(V{}, kStream, (int64_t)x->get_size()[0], (int64_t)x->get_size()[1], x->get_const_values(),
                              (int64_t)x->get_stride(), y->get_const_values(), (int64_t)y->get_stride(),
                              result->get_values(), reduce_ws(exec, tmp)));
}
template <typename V>
void compute_dot_dispatch(Exec exec, const D<V>* x, const D<V>* y, D<V>* r, array<char>& tmp) { compute_dot(exec, x, y, r, tmp); }
template <typename V>
void compute_conj_dot(Exec exec, const D<V>* x, const D<V>* y, D<V>* r, array<char>& tmp) { compute_dot(exec, x, y, r, tmp); }
template <typename V>
void compute_conj_dot_dispatch(Exec exec, const D<V>* x, const D<V>* y, D<V>* r, array<char>& tmp) { compute_dot(exec, x, y, r, tmp); }
template <typename V>
void compute_norm2(Exec exec, const D<V>* x, D<remove_complex<V>>* result, array<char>& tmp)
{
    B200(t::dense_compute_norm2(V{}, kStream, (int64_t)x->get_size()[0], (int64_t)x->get_size()[1], x->get_const_values(),
                                (int64_t)x->get_stride(), result->get_values(), reduce_ws(exec, tmp)));
}
template <typename V>
void compute_norm2_dispatch(Exec exec, const D<V>* x, D<remove_complex<V>>* r, array<char>& tmp) { compute_norm2(exec, x, r, tmp); }
template <typename V>
void compute_squared_norm2(Exec exec, const D<V>* x, D<remove_complex<V>>* result, array<char>& tmp)
{
    B200(t::dense_compute_squared_norm2(V{}, kStream, (int64_t)x->get_size()[0], (int64_t)x->get_size()[1],
                                        x->get_const_values(), (int64_t)x->get_stride(), result->get_values(),
                                        reduce_ws(exec, tmp)));
}
template <typename V>
void compute_norm1(Exec exec, const D<V>* x, D<remove_complex<V>>* result, array<char>& tmp)
{
    B200(t::dense_compute_norm1(V{}, kStream, (int64_t)x->get_size()[0], (int64_t)x->get_size()[1], x->get_const_values(),
                                (int64_t)x->get_stride(), result->get_values(), reduce_ws(exec, tmp)));
}
template <typename V>
void compute_sqrt(Exec, D<V>* data)
{
    B200(t::dense_compute_sqrt(V{}, kStream, (int64_t)data->get_size()[1], data->get_values()));
}
template <typename In, typename Out>
void copy(Exec, const D<In>* in, D<Out>* out);
template <>
void copy<double, double>(Exec, const D<double>* in, D<double>* out)
{
    B200(gkob200_dense_copy_f64(kStream, (int64_t)in->get_size()[0], (int64_t)in->get_size()[1], in->get_const_values(),
                                (int64_t)in->get_stride(), out->get_values(), (int64_t)out->get_stride()));
}
template <>
void copy<float, float>(Exec, const D<float>* in, D<float>* out)
{
    B200(gkob200_dense_copy_f32(kStream, (int64_t)in->get_size()[0], (int64_t)in->get_size()[1], in->get_const_values(),
                                (int64_t)in->get_stride(), out->get_values(), (int64_t)out->get_stride()));
}

#define INST_DENSE(V)                                                                                   \
    template void fill<V>(Exec, D<V>*, V);                                                               \
    template void scale<V, V>(Exec, const D<V>*, D<V>*);                                                 \
    template void inv_scale<V, V>(Exec, const D<V>*, D<V>*);                                             \
    template void add_scaled<V, V>(Exec, const D<V>*, const D<V>*, D<V>*);                               \
    template void sub_scaled<V, V>(Exec, const D<V>*, const D<V>*, D<V>*);                               \
    template void compute_dot<V>(Exec, const D<V>*, const D<V>*, D<V>*, array<char>&);                   \
    template void compute_dot_dispatch<V>(Exec, const D<V>*, const D<V>*, D<V>*, array<char>&);          \
    template void compute_conj_dot<V>(Exec, const D<V>*, const D<V>*, D<V>*, array<char>&);              \
    template void compute_conj_dot_dispatch<V>(Exec, const D<V>*, const D<V>*, D<V>*, array<char>&);     \
    template void compute_norm2<V>(Exec, const D<V>*, D<V>*, array<char>&);                              \
    template void compute_norm2_dispatch<V>(Exec, const D<V>*, D<V>*, array<char>&);                     \
    template void compute_squared_norm2<V>(Exec, const D<V>*, D<V>*, array<char>&);                      \
    template void compute_norm1<V>(Exec, const D<V>*, D<V>*, array<char>&);                              \
    template void compute_sqrt<V>(Exec, D<V>*);
INST_DENSE(double)
INST_DENSE(float)

}  // namespace dense

// ===================================== solvers ========================================
namespace cg {

template <typename V>
void initialize(Exec, const D<V>* b, D<V>* r, D<V>* z, D<V>* p, D<V>* q, D<V>* prev_rho, D<V>* rho,
                array<stopping_status>* stop)
{
    B200(t::cg_initialize(V{}, kStream, (int64_t)b->get_size()[0], (int64_t)b->get_size()[1], b->get_const_values(),
                          (int64_t)b->get_stride(), r->get_values(), z->get_values(), p->get_values(), q->get_values(),
                          (int64_t)r->get_stride(), prev_rho->get_values(), rho->get_values(), status_ptr(stop)));
}
template <typename V>
void step_1(Exec, D<V>* p, const D<V>* z, const D<V>* rho, const D<V>* prev_rho, const array<stopping_status>* stop)
{
    B200(t::cg_step_1(V{}, kStream, (int64_t)p->get_size()[0], (int64_t)p->get_size()[1], p->get_values(),
                      z->get_const_values(), (int64_t)p->get_stride(), rho->get_const_values(),
                      prev_rho->get_const_values(), status_ptr(stop)));
}
template <typename V>
void step_2(Exec, D<V>* x, D<V>* r, const D<V>* p, const D<V>* q, const D<V>* beta, const D<V>* rho,
            const array<stopping_status>* stop)
{
    B200(t::cg_step_2(V{}, kStream, (int64_t)x->get_size()[0], (int64_t)x->get_size()[1], x->get_values(),
                      (int64_t)x->get_stride(), r->get_values(), p->get_const_values(), q->get_const_values(),
                      (int64_t)r->get_stride(), beta->get_const_values(), rho->get_const_values(), status_ptr(stop)));
}
#define INST_CG(V)                                                                                              \
    template void initialize<V>(Exec, const D<V>*, D<V>*, D<V>*, D<V>*, D<V>*, D<V>*, D<V>*, array<stopping_status>*); \
    template void step_1<V>(Exec, D<V>*, const D<V>*, const D<V>*, const D<V>*, const array<stopping_status>*);        \
    template void step_2<V>(Exec, D<V>*, D<V>*, const D<V>*, const D<V>*, const D<V>*, const D<V>*, const array<stopping_status>*);
INST_CG(double)
INST_CG(float)

}  // namespace cg

// [core/solver/bicg_kernels.hpp; the host loop core/solver/bicg.cpp:120-243 transposes the system
// matrix with csr::transpose and the preconditioner with Transposable::conj_transpose]
namespace bicg {

template <typename V>
void initialize(Exec, const D<V>* b, D<V>* r, D<V>* z, D<V>* p, D<V>* q, D<V>* prev_rho, D<V>* rho, D<V>* r2, D<V>* z2,
                D<V>* p2, D<V>* q2, array<stopping_status>* stop_status)
{
    B200(t::bicg_initialize(V{}, kStream, (int64_t)b->get_size()[0], (int64_t)b->get_size()[1], b->get_const_values(),
                            (int64_t)b->get_stride(), r->get_values(), z->get_values(), p->get_values(), q->get_values(),
                            r2->get_values(), z2->get_values(), p2->get_values(), q2->get_values(),
                            (int64_t)r->get_stride(), prev_rho->get_values(), rho->get_values(), status_ptr(stop_status)));
}
template <typename V>
void step_1(Exec, D<V>* p, const D<V>* z, D<V>* p2, const D<V>* z2, const D<V>* rho, const D<V>* prev_rho,
            const array<stopping_status>* stop_status)
{
    B200(t::bicg_step_1(V{}, kStream, (int64_t)p->get_size()[0], (int64_t)p->get_size()[1], p->get_values(),
                        z->get_const_values(), p2->get_values(), z2->get_const_values(), (int64_t)p->get_stride(),
                        rho->get_const_values(), prev_rho->get_const_values(), status_ptr(stop_status)));
}
template <typename V>
void step_2(Exec, D<V>* x, D<V>* r, D<V>* r2, const D<V>* p, const D<V>* q, const D<V>* q2, const D<V>* beta,
            const D<V>* rho, const array<stopping_status>* stop_status)
{
    B200(t::bicg_step_2(V{}, kStream, (int64_t)x->get_size()[0], (int64_t)x->get_size()[1], x->get_values(),
                        (int64_t)x->get_stride(), r->get_values(), r2->get_values(), p->get_const_values(),
                        q->get_const_values(), q2->get_const_values(), (int64_t)r->get_stride(), beta->get_const_values(),
                        rho->get_const_values(), status_ptr(stop_status)));
}
#define INST_BICG(V)                                                                                                  \
    template void initialize<V>(Exec, const D<V>*, D<V>*, D<V>*, D<V>*, D<V>*, D<V>*, D<V>*, D<V>*, D<V>*, D<V>*,      \
                                D<V>*, array<stopping_status>*);                                                      \
    template void step_1<V>(Exec, D<V>*, const D<V>*, D<V>*, const D<V>*, const D<V>*, const D<V>*,                    \
                            const array<stopping_status>*);                                                           \
    template void step_2<V>(Exec, D<V>*, D<V>*, D<V>*, const D<V>*, const D<V>*, const D<V>*, const D<V>*,             \
                            const D<V>*, const array<stopping_status>*);
INST_BICG(double)
INST_BICG(float)

}  // namespace bicg

// [core/solver/ir_kernels.hpp:54-56]: solver::Ir's only kernel resets the stopping status
namespace ir {
void initialize(Exec, array<stopping_status>* stop_status)
{
    const uint8 zero = 0;   // stopping_status::reset()
    B200(gkob200_fill_array(kStream, status_ptr(stop_status), (int64_t)stop_status->get_num_elems(), 1, &zero));
}
}  // namespace ir

namespace fcg {

template <typename V>
void initialize(Exec, const D<V>* b, D<V>* r, D<V>* z, D<V>* p, D<V>* q, D<V>* t, D<V>* prev_rho, D<V>* rho,
                D<V>* rho_t, array<stopping_status>* stop)
{
    B200(t::fcg_initialize(V{}, kStream, (int64_t)b->get_size()[0], (int64_t)b->get_size()[1], b->get_const_values(),
                           (int64_t)b->get_stride(), r->get_values(), z->get_values(), p->get_values(), q->get_values(),
                           t->get_values(), (int64_t)r->get_stride(), prev_rho->get_values(), rho->get_values(),
                           rho_t->get_values(), status_ptr(stop)));
}
template <typename V>
void step_1(Exec, D<V>* p, const D<V>* z, const D<V>* rho_t, const D<V>* prev_rho, const array<stopping_status>* stop)
{
    B200(t::fcg_step_1(V{}, kStream, (int64_t)p->get_size()[0], (int64_t)p->get_size()[1], p->get_values(),
                       z->get_const_values(), (int64_t)p->get_stride(), rho_t->get_const_values(),
                       prev_rho->get_const_values(), status_ptr(stop)));
}
template <typename V>
void step_2(Exec, D<V>* x, D<V>* r, D<V>* t, const D<V>* p, const D<V>* q, const D<V>* beta, const D<V>* rho,
            const array<stopping_status>* stop)
{
    B200(t::fcg_step_2(V{}, kStream, (int64_t)x->get_size()[0], (int64_t)x->get_size()[1], x->get_values(),
                       (int64_t)x->get_stride(), r->get_values(), t->get_values(), p->get_const_values(),
                       q->get_const_values(), (int64_t)r->get_stride(), beta->get_const_values(), rho->get_const_values(),
                       status_ptr(stop)));
}
#define INST_FCG(V)                                                                                                 \
    template void initialize<V>(Exec, const D<V>*, D<V>*, D<V>*, D<V>*, D<V>*, D<V>*, D<V>*, D<V>*, D<V>*,         \
                                array<stopping_status>*);                                                          \
    template void step_1<V>(Exec, D<V>*, const D<V>*, const D<V>*, const D<V>*, const array<stopping_status>*);     \
    template void step_2<V>(Exec, D<V>*, D<V>*, D<V>*, const D<V>*, const D<V>*, const D<V>*, const D<V>*,         \
                            const array<stopping_status>*);
INST_FCG(double)
INST_FCG(float)

}  // namespace fcg

namespace cgs {

template <typename V>
void initialize(Exec, const D<V>* b, D<V>* r, D<V>* r_tld, D<V>* p, D<V>* q, D<V>* u, D<V>* u_hat, D<V>* v_hat, D<V>* t,
                D<V>* alpha, D<V>* beta, D<V>* gamma, D<V>* rho_prev, D<V>* rho, array<stopping_status>* stop)
{
    B200(t::cgs_initialize(V{}, kStream, (int64_t)b->get_size()[0], (int64_t)b->get_size()[1], b->get_const_values(),
                           (int64_t)b->get_stride(), r->get_values(), r_tld->get_values(), p->get_values(),
                           q->get_values(), u->get_values(), u_hat->get_values(), v_hat->get_values(), t->get_values(),
                           (int64_t)r->get_stride(), alpha->get_values(), beta->get_values(), gamma->get_values(),
                           rho_prev->get_values(), rho->get_values(), status_ptr(stop)));
}
template <typename V>
void step_1(Exec, const D<V>* r, D<V>* u, D<V>* p, const D<V>* q, D<V>* beta, const D<V>* rho, const D<V>* rho_prev,
            const array<stopping_status>* stop)
{
    B200(t::cgs_step_1(V{}, kStream, (int64_t)p->get_size()[0], (int64_t)p->get_size()[1], r->get_const_values(),
                       u->get_values(), p->get_values(), q->get_const_values(), (int64_t)p->get_stride(),
                       beta->get_values(), rho->get_const_values(), rho_prev->get_const_values(), status_ptr(stop)));
}
template <typename V>
void step_2(Exec, const D<V>* u, const D<V>* v_hat, D<V>* q, D<V>* t, D<V>* alpha, const D<V>* rho, const D<V>* gamma,
            const array<stopping_status>* stop)
{
    B200(t::cgs_step_2(V{}, kStream, (int64_t)u->get_size()[0], (int64_t)u->get_size()[1], u->get_const_values(),
                       v_hat->get_const_values(), q->get_values(), t->get_values(), (int64_t)u->get_stride(),
                       alpha->get_values(), rho->get_const_values(), gamma->get_const_values(), status_ptr(stop)));
}
template <typename V>
void step_3(Exec, const D<V>* t, const D<V>* u_hat, D<V>* r, D<V>* x, const D<V>* alpha,
            const array<stopping_status>* stop)
{
    B200(t::cgs_step_3(V{}, kStream, (int64_t)x->get_size()[0], (int64_t)x->get_size()[1], t->get_const_values(),
                       u_hat->get_const_values(), r->get_values(), (int64_t)r->get_stride(), x->get_values(),
                       (int64_t)x->get_stride(), alpha->get_const_values(), status_ptr(stop)));
}
#define INST_CGS(V)                                                                                                 \
    template void initialize<V>(Exec, const D<V>*, D<V>*, D<V>*, D<V>*, D<V>*, D<V>*, D<V>*, D<V>*, D<V>*, D<V>*,  \
                                D<V>*, D<V>*, D<V>*, D<V>*, array<stopping_status>*);                              \
    template void step_1<V>(Exec, const D<V>*, D<V>*, D<V>*, const D<V>*, D<V>*, const D<V>*, const D<V>*,         \
                            const array<stopping_status>*);                                                        \
    template void step_2<V>(Exec, const D<V>*, const D<V>*, D<V>*, D<V>*, D<V>*, const D<V>*, const D<V>*,         \
                            const array<stopping_status>*);                                                        \
    template void step_3<V>(Exec, const D<V>*, const D<V>*, D<V>*, D<V>*, const D<V>*, const array<stopping_status>*);
INST_CGS(double)
INST_CGS(float)

}  // namespace cgs

namespace bicgstab {

template <typename V>
void initialize(Exec, const D<V>* b, D<V>* r, D<V>* rr, D<V>* y, D<V>* s, D<V>* t_, D<V>* z, D<V>* v, D<V>* p,
                D<V>* prev_rho, D<V>* rho, D<V>* alpha, D<V>* beta, D<V>* gamma, D<V>* omega,
                array<stopping_status>* stop)
{
    B200(t::bicgstab_initialize(V{}, kStream, (int64_t)b->get_size()[0], (int64_t)b->get_size()[1], b->get_const_values(),
                                (int64_t)b->get_stride(), r->get_values(), rr->get_values(), y->get_values(),
                                s->get_values(), t_->get_values(), z->get_values(), v->get_values(), p->get_values(),
                                (int64_t)r->get_stride(), prev_rho->get_values(), rho->get_values(), alpha->get_values(),
                                beta->get_values(), gamma->get_values(), omega->get_values(), status_ptr(stop)));
}
template <typename V>
void step_1(Exec, const D<V>* r, D<V>* p, const D<V>* v, const D<V>* rho, const D<V>* prev_rho, const D<V>* alpha,
            const D<V>* omega, const array<stopping_status>* stop)
{
    B200(t::bicgstab_step_1(V{}, kStream, (int64_t)p->get_size()[0], (int64_t)p->get_size()[1], r->get_const_values(),
                            p->get_values(), v->get_const_values(), (int64_t)p->get_stride(), rho->get_const_values(),
                            prev_rho->get_const_values(), alpha->get_const_values(), omega->get_const_values(),
                            status_ptr(stop)));
}
template <typename V>
void step_2(Exec, const D<V>* r, D<V>* s, const D<V>* v, const D<V>* rho, D<V>* alpha, const D<V>* beta,
            const array<stopping_status>* stop)
{
    B200(t::bicgstab_step_2(V{}, kStream, (int64_t)s->get_size()[0], (int64_t)s->get_size()[1], r->get_const_values(),
                            s->get_values(), v->get_const_values(), (int64_t)s->get_stride(), rho->get_const_values(),
                            alpha->get_values(), beta->get_const_values(), status_ptr(stop)));
}
template <typename V>
void step_3(Exec, D<V>* x, D<V>* r, const D<V>* s, const D<V>* t_, const D<V>* y, const D<V>* z, const D<V>* alpha,
            const D<V>* beta, const D<V>* gamma, D<V>* omega, const array<stopping_status>* stop)
{
    B200(t::bicgstab_step_3(V{}, kStream, (int64_t)x->get_size()[0], (int64_t)x->get_size()[1], x->get_values(),
                            (int64_t)x->get_stride(), r->get_values(), s->get_const_values(), t_->get_const_values(),
                            y->get_const_values(), z->get_const_values(), (int64_t)r->get_stride(),
                            alpha->get_const_values(), beta->get_const_values(), gamma->get_const_values(),
                            omega->get_values(), status_ptr(stop)));
}
template <typename V>
void finalize(Exec, D<V>* x, const D<V>* y, const D<V>* alpha, array<stopping_status>* stop)
{
    B200(t::bicgstab_finalize(V{}, kStream, (int64_t)x->get_size()[0], (int64_t)x->get_size()[1], x->get_values(),
                              (int64_t)x->get_stride(), y->get_const_values(), (int64_t)y->get_stride(),
                              alpha->get_const_values(), status_ptr(stop)));
}
#define INST_BICGSTAB(V)                                                                                       \
    template void initialize<V>(Exec, const D<V>*, D<V>*, D<V>*, D<V>*, D<V>*, D<V>*, D<V>*, D<V>*, D<V>*, D<V>*, D<V>*, \
                                D<V>*, D<V>*, D<V>*, D<V>*, array<stopping_status>*);                           \
    template void step_1<V>(Exec, const D<V>*, D<V>*, const D<V>*, const D<V>*, const D<V>*, const D<V>*, const D<V>*, \
                            const array<stopping_status>*);                                                    \
    template void step_2<V>(Exec, const D<V>*, D<V>*, const D<V>*, const D<V>*, D<V>*, const D<V>*,              \
                            const array<stopping_status>*);                                                    \
    template void step_3<V>(Exec, D<V>*, D<V>*, const D<V>*, const D<V>*, const D<V>*, const D<V>*, const D<V>*, \
                            const D<V>*, const D<V>*, D<V>*, const array<stopping_status>*);                    \
    template void finalize<V>(Exec, D<V>*, const D<V>*, const D<V>*, array<stopping_status>*);
INST_BICGSTAB(double)
INST_BICGSTAB(float)

}  // namespace bicgstab

namespace common_gmres {

template <typename V>
void initialize(Exec, const D<V>* b, D<V>* residual, D<V>* gsin, D<V>* gcos, stopping_status* stop)
{
    B200(t::gmres_initialize(V{}, kStream, (int64_t)b->get_size()[0], (int64_t)b->get_size()[1],
                             (int64_t)gsin->get_size()[0], b->get_const_values(), (int64_t)b->get_stride(),
                             residual->get_values(), (int64_t)residual->get_stride(), gsin->get_values(),
                             gcos->get_values(), status_ptr(stop)));
}
template <typename V>
void hessenberg_qr(Exec, D<V>* gsin, D<V>* gcos, D<remove_complex<V>>* rnorm, D<V>* rnc, D<V>* hess_iter, size_type iter,
                   size_type* fin, const stopping_status* stop)
{
    B200(t::gmres_hessenberg_qr(V{}, kStream, (int64_t)gsin->get_size()[1], gsin->get_values(), gcos->get_values(),
                                rnorm->get_values(), rnc->get_values(), hess_iter->get_values(),
                                (int64_t)hess_iter->get_stride(), (int64_t)iter, reinterpret_cast<uint64_t*>(fin),
                                status_ptr(stop)));
}
template <typename V>
void solve_krylov(Exec, const D<V>* rnc, const D<V>* hess, D<V>* y, const size_type* fin, const stopping_status* stop)
{
    B200(t::gmres_solve_krylov(V{}, kStream, (int64_t)rnc->get_size()[1], rnc->get_const_values(),
                               hess->get_const_values(), (int64_t)hess->get_stride(), y->get_values(),
                               reinterpret_cast<const uint64_t*>(fin), status_ptr(stop)));
}
#define INST_CGMRES(V)                                                                                    \
    template void initialize<V>(Exec, const D<V>*, D<V>*, D<V>*, D<V>*, stopping_status*);                 \
    template void hessenberg_qr<V>(Exec, D<V>*, D<V>*, D<V>*, D<V>*, D<V>*, size_type, size_type*, const stopping_status*); \
    template void solve_krylov<V>(Exec, const D<V>*, const D<V>*, D<V>*, const size_type*, const stopping_status*);
INST_CGMRES(double)
INST_CGMRES(float)

}  // namespace common_gmres

namespace gmres {

template <typename V>
void restart(Exec, const D<V>* residual, const D<remove_complex<V>>* rnorm, D<V>* rnc, D<V>* kb, size_type* fin)
{
    B200(t::gmres_restart(V{}, kStream, (int64_t)residual->get_size()[0], (int64_t)residual->get_size()[1],
                          residual->get_const_values(), (int64_t)residual->get_stride(), rnorm->get_const_values(),
                          rnc->get_values(), kb->get_values(), reinterpret_cast<uint64_t*>(fin)));
}
template <typename V>
void multi_axpy(Exec, const D<V>* kb, const D<V>* y, D<V>* before, const size_type* fin, stopping_status* stop)
{
    B200(t::gmres_multi_axpy(V{}, kStream, (int64_t)before->get_size()[0], (int64_t)before->get_size()[1],
                             kb->get_const_values(), y->get_const_values(), before->get_values(),
                             (int64_t)before->get_stride(), reinterpret_cast<const uint64_t*>(fin), status_ptr(stop)));
}
#define INST_GMRES(V)                                                                          \
    template void restart<V>(Exec, const D<V>*, const D<V>*, D<V>*, D<V>*, size_type*);         \
    template void multi_axpy<V>(Exec, const D<V>*, const D<V>*, D<V>*, const size_type*, stopping_status*);
INST_GMRES(double)
INST_GMRES(float)

}  // namespace gmres

// ===================================== stopping criteria ===============================
namespace {
// the two booleans travel through a 2-byte pinned buffer: one stream sync instead of the
// reference's two blocking cudaMemcpy (cuda/stop/residual_norm_kernels.cu:117-118)
uint8* flags_buffer()
{
    static thread_local uint8* p = nullptr;
    if (!p && cudaHostAlloc(reinterpret_cast<void**>(&p), 16, cudaHostAllocMapped) != cudaSuccess)
        throw AllocationError(__FILE__, __LINE__, "cudaHostAlloc", 16);
    return p;
}
}  // namespace

namespace residual_norm {
template <typename V>
void residual_norm(Exec, const D<V>* tau, const D<V>* orig_tau, V goal, uint8 id, bool fin,
                   array<stopping_status>* stop, array<bool>*, bool* all_converged, bool* one_changed)
{
    uint8* flags = flags_buffer();
    B200(t::residual_norm(V{}, kStream, (int64_t)tau->get_size()[1], tau->get_const_values(),
                          orig_tau->get_const_values(), goal, id, (int)fin, status_ptr(stop), flags));
    if (cudaStreamSynchronize(nullptr) != cudaSuccess) throw CudaError(__FILE__, __LINE__, "sync", cudaGetLastError());
    *all_converged = flags[0] != 0;
    *one_changed = flags[1] != 0;
}
template void residual_norm<double>(Exec, const D<double>*, const D<double>*, double, uint8, bool,
                                    array<stopping_status>*, array<bool>*, bool*, bool*);
template void residual_norm<float>(Exec, const D<float>*, const D<float>*, float, uint8, bool, array<stopping_status>*,
                                   array<bool>*, bool*, bool*);
}  // namespace residual_norm

namespace implicit_residual_norm {
template <typename V>
void implicit_residual_norm(Exec, const D<V>* tau, const D<remove_complex<V>>* orig_tau, remove_complex<V> goal, uint8 id,
                            bool fin, array<stopping_status>* stop, array<bool>*, bool* all_converged, bool* one_changed)
{
    uint8* flags = flags_buffer();
    B200(t::implicit_residual_norm(V{}, kStream, (int64_t)tau->get_size()[1], tau->get_const_values(),
                                   orig_tau->get_const_values(), goal, id, (int)fin, status_ptr(stop), flags));
    if (cudaStreamSynchronize(nullptr) != cudaSuccess) throw CudaError(__FILE__, __LINE__, "sync", cudaGetLastError());
    *all_converged = flags[0] != 0;
    *one_changed = flags[1] != 0;
}
template void implicit_residual_norm<double>(Exec, const D<double>*, const D<double>*, double, uint8, bool,
                                             array<stopping_status>*, array<bool>*, bool*, bool*);
template void implicit_residual_norm<float>(Exec, const D<float>*, const D<float>*, float, uint8, bool,
                                            array<stopping_status>*, array<bool>*, bool*, bool*);
}  // namespace implicit_residual_norm

namespace set_all_statuses {
void set_all_statuses(Exec, uint8 id, bool fin, array<stopping_status>* stop)
{
    B200(gkob200_set_all_statuses(kStream, (int64_t)stop->get_num_elems(), id, (int)fin, status_ptr(stop)));
}
}  // namespace set_all_statuses

// ===================================== Jacobi ==========================================
namespace jacobi {

// real value types: conj(diag) = diag (what Jacobi::conj_transpose runs for max_block_size 1)
template <typename V>
void scalar_conj(Exec, const array<V>& diag, array<V>& conj_diag)
{
    if (diag.get_num_elems() == 0) return;
    const cudaError_t e = cudaMemcpyAsync(conj_diag.get_data(), diag.get_const_data(), diag.get_num_elems() * sizeof(V),
                                          cudaMemcpyDeviceToDevice, nullptr);
    if (e != cudaSuccess) throw CudaError(__FILE__, __LINE__, "cudaMemcpyAsync", (int)e);
}
template void scalar_conj<double>(Exec, const array<double>&, array<double>&);
template void scalar_conj<float>(Exec, const array<float>&, array<float>&);
template <typename V>
void invert_diagonal(Exec, const array<V>& diag, array<V>& inv)
{
    B200(t::jacobi_invert_diagonal(V{}, kStream, (int64_t)diag.get_num_elems(), diag.get_const_data(), inv.get_data()));
}
template <typename V>
void simple_scalar_apply(Exec, const array<V>& diag, const D<V>* b, D<V>* x)
{
    B200(t::jacobi_simple_scalar_apply(V{}, kStream, (int64_t)x->get_size()[0], (int64_t)x->get_size()[1],
                                       diag.get_const_data(), b->get_const_values(), (int64_t)b->get_stride(),
                                       x->get_values(), (int64_t)x->get_stride()));
}
template <typename V>
void scalar_apply(Exec, const array<V>& diag, const D<V>* alpha, const D<V>* b, const D<V>* beta, D<V>* x)
{
    B200(t::jacobi_scalar_apply(V{}, kStream, (int64_t)x->get_size()[0], (int64_t)x->get_size()[1], diag.get_const_data(),
                                alpha->get_const_values(), b->get_const_values(), (int64_t)b->get_stride(),
                                beta->get_const_values(), x->get_values(), (int64_t)x->get_stride()));
}
template <typename V, typename I>
void find_blocks(Exec, const matrix::Csr<V, I>* mtx, uint32 max_block_size, size_type& num_blocks, array<I>& block_ptrs);
template <typename V>
void find_blocks_i32(const matrix::Csr<V, int32>* mtx, uint32 max_block_size, size_type& num_blocks,
                     array<int32>& block_ptrs)
{
    const int64_t n = mtx->get_size()[0];
    const size_t wsb = gkob200_jacobi_find_blocks_workspace_bytes(n);
    char* ws = static_cast<char*>(scratch().get(wsb + 64));
    int64_t* nb_dev = reinterpret_cast<int64_t*>(ws + ((wsb + 15) / 16) * 16);
    B200(gkob200_jacobi_find_blocks_i32(kStream, n, mtx->get_const_row_ptrs(), mtx->get_const_col_idxs(),
                                        (int32_t)max_block_size, nb_dev, block_ptrs.get_data(), ws, wsb));
    int64_t nb = 0;
    if (cudaMemcpy(&nb, nb_dev, sizeof(nb), cudaMemcpyDeviceToHost) != cudaSuccess)
        throw CudaError(__FILE__, __LINE__, "cudaMemcpy", cudaGetLastError());
    num_blocks = static_cast<size_type>(nb);
}
template <>
void find_blocks<double, int32>(Exec, const matrix::Csr<double, int32>* m, uint32 mbs, size_type& nb, array<int32>& bp)
{
    find_blocks_i32<double>(m, mbs, nb, bp);
}
template <>
void find_blocks<float, int32>(Exec, const matrix::Csr<float, int32>* m, uint32 mbs, size_type& nb, array<int32>& bp)
{
    find_blocks_i32<float>(m, mbs, nb, bp);
}
template <typename V, typename I>
void generate(Exec, const matrix::Csr<V, I>* mtx, size_type num_blocks, uint32, remove_complex<V>,
              const preconditioner::block_interleaved_storage_scheme<I>& scheme, array<remove_complex<V>>&,
              array<precision_reduction>&, const array<I>& block_ptrs, array<V>& blocks)
{
    B200(t::jacobi_block_generate(V{}, kStream, (int64_t)mtx->get_size()[0], mtx->get_const_row_ptrs(),
                                  mtx->get_const_col_idxs(), mtx->get_const_values(), (int64_t)num_blocks,
                                  block_ptrs.get_const_data(), (int64_t)scheme.block_offset, (int64_t)scheme.group_offset,
                                  (int)scheme.group_power, blocks.get_data()));
}
template <typename V, typename I>
void simple_apply(Exec, size_type num_blocks, uint32, const preconditioner::block_interleaved_storage_scheme<I>& scheme,
                  const array<precision_reduction>&, const array<I>& block_ptrs, const array<V>& blocks, const D<V>* b,
                  D<V>* x)
{
    B200(t::jacobi_block_simple_apply(V{}, kStream, (int64_t)num_blocks, block_ptrs.get_const_data(),
                                      blocks.get_const_data(), (int64_t)scheme.block_offset, (int64_t)scheme.group_offset,
                                      (int)scheme.group_power, (int64_t)x->get_size()[0], (int64_t)x->get_size()[1],
                                      b->get_const_values(), (int64_t)b->get_stride(), x->get_values(),
                                      (int64_t)x->get_stride()));
}
template <typename V, typename I>
void apply(Exec, size_type num_blocks, uint32, const preconditioner::block_interleaved_storage_scheme<I>& scheme,
           const array<precision_reduction>&, const array<I>& block_ptrs, const array<V>& blocks, const D<V>* alpha,
           const D<V>* b, const D<V>* beta, D<V>* x)
{
    B200(t::jacobi_block_apply(V{}, kStream, (int64_t)num_blocks, block_ptrs.get_const_data(), blocks.get_const_data(),
                               (int64_t)scheme.block_offset, (int64_t)scheme.group_offset, (int)scheme.group_power,
                               (int64_t)x->get_size()[0], (int64_t)x->get_size()[1], alpha->get_const_values(),
                               b->get_const_values(), (int64_t)b->get_stride(), beta->get_const_values(), x->get_values(),
                               (int64_t)x->get_stride()));
}
template <typename V, typename I>
void transpose_jacobi(Exec, size_type num_blocks, uint32, const array<precision_reduction>&, const array<I>& block_ptrs,
                      const array<V>& blocks, const preconditioner::block_interleaved_storage_scheme<I>& scheme,
                      array<V>& out_blocks)
{
    B200(t::jacobi_block_transpose(V{}, kStream, (int64_t)num_blocks, block_ptrs.get_const_data(), blocks.get_const_data(),
                                   (int64_t)scheme.block_offset, (int64_t)scheme.group_offset, (int)scheme.group_power,
                                   out_blocks.get_data()));
}
template <typename V, typename I>
void conj_transpose_jacobi(Exec exec, size_type num_blocks, uint32 mbs, const array<precision_reduction>& prec,
                           const array<I>& block_ptrs, const array<V>& blocks,
                           const preconditioner::block_interleaved_storage_scheme<I>& scheme, array<V>& out_blocks)
{
    transpose_jacobi<V, I>(exec, num_blocks, mbs, prec, block_ptrs, blocks, scheme, out_blocks);   // real value types
}
void initialize_precisions(Exec exec, const array<precision_reduction>& source, array<precision_reduction>& precisions)
{
    // Blocks are stored in full precision only on this path.  A request for reduced or
    // autodetected storage (adaptive-precision Jacobi, core/preconditioner/jacobi_utils.hpp:45-70)
    // is refused loudly: storing full-precision blocks under metadata that says "reduced" would
    // make every later decode / convert / transpose of the preconditioner wrong.
    const auto n = precisions.get_num_elems();
    const auto m = source.get_num_elems();
    array<precision_reduction> host_src(exec->get_master(), source);
    for (size_type i = 0; i < m; ++i)
        if (!(host_src.get_const_data()[i] == precision_reduction(0, 0)))
            throw NotSupported(__FILE__, __LINE__, __func__, "adaptive-precision block-Jacobi storage");
    array<precision_reduction> host_dst(exec->get_master(), n);
    for (size_type i = 0; i < n; ++i) host_dst.get_data()[i] = precision_reduction(0, 0);
    precisions = host_dst;
}
#define INST_JAC(V)                                                                                             \
    template void invert_diagonal<V>(Exec, const array<V>&, array<V>&);                                          \
    template void simple_scalar_apply<V>(Exec, const array<V>&, const D<V>*, D<V>*);                             \
    template void scalar_apply<V>(Exec, const array<V>&, const D<V>*, const D<V>*, const D<V>*, D<V>*);          \
    template void generate<V, int32>(Exec, const matrix::Csr<V, int32>*, size_type, uint32, V,                   \
                                     const preconditioner::block_interleaved_storage_scheme<int32>&, array<V>&,  \
                                     array<precision_reduction>&, const array<int32>&, array<V>&);               \
    template void simple_apply<V, int32>(Exec, size_type, uint32,                                                \
                                         const preconditioner::block_interleaved_storage_scheme<int32>&,        \
                                         const array<precision_reduction>&, const array<int32>&, const array<V>&, \
                                         const D<V>*, D<V>*);                                                    \
    template void apply<V, int32>(Exec, size_type, uint32, const preconditioner::block_interleaved_storage_scheme<int32>&, \
                                  const array<precision_reduction>&, const array<int32>&, const array<V>&, const D<V>*,  \
                                  const D<V>*, const D<V>*, D<V>*);
INST_JAC(double)
INST_JAC(float)
#define INST_JACT(V)                                                                                               \
    template void transpose_jacobi<V, int32>(Exec, size_type, uint32, const array<precision_reduction>&,            \
                                             const array<int32>&, const array<V>&,                                  \
                                             const preconditioner::block_interleaved_storage_scheme<int32>&,        \
                                             array<V>&);                                                            \
    template void conj_transpose_jacobi<V, int32>(Exec, size_type, uint32, const array<precision_reduction>&,       \
                                                  const array<int32>&, const array<V>&,                             \
                                                  const preconditioner::block_interleaved_storage_scheme<int32>&,   \
                                                  array<V>&);
INST_JACT(double)
INST_JACT(float)


}  // namespace jacobi

}  // namespace cuda
}  // namespace kernels
}  // namespace gko
