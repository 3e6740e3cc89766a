// matrix_apply.cu — format dispatch behind gkob200_matrix (the LinOp::apply of the
// descriptor) and the C entry points of the solver objects.
#include "solver_common.cuh"

namespace gkob200 {

bool matrix_apply_fuses_dot(const gkob200_matrix& A, int64_t nrhs)
{
    if (nrhs != 1) return false;
    if (A.format == GKOB200_FMT_CSR) {
        int strategy = A.csr_strategy;
        if (strategy == GKOB200_CSR_AUTO)
            strategy = A.csr_max_block_nnz > 0
                           ? gkob200_csr_pick_strategy(A.n_rows, A.nnz, -1, A.csr_max_block_nnz)
                           : GKOB200_CSR_MERGE_PATH;
        return strategy == GKOB200_CSR_CLASSICAL;
    }
    // a hybrid matrix whose COO part is empty (strategy automatic on a regular matrix) is its ELL part
    if (A.format == GKOB200_FMT_HYBRID) return A.coo_nnz == 0;
    return A.format == GKOB200_FMT_ELL || A.format == GKOB200_FMT_SELLP || A.format == GKOB200_FMT_CSR_ROWS;
}

template <typename V>
int matrix_apply(cudaStream_t s, const gkob200_matrix& A, const V* b, int64_t b_stride, int64_t nrhs,
                 const V* alpha, const V* beta, V* c, int64_t c_stride, const SpmvFusion<V>* fusion)
{
    const int vt = sizeof(V) == 8 ? GKOB200_F64 : GKOB200_F32;
    if (A.value_type != vt) return GKOB200_EINVAL;
    SpmvFusion<V> fu;
    const SpmvFusion<V>* fp = nullptr;
    if (fusion) {
        fu = *fusion;
        if (!matrix_apply_fuses_dot(A, nrhs)) {
            fu.w = nullptr;
            fu.out = nullptr;
            fu.out_sq = nullptr;
        }
        fp = &fu;
    }
    // a halo exchange attached to the launch exists in one kernel only; dropping it silently
    // would drop the non-local part of a distributed matrix
    if (fusion && fusion->halo && (A.format != GKOB200_FMT_CSR || nrhs != 1)) return GKOB200_EUNSUPPORTED;
    switch (A.format) {
    case GKOB200_FMT_CSR: {
        int strategy = A.csr_strategy;
        if (strategy == GKOB200_CSR_AUTO)
            strategy = A.csr_max_block_nnz > 0
                           ? gkob200_csr_pick_strategy(A.n_rows, A.nnz, -1, A.csr_max_block_nnz)
                           : GKOB200_CSR_MERGE_PATH;
        if (fusion && fusion->halo && strategy != GKOB200_CSR_CLASSICAL) return GKOB200_EUNSUPPORTED;
        if ((strategy == GKOB200_CSR_MERGE_PATH || strategy == GKOB200_CSR_MERGE_PATH_PLANNED) && fp) {
            // the merge-path kernel has no skip/dot fusion; its extra work after the
            // solver stopped only touches solver workspace
            fp = nullptr;
        }
        if (A.index_type == GKOB200_I32)
            return csr_spmv_launch<V, int32_t>(s, A.n_rows, A.n_cols, A.nnz,
                                               static_cast<const int32_t*>(A.row_ptrs),
                                               static_cast<const int32_t*>(A.col_idxs),
                                               static_cast<const V*>(A.values), b, b_stride, nrhs, alpha, beta,
                                               c, c_stride, strategy, A.csr_max_block_nnz, A.workspace,
                                               A.workspace_bytes, nrhs == 1 ? fp : nullptr);
        return csr_spmv_launch<V, int64_t>(s, A.n_rows, A.n_cols, A.nnz, static_cast<const int64_t*>(A.row_ptrs),
                                           static_cast<const int64_t*>(A.col_idxs),
                                           static_cast<const V*>(A.values), b, b_stride, nrhs, alpha, beta, c,
                                           c_stride, strategy, A.csr_max_block_nnz, A.workspace,
                                           A.workspace_bytes, nrhs == 1 ? fp : nullptr);
    }
    case GKOB200_FMT_CSR_ROWS:
        if (A.index_type != GKOB200_I32 || !alpha) return GKOB200_EUNSUPPORTED;
        return csr_rows_spmv_launch<V, int32_t>(s, A.n_listed, A.row_list, static_cast<const int32_t*>(A.row_ptrs),
                                                static_cast<const int32_t*>(A.col_idxs),
                                                static_cast<const V*>(A.values), b, b_stride, nrhs, alpha, beta, c,
                                                c_stride, fusion);
    case GKOB200_FMT_ELL:
        if (A.index_type == GKOB200_I32)
            return ell_spmv_launch<V, int32_t>(s, A.n_rows, A.ell_stride, A.ell_width,
                                               static_cast<const int32_t*>(A.ell_col_idxs),
                                               static_cast<const V*>(A.ell_values), b, b_stride, nrhs, alpha, beta, c,
                                               c_stride, nrhs == 1 ? fp : nullptr);
        return ell_spmv_launch<V, int64_t>(s, A.n_rows, A.ell_stride, A.ell_width,
                                           static_cast<const int64_t*>(A.ell_col_idxs),
                                           static_cast<const V*>(A.ell_values), b, b_stride, nrhs, alpha, beta, c,
                                           c_stride, nrhs == 1 ? fp : nullptr);
    case GKOB200_FMT_SELLP:
        if (nrhs == 1 && A.sellp_max_slice_len > 0) {
            int rc = 0;
            if (A.index_type == GKOB200_I32)
                rc = sellp_spmv_tma_launch<V, int32_t>(s, A.n_rows, A.slice_size, A.slice_sets, A.sellp_max_slice_len,
                                                       A.sellp_total_cols, static_cast<const int32_t*>(A.col_idxs),
                                                       static_cast<const V*>(A.values), b, b_stride, alpha, beta, c,
                                                       c_stride, fp);
            else
                rc = sellp_spmv_tma_launch<V, int64_t>(s, A.n_rows, A.slice_size, A.slice_sets, A.sellp_max_slice_len,
                                                       A.sellp_total_cols, static_cast<const int64_t*>(A.col_idxs),
                                                       static_cast<const V*>(A.values), b, b_stride, alpha, beta, c,
                                                       c_stride, fp);
            if (rc != 0) return rc == 1 ? 0 : rc;
        }
        if (A.index_type == GKOB200_I32)
            return sellp_spmv_launch<V, int32_t>(s, A.n_rows, A.slice_size, A.slice_sets, A.slice_lengths,
                                                 static_cast<const int32_t*>(A.col_idxs),
                                                 static_cast<const V*>(A.values), b, b_stride, nrhs, alpha, beta, c,
                                                 c_stride, nrhs == 1 ? fp : nullptr);
        return sellp_spmv_launch<V, int64_t>(s, A.n_rows, A.slice_size, A.slice_sets, A.slice_lengths,
                                             static_cast<const int64_t*>(A.col_idxs), static_cast<const V*>(A.values),
                                             b, b_stride, nrhs, alpha, beta, c, c_stride, nrhs == 1 ? fp : nullptr);
    case GKOB200_FMT_COO:
        // (no skip/dot fusion: extra work after the solver stopped only touches workspace)
        if (A.index_type == GKOB200_I32)
            return coo_spmv_launch<V, int32_t>(s, A.n_rows, A.nnz, static_cast<const int32_t*>(A.row_ptrs),
                                               static_cast<const int32_t*>(A.col_idxs),
                                               static_cast<const V*>(A.values), b, b_stride, nrhs, alpha, beta, false,
                                               c, c_stride, A.workspace, A.workspace_bytes);
        return coo_spmv_launch<V, int64_t>(s, A.n_rows, A.nnz, static_cast<const int64_t*>(A.row_ptrs),
                                           static_cast<const int64_t*>(A.col_idxs), static_cast<const V*>(A.values), b,
                                           b_stride, nrhs, alpha, beta, false, c, c_stride, A.workspace,
                                           A.workspace_bytes);
    case GKOB200_FMT_HYBRID: {
        // ELL part apply, then COO part apply2 [ref: core/matrix/hybrid.cpp:133-160]
        int rc;
        if (A.index_type == GKOB200_I32) {
            rc = ell_spmv_launch<V, int32_t>(s, A.n_rows, A.ell_stride, A.ell_width,
                                             static_cast<const int32_t*>(A.ell_col_idxs),
                                             static_cast<const V*>(A.ell_values), b, b_stride, nrhs, alpha, beta, c,
                                             c_stride, nrhs == 1 && A.coo_nnz == 0 ? fp : nullptr);
            if (rc || A.coo_nnz == 0) return rc;
            return coo_spmv_launch<V, int32_t>(s, A.n_rows, A.coo_nnz, static_cast<const int32_t*>(A.coo_row_idxs),
                                               static_cast<const int32_t*>(A.coo_col_idxs),
                                               static_cast<const V*>(A.coo_values), b, b_stride, nrhs, alpha, nullptr,
                                               true, c, c_stride, A.workspace, A.workspace_bytes);
        }
        rc = ell_spmv_launch<V, int64_t>(s, A.n_rows, A.ell_stride, A.ell_width,
                                         static_cast<const int64_t*>(A.ell_col_idxs),
                                         static_cast<const V*>(A.ell_values), b, b_stride, nrhs, alpha, beta, c,
                                         c_stride, nrhs == 1 && A.coo_nnz == 0 ? fp : nullptr);
        if (rc || A.coo_nnz == 0) return rc;
        return coo_spmv_launch<V, int64_t>(s, A.n_rows, A.coo_nnz, static_cast<const int64_t*>(A.coo_row_idxs),
                                           static_cast<const int64_t*>(A.coo_col_idxs),
                                           static_cast<const V*>(A.coo_values), b, b_stride, nrhs, alpha, nullptr,
                                           true, c, c_stride, A.workspace, A.workspace_bytes);
    }
    default:
        return GKOB200_EUNSUPPORTED;
    }
}

template int matrix_apply<double>(cudaStream_t, const gkob200_matrix&, const double*, int64_t, int64_t,
                                  const double*, const double*, double*, int64_t, const SpmvFusion<double>*);
template int matrix_apply<float>(cudaStream_t, const gkob200_matrix&, const float*, int64_t, int64_t, const float*,
                                 const float*, float*, int64_t, const SpmvFusion<float>*);

}  // namespace gkob200

using namespace gkob200;

extern "C" {

int gkob200_matrix_apply(void* stream, const gkob200_matrix* A, const void* b, int64_t b_stride, int64_t nrhs,
                         const void* alpha, const void* beta, void* c, int64_t c_stride)
{
    if (!A) return GKOB200_EINVAL;
    if (A->value_type == GKOB200_F64)
        return matrix_apply<double>(as_stream(stream), *A, static_cast<const double*>(b), b_stride, nrhs,
                                    static_cast<const double*>(alpha), static_cast<const double*>(beta),
                                    static_cast<double*>(c), c_stride, nullptr);
    if (A->value_type == GKOB200_F32)
        return matrix_apply<float>(as_stream(stream), *A, static_cast<const float*>(b), b_stride, nrhs,
                                   static_cast<const float*>(alpha), static_cast<const float*>(beta),
                                   static_cast<float*>(c), c_stride, nullptr);
    return GKOB200_EINVAL;
}

int gkob200_solver_create(int kind, const gkob200_matrix* A, const gkob200_precond* M, const gkob200_stop* stop,
                          int64_t nrhs, int64_t krylov_dim, gkob200_solver** out)
{
    if (!A || !stop || !out || nrhs < 1) return GKOB200_EINVAL;
    *out = nullptr;
    // iteration limits are clamped so that `max_iters + chunk` style arithmetic cannot overflow
    gkob200_stop clamped = *stop;
    if (clamped.max_iters < 0) clamped.max_iters = 0;
    if (clamped.max_iters > (int64_t(1) << 40)) clamped.max_iters = int64_t(1) << 40;
    stop = &clamped;
    int rc = GKOB200_EUNSUPPORTED;
    gkob200_solver* s = nullptr;
    if (kind == GKOB200_SOLVER_CG) {
        s = A->value_type == GKOB200_F64 ? make_cg_f64(A, M, stop, nrhs, &rc) : make_cg_f32(A, M, stop, nrhs, &rc);
    }
    else if (kind == GKOB200_SOLVER_BICGSTAB) {
        s = A->value_type == GKOB200_F64 ? make_bicgstab_f64(A, M, stop, nrhs, &rc) : make_bicgstab_f32(A, M, stop, nrhs, &rc);
    } else if (kind == GKOB200_SOLVER_FCG) {
        s = make_fcg(A, M, stop, nrhs, &rc);
    } else if (kind == GKOB200_SOLVER_CGS) {
        s = make_cgs(A, M, stop, nrhs, &rc);
    } else if (kind == GKOB200_SOLVER_GMRES) {
        s = A->value_type == GKOB200_F64 ? make_gmres_f64(A, M, stop, nrhs, krylov_dim, &rc)
                                         : make_gmres_f32(A, M, stop, nrhs, krylov_dim, &rc);
    }
    if (!s) return rc ? rc : GKOB200_EUNSUPPORTED;
    *out = s;
    return 0;
}

int gkob200_solver_destroy(gkob200_solver* s)
{
    delete s;
    return 0;
}

int gkob200_solver_apply(gkob200_solver* s, void* stream, const void* b, int64_t b_stride, void* x, int64_t x_stride)
{
    if (!s) return GKOB200_EINVAL;
    return s->apply(as_stream(stream), b, b_stride, x, x_stride);
}

int gkob200_solver_apply_host(gkob200_solver* s, void* stream, const void* b_host, void* x_host)
{
    if (!s || !b_host || !x_host) return GKOB200_EINVAL;
    return s->apply_host(as_stream(stream), b_host, x_host);
}

int64_t gkob200_solver_num_iterations(const gkob200_solver* s) { return s ? s->num_iterations : -1; }

int gkob200_solver_stop_status(const gkob200_solver* s, uint8_t* out)
{
    if (!s || !out) return GKOB200_EINVAL;
    for (size_t i = 0; i < s->stop_status_host.size(); ++i) out[i] = s->stop_status_host[i];
    return 0;
}

int64_t gkob200_solver_residual_history(const gkob200_solver* s, double* hist, int64_t cap)
{
    if (!s || (!hist && cap > 0)) return GKOB200_EINVAL;
    const int64_t m = static_cast<int64_t>(s->residual_history.size()) < cap
                          ? static_cast<int64_t>(s->residual_history.size())
                          : cap;
    for (int64_t i = 0; i < m; ++i) hist[i] = s->residual_history[i];
    return m;
}

int64_t gkob200_solver_launch_count(const gkob200_solver* s) { return s ? s->launch_count : -1; }

}  // extern "C"
