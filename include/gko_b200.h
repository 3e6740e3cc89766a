/* gko_b200.h — C-ABI of libgko_b200.so: Blackwell (sm_100a) kernels for Ginkgo's
 * sparse-solve hot path.
 *
 * This is the drop-in boundary.  Every entry point below replaces one kernel of
 * the reference's `gko::kernels::cuda::<area>::<kernel>` namespace (the functions
 * `libginkgo_cuda.so` exports and `Executor::run` dispatches to,
 * include/ginkgo/core/base/executor.hpp:456-510), or one host-side solver loop
 * (core/solver/{cg,bicgstab,gmres}.cpp) that we re-issue as a fused, graph-captured
 * sequence.  The reference declaration each function stands in for is cited as
 * `[ref: file:line]`, paths relative to the Ginkgo 1.5.0 tree.  INTEGRATION.md shows
 * the C++ shim a Ginkgo maintainer adds to bind them.
 *
 * Conventions
 *  - plain pointers and sizes only; no C++/torch types cross this boundary.
 *  - all array pointers are DEVICE pointers unless the parameter name ends in
 *    `_host`.  Scalars alpha/beta/rho/... are device-resident (Ginkgo passes them as
 *    1x1 / 1xk Dense on the executor).
 *  - `stream` is a cudaStream_t passed as void*.  Calls enqueue work and return;
 *    they never synchronise, never allocate and never free (the `*_create`,
 *    `*_destroy`, `*_generate` setup calls and functions documented "blocking" excepted).
 *  - return value: 0 on success, a positive cudaError_t on CUDA failure, or one of
 *    the negative GKOB200_E* codes.  Empty inputs are successful no-ops
 *    [ref: cuda/matrix/csr_kernels.cu:435-436].
 *  - dense vectors/multivectors are row-major, element (i,j) at v[i*stride + j]
 *    [ref: include/ginkgo/core/matrix/dense.hpp:685,1173].
 *  - suffixes: _f64/_f32 = ValueType double/float, _i32/_i64 = IndexType.
 *  - there is NO CPU fallback anywhere in this library.
 */
#ifndef GKO_B200_H_
#define GKO_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define GKOB200_EINVAL (-1)     /* bad argument (null pointer, negative size, ...) */
#define GKOB200_EUNSUPPORTED (-2) /* combination not implemented on this path */
#define GKOB200_EWORKSPACE (-3) /* caller-provided workspace too small */
#define GKOB200_ENOTCONVERGED (-4)

/* Bytes of zero-initialised device scratch every reducing kernel needs
 * (ticket + per-block partial sums).  Initialise once with
 * gkob200_reduce_ws_init; kernels leave it re-armed. */
#define GKOB200_REDUCE_WS_BYTES (1024 + (148 * 16 + 256) * 8 * 8)

int gkob200_version(void);
/* Number of SMs of the current device (blocking, cached). */
int gkob200_sm_count(void);
int gkob200_reduce_ws_init(void* stream, void* ws);
/* components::fill_array for elements of 1/2/4/8 bytes; the value is read from the HOST
 * [ref: core/components/fill_array_kernels.hpp] */
int gkob200_fill_array(void* stream, void* data, int64_t n, int elem_bytes, const void* value_host);

/* ------------------------------------------------------------------------- *
 * CSR SpMV / SpMM
 * [ref: core/matrix/csr_kernels.hpp:59-70  csr::spmv, csr::advanced_spmv;
 *       oracle reference/matrix/csr_kernels.cpp:75-129;
 *       replaced cuda/matrix/csr_kernels.cu:430-542]
 * c = A b                      (alpha == NULL && beta == NULL)
 * c = alpha * A b + beta * c   (both non-NULL, device scalars)
 * strategy: which kernel runs (the reference switches on the strategy NAME per call,
 * cuda/matrix/csr_kernels.cu:431-479; here the choice is made once from row-length
 * statistics, see gkob200_csr_row_stats_* / gkob200_csr_pick_strategy).
 * ------------------------------------------------------------------------- */
enum gkob200_csr_strategy {
    GKOB200_CSR_CLASSICAL = 0,  /* row-block kernel: coalesced streams staged in shared
                                   memory, one thread per row, oracle summation order
                                   (bit-identical to the reference executor)          */
    GKOB200_CSR_MERGE_PATH = 1, /* load-balanced merge-path kernel (skewed rows)      */
    GKOB200_CSR_AUTO = 2,       /* = pick_strategy() when stats are given, else MERGE  */
    GKOB200_CSR_MERGE_PATH_PLANNED = 3 /* merge-path with the per-tile split rows precomputed in
                                          `workspace` by gkob200_csr_merge_plan_* (the counterpart of
                                          the srow array Csr::make_srow() fills for load_balance,
                                          include/ginkgo/core/matrix/csr.hpp:421-511)           */
};

/* Row-length statistics of a CSR matrix (device kernel, result on device):
 * stats[0] = max row nnz, stats[1] = max nnz of any block of 128 consecutive rows,
 * stats[2] = number of empty rows, stats[3] = nnz.  int64 each.
 * [ref: include/ginkgo/core/matrix/csr.hpp:242-262 classical::process computes the max
 * row length on the host after a D2H copy of row_ptrs] */
int gkob200_csr_row_stats_i32(void* stream, int64_t n_rows, const int32_t* row_ptrs, int64_t* stats);
int gkob200_csr_row_stats_i64(void* stream, int64_t n_rows, const int64_t* row_ptrs, int64_t* stats);
/* Host-side decision from those statistics (pure function, no CUDA). */
int gkob200_csr_pick_strategy(int64_t n_rows, int64_t nnz, int64_t max_row_nnz, int64_t max_block_nnz);

#define GKOB200_DECL_CSR_SPMV(V, VT, I, IT)                                                   \
    int gkob200_csr_spmv_##V##_##I(void* stream, int64_t n_rows, int64_t n_cols, int64_t nnz, \
                                   const IT* row_ptrs, const IT* col_idxs, const VT* values,  \
                                   const VT* b, int64_t b_stride, int64_t nrhs,               \
                                   const VT* alpha, const VT* beta, VT* c, int64_t c_stride,  \
                                   int strategy, int64_t max_block_nnz, void* workspace,      \
                                   size_t workspace_bytes);
GKOB200_DECL_CSR_SPMV(f64, double, i32, int32_t)
GKOB200_DECL_CSR_SPMV(f32, float, i32, int32_t)
GKOB200_DECL_CSR_SPMV(f64, double, i64, int64_t)
GKOB200_DECL_CSR_SPMV(f32, float, i64, int64_t)
/* Workspace the merge-path kernel needs for its per-tile carries and plan (bytes). */
size_t gkob200_csr_spmv_workspace_bytes(int64_t n_rows, int64_t nnz, int64_t nrhs, int value_bytes);
/* Once per matrix: store the merge-path split row of every tile in `workspace`
 * (sized with value_bytes = 8); enables strategy GKOB200_CSR_MERGE_PATH_PLANNED. */
int gkob200_csr_merge_plan_i32(void* stream, int64_t n_rows, int64_t nnz, const int32_t* row_ptrs, void* workspace,
                               size_t workspace_bytes);
int gkob200_csr_merge_plan_i64(void* stream, int64_t n_rows, int64_t nnz, const int64_t* row_ptrs, void* workspace,
                               size_t workspace_bytes);

/* ------------------------------------------------------------------------- *
 * ELL / SELL-P / COO SpMV and SpMM
 * [ref: core/matrix/ell_kernels.hpp:52-66, sellp_kernels.hpp:52-63, coo_kernels.hpp:53-76;
 *  oracle reference/matrix/{ell,sellp,coo}_kernels.cpp; replaced
 *  cuda/matrix/{ell,sellp,coo}_kernels.cu]
 * ELL: col_idxs/values[stride*width], entry (row,i) at row + i*stride, padding col -1.
 * SELL-P: entry (row,i) of slice s at (slice_sets[s]+i)*slice_size + row%slice_size;
 *         slice_sets / slice_lengths are size_type (uint64) like the reference's.
 * COO: row-sorted triplets; coo_spmv2 ACCUMULATES (c += [alpha] A b), the building block
 *      of Hybrid::apply [ref: core/matrix/hybrid.cpp:133-160].  workspace: carries.
 * alpha/beta: both NULL (simple apply) or both set (coo_spmv2: alpha alone, nullable).
 * ------------------------------------------------------------------------- */
size_t gkob200_coo_spmv_workspace_bytes(int64_t nnz, int value_bytes);
#define GKOB200_DECL_FMT(V, VT, I, IT)                                                                         \
    int gkob200_ell_spmv_##V##_##I(void* stream, int64_t n_rows, int64_t n_cols, int64_t stride, int64_t width, \
                                   const IT* col_idxs, const VT* values, const VT* b, int64_t b_stride,         \
                                   int64_t nrhs, const VT* alpha, const VT* beta, VT* c, int64_t c_stride);     \
    int gkob200_sellp_spmv_##V##_##I(void* stream, int64_t n_rows, int64_t n_cols, int64_t slice_size,          \
                                     const uint64_t* slice_sets, const uint64_t* slice_lengths,                 \
                                     const IT* col_idxs, const VT* values, const VT* b, int64_t b_stride,       \
                                     int64_t nrhs, const VT* alpha, const VT* beta, VT* c, int64_t c_stride);   \
    int gkob200_coo_spmv_##V##_##I(void* stream, int64_t n_rows, int64_t n_cols, int64_t nnz,                   \
                                   const IT* row_idxs, const IT* col_idxs, const VT* values, const VT* b,       \
                                   int64_t b_stride, int64_t nrhs, const VT* alpha, const VT* beta, VT* c,      \
                                   int64_t c_stride, void* workspace, size_t workspace_bytes);                  \
    int gkob200_coo_spmv2_##V##_##I(void* stream, int64_t n_rows, int64_t n_cols, int64_t nnz,                  \
                                    const IT* row_idxs, const IT* col_idxs, const VT* values, const VT* b,      \
                                    int64_t b_stride, int64_t nrhs, const VT* alpha, VT* c, int64_t c_stride,   \
                                    void* workspace, size_t workspace_bytes);
GKOB200_DECL_FMT(f64, double, i32, int32_t)
GKOB200_DECL_FMT(f32, float, i32, int32_t)
GKOB200_DECL_FMT(f64, double, i64, int64_t)
GKOB200_DECL_FMT(f32, float, i64, int64_t)

/* ------------------------------------------------------------------------- *
 * Integer / index kernels (bit-exact with the reference executor)
 * [ref: core/components/prefix_sum_kernels.hpp:67 (exclusive scan, last entry receives the
 *  total), common/unified/components/format_conversion_kernels.cpp:49-112,
 *  reference/matrix/sellp_kernels.cpp:134-160 compute_slice_sets,
 *  reference/matrix/ell_kernels.cpp:159-168 compute_max_row_nnz,
 *  common/unified/matrix/hybrid_kernels.cpp:51-76 compute_coo_row_ptrs,
 *  common/unified/matrix/csr_kernels.cpp:137-243 convert_to_{sellp,ell,hybrid}]
 * ws: gkob200_prefix_sum_workspace_bytes(n) bytes of scratch.
 * ------------------------------------------------------------------------- */
size_t gkob200_prefix_sum_workspace_bytes(int64_t n);
int gkob200_prefix_sum_i32(void* stream, int32_t* data, int64_t n, void* ws, size_t ws_bytes);
int gkob200_prefix_sum_i64(void* stream, int64_t* data, int64_t n, void* ws, size_t ws_bytes);
int gkob200_prefix_sum_u64(void* stream, uint64_t* data, int64_t n, void* ws, size_t ws_bytes);
int gkob200_hybrid_compute_coo_row_ptrs(void* stream, const uint64_t* row_nnz, int64_t n, uint64_t ell_lim,
                                        int64_t* coo_row_ptrs, void* ws, size_t ws_bytes);
#define GKOB200_DECL_CONV_I(I, IT)                                                                             \
    int gkob200_convert_ptrs_to_idxs_##I(void* stream, const IT* ptrs, int64_t n, IT* idxs);                    \
    int gkob200_convert_idxs_to_ptrs_##I(void* stream, const IT* idxs, int64_t num_idxs, int64_t n, IT* ptrs);  \
    int gkob200_convert_ptrs_to_sizes_##I(void* stream, const IT* ptrs, int64_t n, uint64_t* sizes);            \
    int gkob200_compute_max_row_nnz_##I(void* stream, const IT* row_ptrs, int64_t n, uint64_t* max_nnz);        \
    int gkob200_sellp_compute_slice_sets_##I(void* stream, const IT* row_ptrs, int64_t n, int64_t slice_size,   \
                                             int64_t stride_factor, uint64_t* slice_sets,                      \
                                             uint64_t* slice_lengths, void* ws, size_t ws_bytes);              \
    /* hist[b] = #rows with lo + b*width <= row length < lo + (b+1)*width, b < bins <= 8192; used to take   \
     * the order statistics Hybrid's imbalance_limit strategies need without sorting on the host */          \
    int gkob200_row_len_histogram_##I(void* stream, const IT* row_ptrs, int64_t n, uint64_t lo, uint64_t width, \
                                      int bins, uint64_t* hist);
GKOB200_DECL_CONV_I(i32, int32_t)
GKOB200_DECL_CONV_I(i64, int64_t)
#define GKOB200_DECL_CONV_VI(V, VT, I, IT)                                                                     \
    int gkob200_csr_convert_to_ell_##V##_##I(void* stream, int64_t n, const IT* row_ptrs, const IT* col_idxs,   \
                                             const VT* values, int64_t ell_width, int64_t ell_stride,          \
                                             IT* ell_col_idxs, VT* ell_values);                                \
    int gkob200_csr_convert_to_sellp_##V##_##I(void* stream, int64_t n, const IT* row_ptrs,                     \
                                               const IT* col_idxs, const VT* values, int64_t slice_size,       \
                                               const uint64_t* slice_sets, IT* out_col_idxs, VT* out_values);  \
    int gkob200_csr_convert_to_hybrid_##V##_##I(void* stream, int64_t n, const IT* row_ptrs,                    \
                                                const IT* col_idxs, const VT* values,                          \
                                                const int64_t* coo_row_ptrs, int64_t ell_stride,               \
                                                int64_t ell_width, IT* ell_col_idxs, VT* ell_values,           \
                                                IT* coo_row_idxs, IT* coo_col_idxs, VT* coo_values);
GKOB200_DECL_CONV_VI(f64, double, i32, int32_t)
GKOB200_DECL_CONV_VI(f32, float, i32, int32_t)
GKOB200_DECL_CONV_VI(f64, double, i64, int64_t)
GKOB200_DECL_CONV_VI(f32, float, i64, int64_t)

/* ------------------------------------------------------------------------- *
 * Dense BLAS-1  [ref: core/matrix/dense_kernels.hpp; oracle
 * reference/matrix/dense_kernels.cpp:158-378; replaced
 * common/unified/matrix/dense_kernels.cpp:58-466 + cuda/matrix/dense_kernels.cu:76-149]
 * x, y: n x k row-major.  `alpha` is 1x1 (alpha_cols==1) or 1xk (alpha_cols==k).
 * result: 1 x k on device.  ws: GKOB200_REDUCE_WS_BYTES of initialised scratch.
 * ------------------------------------------------------------------------- */
#define GKOB200_DECL_DENSE(V, VT)                                                                        \
    int gkob200_dense_fill_##V(void* stream, int64_t n, int64_t k, VT* x, int64_t stride, VT value);    \
    int gkob200_dense_copy_##V(void* stream, int64_t n, int64_t k, const VT* x, int64_t xs, VT* y,      \
                               int64_t ys);                                                              \
    int gkob200_dense_scale_##V(void* stream, int64_t n, int64_t k, const VT* alpha, int64_t alpha_cols, \
                                VT* x, int64_t xs);                                                      \
    int gkob200_dense_inv_scale_##V(void* stream, int64_t n, int64_t k, const VT* alpha,                 \
                                    int64_t alpha_cols, VT* x, int64_t xs);                              \
    int gkob200_dense_add_scaled_##V(void* stream, int64_t n, int64_t k, const VT* alpha,                \
                                     int64_t alpha_cols, const VT* x, int64_t xs, VT* y, int64_t ys);    \
    int gkob200_dense_sub_scaled_##V(void* stream, int64_t n, int64_t k, const VT* alpha,                \
                                     int64_t alpha_cols, const VT* x, int64_t xs, VT* y, int64_t ys);    \
    int gkob200_dense_compute_dot_##V(void* stream, int64_t n, int64_t k, const VT* x, int64_t xs,       \
                                      const VT* y, int64_t ys, VT* result, void* ws);                    \
    int gkob200_dense_compute_norm2_##V(void* stream, int64_t n, int64_t k, const VT* x, int64_t xs,     \
                                        VT* result, void* ws);                                           \
    int gkob200_dense_compute_squared_norm2_##V(void* stream, int64_t n, int64_t k, const VT* x,         \
                                                int64_t xs, VT* result, void* ws);                       \
    int gkob200_dense_compute_norm1_##V(void* stream, int64_t n, int64_t k, const VT* x, int64_t xs,     \
                                        VT* result, void* ws);                                           \
    int gkob200_dense_compute_sqrt_##V(void* stream, int64_t k, VT* x);                                  \
    int gkob200_dense_row_gather_##V##_i32(void* stream, int64_t n_out, int64_t k, const int32_t* rows,  \
                                           const VT* src, int64_t ss, VT* dst, int64_t ds);              \
    int gkob200_dense_row_gather_##V##_i64(void* stream, int64_t n_out, int64_t k, const int64_t* rows,  \
                                           const VT* src, int64_t ss, VT* dst, int64_t ds);
GKOB200_DECL_DENSE(f64, double)
GKOB200_DECL_DENSE(f32, float)

/* ------------------------------------------------------------------------- *
 * CG step kernels, 1:1 with the reference kernel set
 * [ref: core/solver/cg_kernels.hpp:55-79; oracle reference/solver/cg_kernels.cpp:56-133;
 *       replaced common/unified/solver/cg_kernels.cpp:51-138]
 * All vectors n x k with the same `stride`; scalars 1 x k; stop_status k bytes.
 * ------------------------------------------------------------------------- */
#define GKOB200_DECL_CG(V, VT)                                                                     \
    int gkob200_cg_initialize_##V(void* stream, int64_t n, int64_t k, const VT* b, int64_t b_stride, \
                                  VT* r, VT* z, VT* p, VT* q, int64_t stride, VT* prev_rho, VT* rho, \
                                  uint8_t* stop_status);                                            \
    int gkob200_cg_step_1_##V(void* stream, int64_t n, int64_t k, VT* p, const VT* z, int64_t stride, \
                              const VT* rho, const VT* prev_rho, const uint8_t* stop_status);       \
    int gkob200_cg_step_2_##V(void* stream, int64_t n, int64_t k, VT* x, int64_t x_stride, VT* r,   \
                              const VT* p, const VT* q, int64_t stride, const VT* beta,             \
                              const VT* rho, const uint8_t* stop_status);
GKOB200_DECL_CG(f64, double)
GKOB200_DECL_CG(f32, float)

/* ------------------------------------------------------------------------- *
 * Stopping criteria  [ref: core/stop/residual_norm_kernels.hpp:49-80; oracle
 * reference/stop/residual_norm_kernels.cpp:58-137; replaced
 * cuda/stop/residual_norm_kernels.cu:61-199 which ends in two blocking 1-byte D2H
 * copies per call].  Here the two booleans are written to `flags` (device or
 * pinned-mapped host memory): flags[0] = all_converged, flags[1] = one_changed.
 * ------------------------------------------------------------------------- */
#define GKOB200_DECL_STOP(V, VT)                                                                   \
    int gkob200_residual_norm_##V(void* stream, int64_t k, const VT* tau, const VT* orig_tau,      \
                                  VT rel_residual_goal, uint8_t stopping_id, int set_finalized,    \
                                  uint8_t* stop_status, uint8_t* flags);                            \
    int gkob200_implicit_residual_norm_##V(void* stream, int64_t k, const VT* tau,                 \
                                           const VT* orig_tau, VT rel_residual_goal,               \
                                           uint8_t stopping_id, int set_finalized,                 \
                                           uint8_t* stop_status, uint8_t* flags);
GKOB200_DECL_STOP(f64, double)
GKOB200_DECL_STOP(f32, float)
/* [ref: core/stop/criterion_kernels.hpp set_all_statuses; cuda/stop/criterion_kernels.cu:56-83] */
int gkob200_set_all_statuses(void* stream, int64_t k, uint8_t stopping_id, int set_finalized,
                             uint8_t* stop_status);

/* ------------------------------------------------------------------------- *
 * Scalar Jacobi  [ref: core/preconditioner/jacobi_kernels.hpp:73-107; oracle
 * reference/preconditioner/jacobi_kernels.cpp:565-625; replaced
 * common/unified/preconditioner/jacobi_kernels.cpp] and csr::extract_diagonal
 * [ref: core/matrix/csr_kernels.hpp extract_diagonal; reference/matrix/csr_kernels.cpp]
 * ------------------------------------------------------------------------- */
#define GKOB200_DECL_JACOBI_SCALAR(V, VT)                                                          \
    int gkob200_csr_extract_diagonal_##V##_i32(void* stream, int64_t n_rows, int64_t n_cols,       \
                                               const int32_t* row_ptrs, const int32_t* col_idxs,   \
                                               const VT* values, VT* diag);                        \
    int gkob200_jacobi_invert_diagonal_##V(void* stream, int64_t n, const VT* diag, VT* inv_diag); \
    int gkob200_jacobi_simple_scalar_apply_##V(void* stream, int64_t n, int64_t k, const VT* inv_diag, \
                                               const VT* b, int64_t b_stride, VT* x, int64_t x_stride); \
    int gkob200_jacobi_scalar_apply_##V(void* stream, int64_t n, int64_t k, const VT* inv_diag,    \
                                        const VT* alpha, const VT* b, int64_t b_stride,            \
                                        const VT* beta, VT* x, int64_t x_stride);
GKOB200_DECL_JACOBI_SCALAR(f64, double)
GKOB200_DECL_JACOBI_SCALAR(f32, float)

/* ------------------------------------------------------------------------- *
 * BiCGSTAB step kernels  [ref: core/solver/bicgstab_kernels.hpp:55-104; oracle
 * reference/solver/bicgstab_kernels.cpp:53-213; replaced
 * common/unified/solver/bicgstab_kernels.cpp:53-213].  step_2 also writes alpha,
 * step_3 writes omega (from row 0, like the reference's device kernels).
 * GMRES kernels  [ref: core/solver/{gmres,common_gmres}_kernels.hpp; oracle
 * reference/solver/{gmres,common_gmres}_kernels.cpp].  krylov_bases is
 * (krylov_dim+1)*n x k, hessenberg (krylov_dim+1) x krylov_dim*k (entry (i, iter*k + rhs)),
 * hessenberg_iter points at column block `iter`; final_iter_nums are size_type (uint64).
 * ------------------------------------------------------------------------- */
#define GKOB200_DECL_KRYLOV(V, VT)                                                                                \
    /* FCG / CGS step kernels [ref: core/solver/{fcg,cgs}_kernels.hpp; oracle                      \
     * reference/solver/fcg_kernels.cpp:50-128, cgs_kernels.cpp:50-167] */                                         \
    int gkob200_fcg_initialize_##V(void* stream, int64_t n, int64_t k, const VT* b, int64_t b_stride, VT* r,       \
                                   VT* z, VT* p, VT* q, VT* t, int64_t stride, VT* prev_rho, VT* rho, VT* rho_t,   \
                                   uint8_t* stop_status);                                                         \
    int gkob200_fcg_step_1_##V(void* stream, int64_t n, int64_t k, VT* p, const VT* z, int64_t stride,             \
                               const VT* rho_t, const VT* prev_rho, const uint8_t* stop_status);                   \
    int gkob200_fcg_step_2_##V(void* stream, int64_t n, int64_t k, VT* x, int64_t x_stride, VT* r, VT* t,          \
                               const VT* p, const VT* q, int64_t stride, const VT* beta, const VT* rho,            \
                               const uint8_t* stop_status);                                                       \
    /* BiCG step kernels [ref: core/solver/bicg_kernels.hpp; replaced                               \
     * common/unified/solver/bicg_kernels.cpp:53-170]: r2/z2/p2/q2 belong to the transposed system; the           \
     * host loop (core/solver/bicg.cpp:160-243) builds A^T with csr::transpose.  solver::Ir needs no kernel of     \
     * its own beyond ir::initialize (a status reset: gkob200_set_all_statuses / fill_array). */                   \
    int gkob200_bicg_initialize_##V(void* stream, int64_t n, int64_t k, const VT* b, int64_t b_stride, VT* r,      \
                                    VT* z, VT* p, VT* q, VT* r2, VT* z2, VT* p2, VT* q2, int64_t stride,           \
                                    VT* prev_rho, VT* rho, uint8_t* stop_status);                                 \
    int gkob200_bicg_step_1_##V(void* stream, int64_t n, int64_t k, VT* p, const VT* z, VT* p2, const VT* z2,      \
                                int64_t stride, const VT* rho, const VT* prev_rho, const uint8_t* stop_status);    \
    int gkob200_bicg_step_2_##V(void* stream, int64_t n, int64_t k, VT* x, int64_t x_stride, VT* r, VT* r2,        \
                                const VT* p, const VT* q, const VT* q2, int64_t stride, const VT* beta,            \
                                const VT* rho, const uint8_t* stop_status);                                       \
    int gkob200_cgs_initialize_##V(void* stream, int64_t n, int64_t k, const VT* b, int64_t b_stride, VT* r,       \
                                   VT* r_tld, VT* p, VT* q, VT* u, VT* u_hat, VT* v_hat, VT* t, int64_t stride,    \
                                   VT* alpha, VT* beta, VT* gamma, VT* rho_prev, VT* rho, uint8_t* stop_status);   \
    int gkob200_cgs_step_1_##V(void* stream, int64_t n, int64_t k, const VT* r, VT* u, VT* p, const VT* q,         \
                               int64_t stride, VT* beta, const VT* rho, const VT* rho_prev,                        \
                               const uint8_t* stop_status);                                                       \
    int gkob200_cgs_step_2_##V(void* stream, int64_t n, int64_t k, const VT* u, const VT* v_hat, VT* q, VT* t,     \
                               int64_t stride, VT* alpha, const VT* rho, const VT* gamma,                          \
                               const uint8_t* stop_status);                                                       \
    int gkob200_cgs_step_3_##V(void* stream, int64_t n, int64_t k, const VT* t, const VT* u_hat, VT* r,            \
                               int64_t stride, VT* x, int64_t x_stride, const VT* alpha,                           \
                               const uint8_t* stop_status);                                                       \
    int gkob200_bicgstab_initialize_##V(void* stream, int64_t n, int64_t k, const VT* b, int64_t b_stride, VT* r,  \
                                        VT* rr, VT* y, VT* s, VT* t, VT* z, VT* v, VT* p, int64_t stride,          \
                                        VT* prev_rho, VT* rho, VT* alpha, VT* beta, VT* gamma, VT* omega,          \
                                        uint8_t* stop_status);                                                    \
    int gkob200_bicgstab_step_1_##V(void* stream, int64_t n, int64_t k, const VT* r, VT* p, const VT* v,           \
                                    int64_t stride, const VT* rho, const VT* prev_rho, const VT* alpha,            \
                                    const VT* omega, const uint8_t* stop_status);                                 \
    int gkob200_bicgstab_step_2_##V(void* stream, int64_t n, int64_t k, const VT* r, VT* s, const VT* v,           \
                                    int64_t stride, const VT* rho, VT* alpha, const VT* beta,                      \
                                    const uint8_t* stop_status);                                                  \
    int gkob200_bicgstab_step_3_##V(void* stream, int64_t n, int64_t k, VT* x, int64_t x_stride, VT* r,            \
                                    const VT* s, const VT* t, const VT* y, const VT* z, int64_t stride,            \
                                    const VT* alpha, const VT* beta, const VT* gamma, VT* omega,                   \
                                    const uint8_t* stop_status);                                                  \
    int gkob200_bicgstab_finalize_##V(void* stream, int64_t n, int64_t k, VT* x, int64_t x_stride, const VT* y,    \
                                      int64_t stride, const VT* alpha, uint8_t* stop_status);                      \
    int gkob200_gmres_initialize_##V(void* stream, int64_t n, int64_t k, int64_t krylov_dim, const VT* b,          \
                                     int64_t b_stride, VT* residual, int64_t r_stride, VT* givens_sin,             \
                                     VT* givens_cos, uint8_t* stop_status);                                       \
    int gkob200_gmres_restart_##V(void* stream, int64_t n, int64_t k, const VT* residual, int64_t r_stride,        \
                                  const VT* residual_norm, VT* residual_norm_collection, VT* krylov_bases,         \
                                  uint64_t* final_iter_nums);                                                     \
    int gkob200_gmres_multi_axpy_##V(void* stream, int64_t n, int64_t k, const VT* krylov_bases, const VT* y,      \
                                     VT* before_preconditioner, int64_t bp_stride,                                \
                                     const uint64_t* final_iter_nums, uint8_t* stop_status);                      \
    int gkob200_gmres_hessenberg_qr_##V(void* stream, int64_t k, VT* givens_sin, VT* givens_cos,                   \
                                        VT* residual_norm, VT* residual_norm_collection, VT* hessenberg_iter,      \
                                        int64_t hessenberg_stride, int64_t iter, uint64_t* final_iter_nums,        \
                                        const uint8_t* stop_status);                                              \
    int gkob200_gmres_solve_krylov_##V(void* stream, int64_t k, const VT* residual_norm_collection,                \
                                       const VT* hessenberg, int64_t hessenberg_stride, VT* y,                     \
                                       const uint64_t* final_iter_nums, const uint8_t* stop_status);
GKOB200_DECL_KRYLOV(f64, double)
GKOB200_DECL_KRYLOV(f32, float)

/* ------------------------------------------------------------------------- *
 * Block-Jacobi  [ref: core/preconditioner/jacobi_kernels.hpp:50-105; oracle
 * reference/preconditioner/jacobi_kernels.cpp:66-148 (find_blocks), :157-438 (generate:
 * Gauss-Jordan inversion with max-abs column pivoting), :447-562 (apply); replaced
 * cuda/preconditioner/jacobi_*.cu].  Storage = Ginkgo's block_interleaved_storage_scheme
 * (include/ginkgo/core/preconditioner/jacobi.hpp:62-165): block b starts at
 * group_offset*(b >> group_power) + block_offset*(b & (2^group_power - 1)), entry (r,c) at
 * r + c*stride with stride = block_offset << group_power.  No adaptive precision.
 * find_blocks: num_blocks (device, 1 int64) and block_pointers (int32[n+1] capacity).
 * generate: a zero pivot stops the elimination of that block where it is, exactly like the
 * reference executor (invert_block's status is ignored, jacobi_kernels.cpp:375-377).
 * ------------------------------------------------------------------------- */
size_t gkob200_jacobi_find_blocks_workspace_bytes(int64_t n_rows);
int gkob200_jacobi_find_blocks_i32(void* stream, int64_t n_rows, const int32_t* row_ptrs, const int32_t* col_idxs,
                                   int32_t max_block_size, int64_t* num_blocks, int32_t* block_pointers,
                                   void* workspace, size_t workspace_bytes);
#define GKOB200_DECL_JACOBI_BLOCK(V, VT)                                                                          \
    int gkob200_jacobi_block_generate_##V(void* stream, int64_t n_rows, const int32_t* row_ptrs,                   \
                                          const int32_t* col_idxs, const VT* values, int64_t num_blocks,           \
                                          const int32_t* block_pointers, int64_t block_offset,                     \
                                          int64_t group_offset, int group_power, VT* blocks);                      \
    /* out block = transpose of the stored block [ref: jacobi::transpose_jacobi / conj_transpose_jacobi,        \
     * core/preconditioner/jacobi_kernels.hpp:114-134] (what solver::Bicg asks of its preconditioner) */          \
    int gkob200_jacobi_block_transpose_##V(void* stream, int64_t num_blocks, const int32_t* block_pointers,        \
                                           const VT* blocks, int64_t block_offset, int64_t group_offset,           \
                                           int group_power, VT* out_blocks);                                      \
    int gkob200_jacobi_block_simple_apply_##V(void* stream, int64_t num_blocks, const int32_t* block_pointers,     \
                                              const VT* blocks, int64_t block_offset, int64_t group_offset,        \
                                              int group_power, int64_t n, int64_t k, const VT* b, int64_t b_stride, \
                                              VT* x, int64_t x_stride);                                            \
    int gkob200_jacobi_block_apply_##V(void* stream, int64_t num_blocks, const int32_t* block_pointers,            \
                                       const VT* blocks, int64_t block_offset, int64_t group_offset,               \
                                       int group_power, int64_t n, int64_t k, const VT* alpha, const VT* b,        \
                                       int64_t b_stride, const VT* beta, VT* x, int64_t x_stride);
GKOB200_DECL_JACOBI_BLOCK(f64, double)
GKOB200_DECL_JACOBI_BLOCK(f32, float)

/* ------------------------------------------------------------------------- *
 * Synthetic matrix generators (HOST functions writing HOST buffers): the BASELINE
 * shapes, rows [row_begin,row_end) of the global matrix.  kind 0: 2D 5-pt (diag 4),
 * 1: 3D 7-pt (diag 6), 2: 3D 27-pt (diag 26); off-diagonals -1, Dirichlet truncation.
 * [ref: the reference assembles its stencil inputs on the host the same way,
 *  examples/distributed-solver/distributed-solver.cpp:176-186]
 * Suffix: value type, row_ptr type, column type.
 * ------------------------------------------------------------------------- */
int64_t gkob200_gen_stencil_nnz(int kind, int64_t nx, int64_t ny, int64_t nz, int64_t row_begin, int64_t row_end);
#define GKOB200_DECL_GEN(V, VT, P, PT, C, CT)                                                          \
    int gkob200_gen_stencil_csr_##V##_##P##_##C(int kind, int64_t nx, int64_t ny, int64_t nz,          \
                                                int64_t row_begin, int64_t row_end, PT* row_ptrs_host, \
                                                CT* col_idxs_host, VT* values_host);
GKOB200_DECL_GEN(f64, double, i32, int32_t, i32, int32_t)
GKOB200_DECL_GEN(f32, float, i32, int32_t, i32, int32_t)
GKOB200_DECL_GEN(f64, double, i64, int64_t, i64, int64_t)
GKOB200_DECL_GEN(f32, float, i64, int64_t, i64, int64_t)
/* power-law matrix of config 3: returns nnz */
int64_t gkob200_gen_powerlaw_row_ptrs_i64(int64_t n, uint64_t seed, double lmin, double alpha, int64_t lmax,
                                          int64_t* row_ptrs_host);
int gkob200_gen_powerlaw_fill_f64_i32(int64_t n, uint64_t seed, const int64_t* row_ptrs_host,
                                      int32_t* row_ptrs32_host, int32_t* col_idxs_host, double* values_host);

/* ------------------------------------------------------------------------- *
 * Matrix assembly on the device (setup path, not the hot loop)
 * [ref: components::sort_row_major / sum_duplicates / remove_zeros
 *       core/base/device_matrix_data.cpp, reference/base/device_matrix_data_kernels.cpp:82-172;
 *       csr::transpose reference/matrix/csr_kernels.cpp:551-587;
 *       csr::sort_by_column_index reference/matrix/csr_kernels.cpp:969-987]
 * sort_row_major is a STABLE sort by (row, col) (the reference's std::sort leaves the order of
 * duplicates unspecified); sum_duplicates / remove_zeros expect what the reference's callers
 * give them (row-major sorted input for sum_duplicates), write to separate output arrays of
 * capacity nnz and leave the new entry count in *out_nnz (device memory); duplicates are added
 * in input order starting from zero, like the reference loop.  Integer outputs bit-exact,
 * values bit-exact.  Workspaces: gkob200_setup_sort_workspace_bytes for sort_row_major /
 * transpose / sort_by_column_index, gkob200_setup_compact_workspace_bytes for the other two.
 * ------------------------------------------------------------------------- */
size_t gkob200_setup_sort_workspace_bytes(int64_t nnz, int value_bytes, int index_bytes);
size_t gkob200_setup_compact_workspace_bytes(int64_t nnz);
#define GKOB200_DECL_SETUP(V, VT, I, IT)                                                                           \
    int gkob200_coo_sort_row_major_##V##_##I(void* stream, int64_t n_rows, int64_t n_cols, int64_t nnz, IT* rows,   \
                                             IT* cols, VT* vals, void* ws, size_t ws_bytes);                       \
    int gkob200_coo_sum_duplicates_##V##_##I(void* stream, int64_t nnz, const IT* rows, const IT* cols,             \
                                             const VT* vals, IT* out_rows, IT* out_cols, VT* out_vals,              \
                                             int64_t* out_nnz, void* ws, size_t ws_bytes);                         \
    int gkob200_coo_remove_zeros_##V##_##I(void* stream, int64_t nnz, const IT* rows, const IT* cols,               \
                                           const VT* vals, IT* out_rows, IT* out_cols, VT* out_vals,                \
                                           int64_t* out_nnz, void* ws, size_t ws_bytes);                           \
    int gkob200_csr_transpose_##V##_##I(void* stream, int64_t n_rows, int64_t n_cols, int64_t nnz,                  \
                                        const IT* row_ptrs, const IT* col_idxs, const VT* values, IT* out_row_ptrs, \
                                        IT* out_col_idxs, VT* out_values, void* ws, size_t ws_bytes);              \
    int gkob200_csr_sort_by_column_index_##V##_##I(void* stream, int64_t n_rows, int64_t n_cols, int64_t nnz,       \
                                                   const IT* row_ptrs, IT* col_idxs, VT* values, void* ws,          \
                                                   size_t ws_bytes);
GKOB200_DECL_SETUP(f64, double, i32, int32_t)
GKOB200_DECL_SETUP(f32, float, i32, int32_t)
GKOB200_DECL_SETUP(f64, double, i64, int64_t)
GKOB200_DECL_SETUP(f32, float, i64, int64_t)
/* device_matrix_data <-> array<matrix_data_entry> [ref: components::aos_to_soa / soa_to_aos,
 * core/base/device_matrix_data_kernels.hpp:54-62; reference/base/device_matrix_data_kernels.cpp:50-80].
 * `entries`: nnz structs { I row; I column; V value; } with natural alignment, device memory. */
#define GKOB200_DECL_AOS(V, VT, I, IT)                                                                             \
    int gkob200_aos_to_soa_##V##_##I(void* stream, int64_t nnz, const void* entries, IT* rows, IT* cols, VT* vals); \
    int gkob200_soa_to_aos_##V##_##I(void* stream, int64_t nnz, const IT* rows, const IT* cols, const VT* vals,    \
                                     void* entries);
GKOB200_DECL_AOS(f64, double, i32, int32_t)
GKOB200_DECL_AOS(f32, float, i32, int32_t)
GKOB200_DECL_AOS(f64, double, i64, int64_t)
GKOB200_DECL_AOS(f32, float, i64, int64_t)

/* ------------------------------------------------------------------------- *
 * Wire / disk formats of the callers (host code, no GPU needed)
 * [ref: core/base/mtx_io.cpp — read (MatrixMarket text) :77-103, header grammar :690-722,
 *  storage modifiers :293-463, layouts :509-655, GINKGO binary format :776-958,
 *  read_generic_raw :911-925]
 * gkob200_mtx_read_open parses a whole file (text or binary, chosen by its first byte like
 * read_generic_raw) into row-major sorted triplets; _copy_ converts them into caller-owned HOST
 * arrays of nnz entries; complex files are rejected (real value types only), indices that do
 * not fit the requested index type are rejected.  Errors return GKOB200_EINVAL and leave a
 * message in gkob200_mtx_last_error() (thread-local), the reference's stream-error texts.
 * gkob200_mtx_write_*: format 0 = "%%MatrixMarket matrix coordinate real general" (precision <= 0:
 * 6 significant digits, what the reference's ostream default gives), 1 = GINKGO binary,
 * 2 = "%%MatrixMarket matrix array real general" (dense, column-major; what gko::write gives for Dense).
 * ------------------------------------------------------------------------- */
const char* gkob200_mtx_last_error(void);
int gkob200_mtx_read_open(const char* path, void** handle, int64_t* n_rows, int64_t* n_cols, int64_t* nnz);
int gkob200_mtx_read_close(void* handle);
#define GKOB200_DECL_MTX(V, VT, I, IT)                                                                             \
    int gkob200_mtx_read_copy_##V##_##I(void* handle, IT* rows, IT* cols, VT* vals);                                \
    int gkob200_mtx_write_##V##_##I(const char* path, int format, int precision, int64_t n_rows, int64_t n_cols,    \
                                    int64_t nnz, const IT* rows, const IT* cols, const VT* vals);
GKOB200_DECL_MTX(f64, double, i32, int32_t)
GKOB200_DECL_MTX(f32, float, i32, int32_t)
GKOB200_DECL_MTX(f64, double, i64, int64_t)
GKOB200_DECL_MTX(f32, float, i64, int64_t)

#ifdef __cplusplus
} /* extern "C" */
#endif

#include "gko_b200_solver.h"

#endif /* GKO_B200_H_ */
