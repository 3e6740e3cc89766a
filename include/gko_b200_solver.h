/* gko_b200_solver.h — operator descriptors and fused Krylov solver objects.
 * Included by gko_b200.h.
 *
 * A solver object corresponds to what `solver::Cg<V>::build().with_criteria(...)
 * .with_preconditioner(...).on(exec)->generate(A)` returns in the reference
 * [ref: include/ginkgo/core/solver/cg.hpp:84-170, solver_base.hpp:738-779]:
 * it owns its workspace vectors (allocated once at create, reused by every apply —
 * the reference caches them the same way, core/solver/solver_boilerplate.hpp, and its
 * tests assert a second apply allocates nothing, test/solver/solver.cpp:494-513),
 * borrows the matrix / preconditioner arrays described by the descriptors, and its
 * apply() runs the reference's iteration (core/solver/cg.cpp:107-194) as a CUDA-graph
 * of fused kernels with every scalar (rho, beta, residual norm, stopping status,
 * iteration counter) resident on the device.
 */
#ifndef GKO_B200_SOLVER_H_
#define GKO_B200_SOLVER_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

enum gkob200_value_type { GKOB200_F64 = 0, GKOB200_F32 = 1 };
enum gkob200_index_type { GKOB200_I32 = 0, GKOB200_I64 = 1 };
enum gkob200_format {
    GKOB200_FMT_CSR = 0,    /* row_ptrs[n+1], col_idxs[nnz], values[nnz]                           */
    GKOB200_FMT_ELL = 1,    /* col_idxs/values[ell_stride*ell_width] column-major, pad col = -1    */
    GKOB200_FMT_SELLP = 2,  /* slice_sets[ns+1] (uint64), slice_lengths[ns] (uint64), slice_size   */
    GKOB200_FMT_COO = 3,    /* row_idxs in row_ptrs, col_idxs, values, all [nnz], row-sorted       */
    GKOB200_FMT_HYBRID = 4  /* ELL part in the ell_* fields + COO part in the coo_* fields         */
};

/* Borrowed description of a sparse matrix living on the current device
 * [ref: the array members of matrix::Csr/Ell/Sellp/Coo/Hybrid,
 *  include/ginkgo/core/matrix/{csr,ell,sellp,coo,hybrid}.hpp]. */
typedef struct gkob200_matrix {
    int32_t format;      /* gkob200_format */
    int32_t value_type;  /* gkob200_value_type */
    int32_t index_type;  /* gkob200_index_type */
    int32_t csr_strategy; /* gkob200_csr_strategy */
    int64_t n_rows, n_cols, nnz;
    const void* row_ptrs; /* CSR row_ptrs / COO row_idxs */
    const void* col_idxs;
    const void* values;
    int64_t csr_max_block_nnz; /* stats[1] of gkob200_csr_row_stats, or 0 if unknown */
    /* ELL (also the ELL part of HYBRID) */
    int64_t ell_stride, ell_width;
    const void* ell_col_idxs;
    const void* ell_values;
    /* SELL-P */
    int64_t slice_size, stride_factor, n_slices;
    const uint64_t* slice_sets;
    const uint64_t* slice_lengths;
    /* COO part of HYBRID */
    int64_t coo_nnz;
    const void* coo_row_idxs;
    const void* coo_col_idxs;
    const void* coo_values;
    /* scratch for kernels that need it (merge-path carries); may be NULL if unused */
    void* workspace;
    size_t workspace_bytes;
} gkob200_matrix;

/* c = A b  /  c = alpha A b + beta c  for any format of the descriptor
 * [ref: LinOp::apply(b,x) / apply(alpha,b,beta,x), include/ginkgo/core/base/lin_op.hpp:158-226] */
int gkob200_matrix_apply(void* stream, const gkob200_matrix* A, const void* b, int64_t b_stride, int64_t nrhs,
                         const void* alpha, const void* beta, void* c, int64_t c_stride);

enum gkob200_precond_kind { GKOB200_PRECOND_NONE = 0, GKOB200_PRECOND_JACOBI_SCALAR = 1, GKOB200_PRECOND_JACOBI_BLOCK = 2 };

/* Borrowed description of a generated preconditioner
 * [ref: preconditioner::Jacobi members, include/ginkgo/core/preconditioner/jacobi.hpp:578-610] */
typedef struct gkob200_precond {
    int32_t kind;
    int32_t value_type;
    const void* inv_diag;        /* scalar Jacobi: n values */
    /* block Jacobi: explicit inverses in Ginkgo's block_interleaved_storage_scheme */
    int64_t num_blocks;
    const void* block_pointers;  /* int32[num_blocks+1] */
    const void* blocks;
    int64_t block_offset, group_offset;
    int32_t group_power, max_block_size;
} gkob200_precond;

enum gkob200_stop_baseline { GKOB200_STOP_RHS_NORM = 0, GKOB200_STOP_INITIAL_RESNORM = 1, GKOB200_STOP_ABSOLUTE = 2 };

/* Combined(Iteration(max_iters), ResidualNorm(reduction_factor, baseline)) — criterion
 * ids 1 and 2 as stop::Combined assigns them [ref: core/stop/combined.cpp:40-58,
 * core/stop/iteration.cpp:40-51, core/stop/residual_norm.cpp:129-232]. */
typedef struct gkob200_stop {
    int64_t max_iters;
    double reduction_factor; /* <= 0 disables the residual criterion */
    int32_t baseline;        /* gkob200_stop_baseline */
    int32_t check_every;     /* host polls the device stop flag every this many iterations (>=1) */
} gkob200_stop;

typedef struct gkob200_solver gkob200_solver;

enum gkob200_solver_kind { GKOB200_SOLVER_CG = 0, GKOB200_SOLVER_BICGSTAB = 1, GKOB200_SOLVER_GMRES = 2 };

/* generate: blocking; allocates the solver workspace on the current device.
 * krylov_dim only for GMRES [ref: include/ginkgo/core/solver/gmres.hpp:57]. */
int gkob200_solver_create(int kind, const gkob200_matrix* A, const gkob200_precond* M, const gkob200_stop* stop,
                          int64_t nrhs, int64_t krylov_dim, gkob200_solver** out);
int gkob200_solver_destroy(gkob200_solver* s);

/* x <- solve(A, b) starting from the initial guess in x (apply_uses_initial_guess,
 * cg.hpp:72).  b, x: n x nrhs device arrays.  Blocking: returns when the device-side
 * stopping status says every column stopped. */
int gkob200_solver_apply(gkob200_solver* s, void* stream, const void* b, int64_t b_stride, void* x,
                         int64_t x_stride);
/* Same with HOST b/x: copies b and x in, solves, copies x back — what LinOp::apply does
 * through make_temporary_clone when the vectors live on the host executor
 * [ref: include/ginkgo/core/base/lin_op.hpp:158-167].  Blocking. */
int gkob200_solver_apply_host(gkob200_solver* s, void* stream, const void* b_host, void* x_host);

/* Results of the last apply (blocking reads). */
int64_t gkob200_solver_num_iterations(const gkob200_solver* s);
/* stop_status_host: nrhs bytes (gko::stopping_status encoding). */
int gkob200_solver_stop_status(const gkob200_solver* s, uint8_t* stop_status_host);
/* Copies min(cap, iterations+1) residual norms (column 0; as value type double) of the last
 * apply into hist_host; returns the number copied.  Entry i is ||r_i|| as the reference's
 * ResidualNorm criterion saw it at iteration i. */
int64_t gkob200_solver_residual_history(const gkob200_solver* s, double* hist_host, int64_t cap);
/* number of kernel launches + memcpy nodes the last apply enqueued (for gpu_launches) */
int64_t gkob200_solver_launch_count(const gkob200_solver* s);

#ifdef __cplusplus
}
#endif
#endif /* GKO_B200_SOLVER_H_ */
