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
    GKOB200_FMT_HYBRID = 4, /* ELL part in the ell_* fields + COO part in the coo_* fields         */
    GKOB200_FMT_CSR_ROWS = 5 /* CSR over the listed rows only: row_list[n_listed] (int32, increasing),
                                row_ptrs[n_listed+1]; apply ACCUMULATES into the listed rows
                                (c[row] = beta*c[row] + alpha*sum).  Used for the non-local block of
                                the distributed matrix, whose rows are empty except at the slab faces */
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
    /* CSR_ROWS */
    const int32_t* row_list;
    int64_t n_listed;
    /* SELL-P, optional host-side facts that enable the bulk-async kernel (0 = unknown:
     * the thread-per-row kernel runs): max(slice_lengths) and slice_sets[n_slices] */
    int64_t sellp_max_slice_len;
    int64_t sellp_total_cols;
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

enum gkob200_solver_kind {
    GKOB200_SOLVER_CG = 0,
    GKOB200_SOLVER_BICGSTAB = 1,
    GKOB200_SOLVER_GMRES = 2,
    GKOB200_SOLVER_FCG = 3, /* core/solver/fcg.cpp:104-196 */
    GKOB200_SOLVER_CGS = 4  /* core/solver/cgs.cpp:104-212 */
};

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

/* ------------------------------------------------------------------------- *
 * Row-partitioned distributed matrix over the GPUs of one box (one process per GPU)
 * [ref: experimental::distributed::{Partition,Matrix,Vector},
 *  core/distributed/{partition,matrix,vector}.cpp; the reference uses MPI
 *  (include/ginkgo/core/base/mpi.hpp), here the used subset — all_to_all of sizes,
 *  all_to_all_v of gather indices, the per-apply halo exchange and the scalar all_reduce —
 *  runs on NCCL over NVLink].
 * ------------------------------------------------------------------------- */
typedef struct gkob200_dist_comm gkob200_dist_comm;
typedef struct gkob200_dist_matrix gkob200_dist_matrix;

/* 128-byte NCCL unique id; rank 0 creates it and hands it to the others out of band
 * (torch.distributed broadcast in the Python harness, MPI_Bcast in a Ginkgo shim). */
int gkob200_nccl_unique_id(void* out128);
/* blocking; size == 1 needs no id */
int gkob200_dist_comm_create(const void* id128, int rank, int size, gkob200_dist_comm** out);
int gkob200_dist_comm_destroy(gkob200_dist_comm* comm);
int gkob200_dist_comm_rank(const gkob200_dist_comm* comm);
int gkob200_dist_comm_size(const gkob200_dist_comm* comm);
/* 1 when the scalar all-reduces of this communicator run over peer memory (CUDA IPC over
 * NVLink, one small kernel) instead of ncclAllReduce: every rank on its own GPU and every peer
 * mappable; GKOB200_P2P=0 in the environment forces NCCL. */
int gkob200_dist_comm_uses_p2p(const gkob200_dist_comm* comm);
/* [ref: mpi::communicator::all_reduce / all_to_all / all_to_all_v,
 *  call sites core/distributed/vector.cpp:328-440, matrix.cpp:204-221]; device buffers */
int gkob200_dist_allreduce_sum_f64(gkob200_dist_comm* comm, void* stream, double* buf, int64_t count);
int gkob200_dist_allreduce_sum_f32(gkob200_dist_comm* comm, void* stream, float* buf, int64_t count);
int gkob200_dist_alltoall_i64(gkob200_dist_comm* comm, void* stream, const int64_t* send, int64_t* recv, int64_t count);
int gkob200_dist_alltoallv_i32(gkob200_dist_comm* comm, void* stream, const int32_t* send,
                               const int64_t* send_sizes_host, const int64_t* send_offsets_host, int32_t* recv,
                               const int64_t* recv_sizes_host, const int64_t* recv_offsets_host);

/* Partition kernels [ref: core/distributed/partition_kernels.hpp;
 * reference/distributed/partition_kernels.cpp:42-160] */
int gkob200_partition_build_ranges_from_global_size_i64(void* stream, int32_t num_parts, int64_t global_size,
                                                        int64_t* ranges);
int gkob200_partition_build_from_contiguous_i64(void* stream, int32_t num_parts, const int64_t* ranges,
                                                int64_t* range_bounds, int32_t* part_ids);
int gkob200_partition_build_from_mapping_i64(void* stream, int64_t n, const int32_t* mapping, int64_t* range_bounds,
                                             int32_t* part_ids, int64_t* num_ranges_dev, void* ws, size_t ws_bytes);
int gkob200_partition_build_starting_indices_i32_i64(void* stream, const int64_t* range_offsets,
                                                     const int32_t* range_parts, int64_t num_ranges, int32_t num_parts,
                                                     int32_t* num_empty_parts, int32_t* ranks, int32_t* sizes);

/* distributed_matrix::build_local_nonlocal [ref: core/distributed/matrix_kernels.hpp:51-70;
 * reference/distributed/matrix_kernels.cpp:49-236].  Input: global-index COO (int64), any
 * subset of the global matrix (entries of rows owned by other parts are ignored).
 * Outputs (capacity nnz each unless noted): local COO with local row/col indices; non-local
 * COO whose columns are renumbered by rank in the list of unique ghost columns sorted by
 * (owner part, global column); local_gather_idxs / non_local_to_global (one per ghost column);
 * recv_sizes[num_parts] (device).  counts_host = {n_local, n_non_local, n_ghost_cols}.
 * Blocking setup call (allocates its temporaries). */
#define GKOB200_DECL_DIST(V, VT)                                                                                  \
    int gkob200_dist_build_local_nonlocal_##V(                                                                    \
        void* stream, int64_t nnz, const int64_t* rows, const int64_t* cols, const VT* vals, int64_t row_num_ranges, \
        const int64_t* row_range_bounds, const int32_t* row_part_ids, const int32_t* row_range_starts,             \
        int64_t col_num_ranges, const int64_t* col_range_bounds, const int32_t* col_part_ids,                      \
        const int32_t* col_range_starts, int64_t global_cols, int32_t num_parts, int32_t local_part,               \
        int32_t* local_row_idxs, int32_t* local_col_idxs, VT* local_values, int32_t* non_local_row_idxs,           \
        int32_t* non_local_col_idxs, VT* non_local_values, int32_t* local_gather_idxs, int32_t* recv_sizes,        \
        int64_t* non_local_to_global, int64_t* counts_host);                                                      \
    int gkob200_dist_vector_build_local_##V(void* stream, int64_t nnz, const int64_t* rows, const int64_t* cols,   \
                                            const VT* vals, int64_t num_ranges, const int64_t* range_bounds,       \
                                            const int32_t* part_ids, const int32_t* range_starts,                  \
                                            int32_t local_part, VT* local, int64_t local_stride);
GKOB200_DECL_DIST(f64, double)
GKOB200_DECL_DIST(f32, float)

/* distributed::Matrix: local block + non-local block + halo plan.  gather_idxs (device): the
 * local row indices this rank sends, grouped by destination; send/recv sizes per peer (host).
 * All descriptors and arrays are borrowed. */
int gkob200_dist_matrix_create(gkob200_dist_comm* comm, const gkob200_matrix* local, const gkob200_matrix* non_local,
                               const int32_t* gather_idxs, const int64_t* send_sizes_host,
                               const int64_t* recv_sizes_host, gkob200_dist_matrix** out);
int gkob200_dist_matrix_destroy(gkob200_dist_matrix* m);
/* 1 when apply() is ONE launch: the first CTAs of the local SpMV store the halo entries into
 * the neighbours' receive windows over peer memory (NVLink) and the row blocks with non-local
 * entries, scheduled last, wait for the neighbours' epoch flags and finish their row sums with
 * the non-local block.  Needs: every rank on its own GPU with all peers mappable (CUDA IPC),
 * one right-hand side, local block on the CSR row-block kernel (int32), non-local block
 * row-compressed.  0: pack -> ncclSend/ncclRecv on a side stream overlapped with the local
 * SpMV -> separate non-local SpMV.  GKOB200_FUSED_HALO=0 / GKOB200_P2P=0 force the latter.
 * gkob200_dist_matrix_create is collective over the communicator either way. */
int gkob200_dist_matrix_uses_fused_halo(const gkob200_dist_matrix* m);
/* x_local = A b  /  x_local = alpha A b + beta x_local
 * [ref: core/distributed/matrix.cpp:263-369 communicate + apply_impl] */
int gkob200_dist_matrix_apply(gkob200_dist_matrix* m, void* stream, const void* b_local, int64_t b_stride,
                              int64_t nrhs, const void* alpha, const void* beta, void* x_local, int64_t x_stride);
/* Distributed CG (kind must be GKOB200_SOLVER_CG; preconditioner none / scalar Jacobi on the
 * local block): the handle is used with gkob200_solver_apply / _num_iterations / ... above. */
int gkob200_dist_solver_create(int kind, gkob200_dist_matrix* A, const gkob200_precond* M, const gkob200_stop* stop,
                               int64_t nrhs, gkob200_solver** out);

#ifdef __cplusplus
}
#endif
#endif /* GKO_B200_SOLVER_H_ */
