// internal.h — C++ entry points shared between translation units of libgko_b200.so
// (not part of the C-ABI).
#pragma once
#include "common.cuh"

namespace gkob200 {

// Optional fusions for an SpMV launch issued from inside a solver iteration:
//  skip : device flag; a non-zero value turns the launch into a no-op (the solver
//         has stopped — mirrors the reference's per-column has_stopped() guards)
//  w/out: additionally compute out[0] = sum_i w[i] * (A b)[i] in the same pass
//         (single-pass grid reduction through `ws`)
constexpr int kHaloRuns = 4;

template <typename V>
struct SpmvFusion {
    const int* skip = nullptr;
    const V* w = nullptr;
    V* out = nullptr;
    V* out_sq = nullptr;    // optional second fused reduction: sum of result^2 (needs `out`)
    // optional: the kernel whose finaliser completes `out` also all-reduces p2p_count values at
    // p2p_buf over peer memory (p2p.cuh: finish_partials and the row-compressed non-local SpMV);
    // *on_fail is set to 1 when a peer did not show up (the solver's stopped flag)
    const void* p2p = nullptr;
    V* p2p_buf = nullptr;
    int p2p_count = 0;
    int* on_fail = nullptr;
    // optional (CSR row-block kernel, one right-hand side): the halo exchange and the non-local
    // block of a distributed matrix run inside this launch (device-resident plan, p2p.cuh)
    const HaloDev* halo = nullptr;
    int halo_push_ctas = 0;  // host copy of halo->n_push_ctas (extra CTAs at the front of the grid)
    // CTA slot -> row block for the interior slots, as up to kHaloRuns runs of consecutive blocks
    // held in the kernel parameters (no dependent load in front of the bulk copy): run i covers
    // slots [halo_run_slot[i], halo_run_slot[i+1]) -> blocks halo_run_block[i] + (slot - ...).
    // halo_runs == 0: look the slot up in halo->order (scattered boundary rows).
    int halo_debug = 0;      // measurement only (GKOB200_DIST_DEBUG): 4 = push CTAs idle, 8 = no non-local tail
    int halo_n_interior = 0;
    int halo_runs = 0;
    int halo_run_slot[kHaloRuns + 1] = {};
    int halo_run_block[kHaloRuns] = {};
    void* ws = nullptr;
    int64_t ws_blocks = 0;  // number of per-block partials `ws` has room for
};

// second stage of a deferred SpMV reduction (`grid` per-CTA partials per array in fu.ws)
template <typename V>
inline int launch_finish_partials(cudaStream_t s, int64_t grid, const SpmvFusion<V>& fu)
{
    const int64_t arrays = fu.out_sq ? 2 : 1;
    if (grid * arrays + 2 * kFinishMaxCtas > (fu.ws_blocks + 256) * kReduceMaxVals) return GKOB200_EWORKSPACE;
    V* parts = ws_partials<V>(fu.ws);
    finish_partials<V><<<finish_grid(grid), kFinishThreads, 0, s>>>(
        grid, parts, fu.out, fu.skip, fu.out_sq, parts + arrays * grid, ws_ticket(fu.ws),
        static_cast<const P2pDev*>(fu.p2p), fu.p2p_buf, fu.p2p_count, fu.on_fail);
    GKOB200_CHECK_LAUNCH();
    return 0;
}

template <typename V, typename I>
int csr_spmv_launch(cudaStream_t s, int64_t n_rows, int64_t n_cols, int64_t nnz, const I* row_ptrs,
                    const I* col_idxs, const V* values, const V* b, int64_t b_stride, int64_t nrhs,
                    const V* alpha, const V* beta, V* c, int64_t c_stride, int strategy,
                    int64_t max_block_nnz, void* workspace, size_t workspace_bytes,
                    const SpmvFusion<V>* fusion);

template <typename V, typename I>
int ell_spmv_launch(cudaStream_t s, int64_t n_rows, int64_t stride, int64_t width, const I* cols, const V* vals,
                    const V* b, int64_t b_stride, int64_t nrhs, const V* alpha, const V* beta, V* c,
                    int64_t c_stride, const SpmvFusion<V>* fusion);
template <typename V, typename I>
int sellp_spmv_launch(cudaStream_t s, int64_t n_rows, int64_t slice_size, const uint64_t* slice_sets,
                      const uint64_t* slice_lengths, const I* cols, const V* vals, const V* b, int64_t b_stride,
                      int64_t nrhs, const V* alpha, const V* beta, V* c, int64_t c_stride,
                      const SpmvFusion<V>* fusion);
// bulk-async (TMA) variants: return 1 when launched, 0 when the layout needs the fallback
template <typename V, typename I>
int sellp_spmv_tma_launch(cudaStream_t s, int64_t n_rows, int64_t slice_size, const uint64_t* slice_sets,
                          int64_t max_slice_len, int64_t total_cols, const I* cols, const V* vals, const V* b,
                          int64_t b_stride, const V* alpha, const V* beta, V* c, int64_t c_stride,
                          const SpmvFusion<V>* fusion);
template <typename V, typename I>
int ell_spmv_tma_launch(cudaStream_t s, int64_t n_rows, int64_t stride, int64_t width, const I* cols, const V* vals,
                        const V* b, int64_t b_stride, const V* alpha, const V* beta, V* c, int64_t c_stride,
                        const SpmvFusion<V>* fusion);
// accumulate == true: c += [alpha] A b (spmv2); false: c = A b / c = alpha A b + beta c
template <typename V, typename I>
int coo_spmv_launch(cudaStream_t s, int64_t n_rows, int64_t nnz, const I* rows, const I* cols, const V* vals,
                    const V* b, int64_t b_stride, int64_t nrhs, const V* alpha, const V* beta, bool accumulate,
                    V* c, int64_t c_stride, void* workspace, size_t workspace_bytes);

// c[row] = beta c[row] + alpha (A b)[row] on the listed rows only (alpha, beta required)
template <typename V, typename I>
int csr_rows_spmv_launch(cudaStream_t s, int64_t n_listed, const int32_t* row_list, const I* row_ptrs, const I* cols,
                         const V* vals, const V* b, int64_t b_stride, int64_t nrhs, const V* alpha, const V* beta,
                         V* c, int64_t c_stride, const SpmvFusion<V>* fusion);

// y = A x (alpha == beta == nullptr) or y = alpha A x + beta y, any format.
template <typename V>
int matrix_apply(cudaStream_t s, const gkob200_matrix& A, const V* b, int64_t b_stride, int64_t nrhs,
                 const V* alpha, const V* beta, V* c, int64_t c_stride, const SpmvFusion<V>* fusion);

// true if matrix_apply honours fusion->w/out itself for this descriptor
bool matrix_apply_fuses_dot(const gkob200_matrix& A, int64_t nrhs);

}  // namespace gkob200
