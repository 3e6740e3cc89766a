// solver_common.cuh — device-resident solver state, the on-device stopping
// criterion, and the host-side solver base class shared by CG / BiCGSTAB / GMRES.
#pragma once
#include <vector>

#include "internal.h"
#include "launch.cuh"

namespace gkob200 {

// Lives in device memory; one per solver object.
struct SolverState {
    int stopped;     // every column has stopped: all later kernels of the apply are no-ops
    int iter;        // the reference's `iter` (number of completed iterations)
    int final_iter;  // value of `iter` when the solver stopped
    int one_changed; // last criterion check changed at least one column (BiCGSTAB finalize)
    int hist_cap;    // entries the residual-history buffer holds (writes beyond it are dropped)
};

// Residual-history length for a solver: every iteration when an Iteration criterion bounds
// the solve, a fixed 64 Ki window otherwise (stop.py passes max_iters = 2^31-2 for "no
// Iteration criterion"; the history of such a solve is truncated, never overrun).
inline int64_t history_capacity(int64_t max_iters)
{
    if (max_iters < 0) max_iters = 0;
    return max_iters <= (int64_t(1) << 20) ? max_iters + 2 : (int64_t(1) << 16);
}

// What stop::Combined{Iteration(max_iters), ResidualNorm(factor, baseline)} does in
// one check [ref: core/stop/combined.cpp:40-58, iteration.cpp:40-51,
// reference/stop/residual_norm_kernels.cpp:58-84] — executed by ONE device thread
// so that no host round trip is needed per iteration (the reference's CUDA
// criterion ends in two blocking D2H copies, cuda/stop/residual_norm_kernels.cu:117-118).
//   tau[j]      residual norm of column j
//   hist        residual history (column 0), entry st->iter, nullable
//   advance     count this check as the start of a new iteration
template <typename V>
__device__ __forceinline__ void criterion_check(SolverState* st, int64_t k, const V* tau, const V* orig_tau,
                                                V factor, int64_t max_iters, bool set_finalized,
                                                uint8_t* stop_status, V* hist, bool advance)
{
    // st->iter is what the reference's `iter` becomes at its next "++iter"; a check that
    // does not start a new iteration (BiCGSTAB's mid-iteration check) sees the current one
    const int it = advance ? st->iter : st->iter - 1;
    if (hist && advance && it < st->hist_cap) hist[it] = tau[0];
    bool one_changed = false, all = false;
    // criterion 1: Iteration
    if (it >= max_iters) {
        for (int64_t j = 0; j < k; ++j) {
            uint8_t s = stop_status[j];
            if (!status_has_stopped(s)) {
                s |= 1;
                if (set_finalized) s |= 0x40;
                stop_status[j] = s;
            }
        }
        one_changed = true;
        all = true;
    } else if (factor > V(0)) {
        // criterion 2: ResidualNorm
        all = true;
        for (int64_t j = 0; j < k; ++j) {
            uint8_t s = stop_status[j];
            if (tau[j] < mul_rn(factor, orig_tau[j])) {
                if (!status_has_stopped(s)) {
                    s |= 0x80 | 2;
                    if (set_finalized) s |= 0x40;
                    stop_status[j] = s;
                }
                one_changed = true;
            }
            if (!status_has_stopped(s)) all = false;
        }
    }
    st->one_changed = one_changed;
    if (all) {
        st->stopped = 1;
        st->final_iter = it;
    } else if (advance) {
        st->iter = it + 1;
    }
}

}  // namespace gkob200

// The opaque C handle is the base class.
struct gkob200_solver {
    virtual ~gkob200_solver() {}
    virtual int apply(cudaStream_t s, const void* b, int64_t bs, void* x, int64_t xs) = 0;
    virtual int apply_host(cudaStream_t s, const void* b_host, void* x_host) = 0;
    int64_t num_iterations = 0;
    int64_t launch_count = 0;
    int64_t nrhs = 1;
    std::vector<uint8_t> stop_status_host;
    std::vector<double> residual_history;
};

namespace gkob200 {

// Simple owning device buffer (solver workspace; allocated at create, freed at destroy).
struct DevBuf {
    void* p = nullptr;
    size_t bytes = 0;
    int alloc(size_t b)
    {
        if (p) cudaFree(p);
        p = nullptr;
        bytes = b;
        if (b == 0) return 0;
        cudaError_t e = cudaMalloc(&p, b);
        if (e != cudaSuccess) {
            p = nullptr;
            return static_cast<int>(e);
        }
        return static_cast<int>(cudaMemset(p, 0, b));
    }
    ~DevBuf()
    {
        if (p) cudaFree(p);
    }
    template <typename T>
    T* as()
    {
        return reinterpret_cast<T*>(p);
    }
};

// ---- typed dispatch to the C-ABI BLAS-1 / step kernels --------------------------
#define GKOB200_TYPED2(name)                                                    \
    template <typename... A> inline int name(double*, A... a) { return gkob200_##name##_f64(a...); } \
    template <typename... A> inline int name(float*, A... a) { return gkob200_##name##_f32(a...); }
namespace typed {
// first argument is a type tag (nullptr cast to V*)
GKOB200_TYPED2(dense_fill)
GKOB200_TYPED2(dense_copy)
GKOB200_TYPED2(dense_scale)
GKOB200_TYPED2(dense_inv_scale)
GKOB200_TYPED2(dense_add_scaled)
GKOB200_TYPED2(dense_sub_scaled)
GKOB200_TYPED2(dense_compute_dot)
GKOB200_TYPED2(dense_compute_norm2)
GKOB200_TYPED2(jacobi_simple_scalar_apply)
GKOB200_TYPED2(jacobi_block_simple_apply)
GKOB200_TYPED2(fcg_initialize)
GKOB200_TYPED2(fcg_step_1)
GKOB200_TYPED2(fcg_step_2)
GKOB200_TYPED2(cgs_initialize)
GKOB200_TYPED2(cgs_step_1)
GKOB200_TYPED2(cgs_step_2)
GKOB200_TYPED2(cgs_step_3)
GKOB200_TYPED2(bicgstab_initialize)
GKOB200_TYPED2(bicgstab_step_1)
GKOB200_TYPED2(bicgstab_step_2)
GKOB200_TYPED2(bicgstab_step_3)
GKOB200_TYPED2(bicgstab_finalize)
GKOB200_TYPED2(gmres_initialize)
GKOB200_TYPED2(gmres_restart)
GKOB200_TYPED2(gmres_multi_axpy)
GKOB200_TYPED2(gmres_hessenberg_qr)
GKOB200_TYPED2(gmres_solve_krylov)
}  // namespace typed
#undef GKOB200_TYPED2

template <typename V>
__global__ void criterion_kernel(SolverState* st, int64_t k, const V* tau, const V* orig_tau, V factor,
                                 int64_t max_iters, bool set_finalized, uint8_t* stop_status, V* hist, bool advance)
{
    if (st->stopped) return;
    criterion_check(st, k, tau, orig_tau, factor, max_iters, set_finalized, stop_status, hist, advance);
}

template <typename V>
__global__ void init_state_kernel(SolverState* st, V* one, V* neg_one, int64_t k, int hist_cap)
{
    st->hist_cap = hist_cap;
    st->stopped = 0;
    st->iter = 0;
    st->final_iter = 0;
    st->one_changed = 0;
    for (int64_t j = 0; j < k; ++j) {
        one[j] = V(1);
        neg_one[j] = V(-1);
    }
}

// Shared plumbing of the solver objects: descriptors, device state, criterion,
// preconditioner application, result read-back, host-buffer apply.
template <typename V>
struct SolverBase : gkob200_solver {
    gkob200_matrix A;
    gkob200_precond M;
    gkob200_stop stop;
    int64_t n = 0, k = 1;
    DevBuf state, status, hist, ws, consts, taus;
    SolverState* h_state = nullptr;  // pinned
    int chunk = 8;
    DevBuf host_b, host_x;
    static constexpr V* tag() { return static_cast<V*>(nullptr); }

    ~SolverBase() override
    {
        if (h_state) cudaFreeHost(h_state);
    }
    SolverState* st() { return state.template as<SolverState>(); }
    uint8_t* stat() { return status.template as<uint8_t>(); }
    V* one() { return consts.template as<V>(); }
    V* neg_one() { return consts.template as<V>() + k; }
    V* tau() { return taus.template as<V>(); }
    V* orig_tau() { return taus.template as<V>() + k; }

    int init_base()
    {
        n = A.n_rows;
        if (A.n_rows != A.n_cols) return GKOB200_EINVAL;
        // iterations per CUDA graph / per host poll; bounded so that a huge check_every cannot
        // turn into a graph of hundreds of thousands of nodes
        chunk = stop.check_every > 0 ? (stop.check_every < 512 ? stop.check_every : 512) : 8;
        int rc;
        if ((rc = state.alloc(sizeof(SolverState)))) return rc;
        if ((rc = status.alloc(static_cast<size_t>(k) + 16))) return rc;
        if ((rc = consts.alloc(2 * k * sizeof(V)))) return rc;
        if ((rc = taus.alloc(2 * k * sizeof(V)))) return rc;
        if ((rc = hist.alloc(static_cast<size_t>(history_capacity(stop.max_iters)) * sizeof(V)))) return rc;
        if ((rc = ws.alloc(GKOB200_REDUCE_WS_BYTES))) return rc;
        GKOB200_CUDA(cudaHostAlloc(reinterpret_cast<void**>(&h_state), 2 * sizeof(SolverState), cudaHostAllocDefault));
        stop_status_host.assign(k, 0);
        return 0;
    }

    int reset_state(cudaStream_t s)
    {
        init_state_kernel<V><<<1, 1, 0, s>>>(st(), one(), neg_one(), k, static_cast<int>(hist.bytes / sizeof(V)));
        ++launch_count;
        GKOB200_CHECK_LAUNCH();
        return 0;
    }

    // z = M^-1 r  (identity: copy, reference core/matrix/identity.cpp:47-50)
    int precond_apply(cudaStream_t s, const V* in, int64_t is, V* out, int64_t os)
    {
        ++launch_count;
        if (M.kind == GKOB200_PRECOND_NONE) return typed::dense_copy(tag(), s, n, k, in, is, out, os);
        if (M.kind == GKOB200_PRECOND_JACOBI_SCALAR)
            return typed::jacobi_simple_scalar_apply(tag(), s, n, k, static_cast<const V*>(M.inv_diag), in, is, out, os);
        if (M.kind == GKOB200_PRECOND_JACOBI_BLOCK)
            return typed::jacobi_block_simple_apply(tag(), s, M.num_blocks, static_cast<const int32_t*>(M.block_pointers),
                                                    static_cast<const V*>(M.blocks), M.block_offset, M.group_offset,
                                                    static_cast<int>(M.group_power), n, k, in, is, out, os);
        return GKOB200_EUNSUPPORTED;
    }

    // starting_tau of ResidualNormBase [ref: core/stop/residual_norm.cpp:129-186]
    int baseline_norm(cudaStream_t s, const V* b, int64_t bs, const V* r, int64_t rs)
    {
        ++launch_count;
        switch (stop.baseline) {
        case GKOB200_STOP_RHS_NORM: return typed::dense_compute_norm2(tag(), s, n, k, b, bs, orig_tau(), ws.p);
        case GKOB200_STOP_INITIAL_RESNORM: return typed::dense_compute_norm2(tag(), s, n, k, r, rs, orig_tau(), ws.p);
        case GKOB200_STOP_ABSOLUTE: return typed::dense_fill(tag(), s, int64_t(1), k, orig_tau(), k, V(1));
        default: return GKOB200_EINVAL;
        }
    }

    int check(cudaStream_t s, const V* tau_, bool set_finalized, bool advance)
    {
        criterion_kernel<V><<<1, 1, 0, s>>>(st(), k, tau_, orig_tau(), static_cast<V>(stop.reduction_factor),
                                            stop.max_iters, set_finalized, stat(), hist.template as<V>(), advance);
        ++launch_count;
        GKOB200_CHECK_LAUNCH();
        return 0;
    }

    // blocking read of the device state
    int poll(cudaStream_t s, bool* stopped)
    {
        GKOB200_CUDA(cudaMemcpyAsync(&h_state[0], st(), sizeof(SolverState), cudaMemcpyDeviceToHost, s));
        GKOB200_CUDA(cudaStreamSynchronize(s));
        *stopped = h_state[0].stopped != 0;
        return 0;
    }

    int finish(cudaStream_t s)
    {
        GKOB200_CUDA(cudaMemcpyAsync(&h_state[0], st(), sizeof(SolverState), cudaMemcpyDeviceToHost, s));
        GKOB200_CUDA(cudaMemcpyAsync(stop_status_host.data(), status.p, k, cudaMemcpyDeviceToHost, s));
        GKOB200_CUDA(cudaStreamSynchronize(s));
        num_iterations = h_state[0].final_iter;
        const int64_t cap = static_cast<int64_t>(hist.bytes / sizeof(V));
        const int64_t m = num_iterations + 1 < cap ? num_iterations + 1 : cap;
        std::vector<V> tmp(m);
        GKOB200_CUDA(cudaMemcpy(tmp.data(), hist.p, m * sizeof(V), cudaMemcpyDeviceToHost));
        residual_history.assign(tmp.begin(), tmp.end());
        return 0;
    }

    int apply_host(cudaStream_t s, const void* b_host, void* x_host) override
    {
        const size_t bytes = static_cast<size_t>(n) * k * sizeof(V);
        int rc;
        if (host_b.bytes != bytes) {
            if ((rc = host_b.alloc(bytes))) return rc;
            if ((rc = host_x.alloc(bytes))) return rc;
        }
        if (bytes == 0) return 0;
        GKOB200_CUDA(cudaMemcpyAsync(host_b.p, b_host, bytes, cudaMemcpyHostToDevice, s));
        GKOB200_CUDA(cudaMemcpyAsync(host_x.p, x_host, bytes, cudaMemcpyHostToDevice, s));
        if ((rc = this->apply(s, host_b.p, k, host_x.p, k))) return rc;
        GKOB200_CUDA(cudaMemcpyAsync(x_host, host_x.p, bytes, cudaMemcpyDeviceToHost, s));
        GKOB200_CUDA(cudaStreamSynchronize(s));
        launch_count += 3;
        return 0;
    }
};

gkob200_solver* make_bicgstab_f64(const gkob200_matrix*, const gkob200_precond*, const gkob200_stop*, int64_t nrhs, int* rc);
gkob200_solver* make_bicgstab_f32(const gkob200_matrix*, const gkob200_precond*, const gkob200_stop*, int64_t nrhs, int* rc);
gkob200_solver* make_gmres_f64(const gkob200_matrix*, const gkob200_precond*, const gkob200_stop*, int64_t nrhs, int64_t krylov_dim, int* rc);
gkob200_solver* make_gmres_f32(const gkob200_matrix*, const gkob200_precond*, const gkob200_stop*, int64_t nrhs, int64_t krylov_dim, int* rc);
gkob200_solver* make_fcg(const gkob200_matrix*, const gkob200_precond*, const gkob200_stop*, int64_t nrhs, int* rc);
gkob200_solver* make_cgs(const gkob200_matrix*, const gkob200_precond*, const gkob200_stop*, int64_t nrhs, int* rc);

gkob200_solver* make_cg_f64(const gkob200_matrix*, const gkob200_precond*, const gkob200_stop*, int64_t nrhs, int* rc);
gkob200_solver* make_cg_f32(const gkob200_matrix*, const gkob200_precond*, const gkob200_stop*, int64_t nrhs, int* rc);

}  // namespace gkob200
