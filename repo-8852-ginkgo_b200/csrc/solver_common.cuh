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
};

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
    const int it = st->iter;
    if (hist && advance) hist[it] = tau[0];
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

gkob200_solver* make_cg_f64(const gkob200_matrix*, const gkob200_precond*, const gkob200_stop*, int64_t nrhs, int* rc);
gkob200_solver* make_cg_f32(const gkob200_matrix*, const gkob200_precond*, const gkob200_stop*, int64_t nrhs, int* rc);

}  // namespace gkob200
