// solver_cg.cu — the CG solver object: the reference iteration
// (core/solver/cg.cpp:107-194) re-issued as three fused kernels per iteration,
// graph-captured, with all scalars and the stopping criterion on the device.
//
// Reference iteration                          this file
//   z = M^-1 r ; rho = r.z ; tau = ||r||        \
//   ++iter ; criterion check                     > cg_update  (x,r,z in one pass + 2 dots
//   [previous] x += t p ; r -= t q ; swap rho   /              + on-device criterion)
//   p = z + (rho/prev_rho) p                       cg_direction
//   q = A p ; beta = p.q                           SpMV with the dot fused in its epilogue
//
// HBM passes per iteration (values): cg_update 5 reads + 3 writes (4+2 without
// preconditioner), cg_direction 2+1, SpMV matrix + 1 + 1  => 13 n + matrix, against the
// reference's 18 n + matrix + preconditioner (core/solver/cg.cpp:148-156), in 3 launches
// instead of ~11 launches + 2 blocking D2H copies.
//
// Multi-RHS / strided vectors take the general path built from the 1:1 step kernels
// of cg_kernels.cu (still entirely on the GPU).
#include "solver_common.cuh"

namespace gkob200 {
namespace {

enum { S_RHO = 0, S_PREV_RHO, S_BETA, S_TAU, S_ORIG_TAU, S_ONE, S_NEG_ONE, S_RR, S_COUNT };

template <typename V>
struct CgParams {
    int64_t n;
    V* x;
    V* r;
    V* z;
    V* p;
    V* q;
    const V* inv_diag;  // scalar Jacobi or nullptr
    V* sc;              // scalars, column-major blocks of `k` (here k == 1)
    SolverState* st;
    uint8_t* stop_status;
    V* hist;
    V factor;
    int64_t max_iters;
    void* ws;
};

// mode 0: no preconditioner (z == r, never stored)   1: scalar Jacobi fused
// mode 2: generic preconditioner: only x/r are updated here, dots follow in cg_dots
// Wide: every vector is 16-byte aligned — two rows per thread and iteration through 128-bit
// accesses (the 64-bit version ran the 8 streams of the Jacobi variant at 0.83 of the HBM peak).
// Element-wise results are the same bits either way; only the association of the two
// reductions differs.
template <typename V>
struct Pair;
template <>
struct Pair<double> { using type = double2; };
template <>
struct Pair<float> { using type = float2; };

template <typename V, int Mode, bool First, bool Wide>
__global__ void __launch_bounds__(256) cg_update(CgParams<V> P)
{
    if (P.st->stopped) return;
    V t = V(0);
    bool upd = false;
    if (!First) {
        const V beta = P.sc[S_BETA];
        upd = beta != V(0);
        if (upd) t = div_rn(P.sc[S_RHO], beta);
    }
    V acc[2] = {V(0), V(0)};
    const int64_t step = static_cast<int64_t>(gridDim.x) * blockDim.x;
    const int64_t tid0 = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x;
    auto one = [&](V ri, V xi, V pi, V qi, V di, V& x_out, V& r_out, V& z_out) {
        if (upd) {
            x_out = add_rn(xi, mul_rn(t, pi));
            ri = sub_rn(ri, mul_rn(t, qi));
        }
        r_out = ri;
        if (Mode == 2) return;
        V zi = ri;
        if (Mode == 1) zi = mul_rn(ri, di);
        z_out = zi;
        acc[0] += ri * zi;
        if (Mode == 1) acc[1] += ri * ri;
    };
    if (Wide) {
        using V2 = typename Pair<V>::type;
        const int64_t n2 = P.n / 2;
        V2* __restrict__ x2 = reinterpret_cast<V2*>(P.x);
        V2* __restrict__ r2 = reinterpret_cast<V2*>(P.r);
        V2* __restrict__ z2 = reinterpret_cast<V2*>(P.z);
        const V2* __restrict__ p2 = reinterpret_cast<const V2*>(P.p);
        const V2* __restrict__ q2 = reinterpret_cast<const V2*>(P.q);
        const V2* __restrict__ d2 = reinterpret_cast<const V2*>(P.inv_diag);
        for (int64_t i = tid0; i < n2; i += step) {
            const V2 rv = r2[i];
            V2 xv = {V(0), V(0)}, pv = xv, qv = xv, dv = xv;
            if (upd) {
                xv = x2[i];
                pv = p2[i];
                qv = q2[i];
            }
            if (Mode == 1) dv = d2[i];
            V2 xo = xv, ro, zo;
            one(rv.x, xv.x, pv.x, qv.x, dv.x, xo.x, ro.x, zo.x);
            one(rv.y, xv.y, pv.y, qv.y, dv.y, xo.y, ro.y, zo.y);
            if (upd) {
                x2[i] = xo;
                r2[i] = ro;
            }
            if (Mode == 1) z2[i] = zo;
        }
        if ((P.n & 1) && tid0 == 0) {
            const int64_t i = P.n - 1;
            V xo = V(0), ro, zo;
            one(P.r[i], upd ? P.x[i] : V(0), upd ? P.p[i] : V(0), upd ? P.q[i] : V(0), Mode == 1 ? P.inv_diag[i] : V(0), xo,
                ro, zo);
            if (upd) {
                P.x[i] = xo;
                P.r[i] = ro;
            }
            if (Mode == 1) P.z[i] = zo;
        }
    } else {
        for (int64_t i = tid0; i < P.n; i += step) {
            V xo = V(0), ro, zo;
            one(P.r[i], upd ? P.x[i] : V(0), upd ? P.p[i] : V(0), upd ? P.q[i] : V(0), Mode == 1 ? P.inv_diag[i] : V(0), xo,
                ro, zo);
            if (upd) {
                P.x[i] = xo;
                P.r[i] = ro;
            }
            if (Mode == 1) P.z[i] = zo;
        }
    }
    if (Mode == 2) return;
    CgParams<V> Q = P;
    grid_reduce<2>(acc, ws_partials<V>(P.ws), ws_ticket(P.ws), [Q](V(&tot)[2]) {
        if (!First) Q.sc[S_PREV_RHO] = Q.sc[S_RHO];
        Q.sc[S_RHO] = tot[0];
        const V rr = Mode == 1 ? tot[1] : tot[0];
        Q.sc[S_TAU] = sqrt_rn(rr);
        criterion_check(Q.st, 1, Q.sc + S_TAU, Q.sc + S_ORIG_TAU, Q.factor, Q.max_iters, true, Q.stop_status,
                        Q.hist, true);
    });
}

// generic preconditioner: rho = r.z, tau = ||r|| after z = M^-1 r was applied
template <typename V, bool First>
__global__ void __launch_bounds__(256) cg_dots(CgParams<V> P)
{
    if (P.st->stopped) return;
    V acc[2] = {V(0), V(0)};
    const int64_t step = static_cast<int64_t>(gridDim.x) * blockDim.x;
    for (int64_t i = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; i < P.n; i += step) {
        const V ri = P.r[i];
        acc[0] += ri * P.z[i];
        acc[1] += ri * ri;
    }
    CgParams<V> Q = P;
    grid_reduce<2>(acc, ws_partials<V>(P.ws), ws_ticket(P.ws), [Q](V(&tot)[2]) {
        if (!First) Q.sc[S_PREV_RHO] = Q.sc[S_RHO];
        Q.sc[S_RHO] = tot[0];
        Q.sc[S_TAU] = sqrt_rn(tot[1]);
        criterion_check(Q.st, 1, Q.sc + S_TAU, Q.sc + S_ORIG_TAU, Q.factor, Q.max_iters, true, Q.stop_status,
                        Q.hist, true);
    });
}

template <typename V, bool ZisR, bool Wide>
__global__ void __launch_bounds__(256) cg_direction(CgParams<V> P)
{
    if (P.st->stopped) return;
    const V prev = P.sc[S_PREV_RHO];
    const bool zero_prev = prev == V(0);
    const V t = zero_prev ? V(0) : div_rn(P.sc[S_RHO], prev);
    const V* __restrict__ z = ZisR ? P.r : P.z;
    const int64_t step = static_cast<int64_t>(gridDim.x) * blockDim.x;
    const int64_t tid0 = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x;
    if (Wide) {
        using V2 = typename Pair<V>::type;
        const V2* __restrict__ z2 = reinterpret_cast<const V2*>(z);
        V2* __restrict__ p2 = reinterpret_cast<V2*>(P.p);
        for (int64_t i = tid0; i < P.n / 2; i += step) {
            const V2 zv = z2[i];
            if (zero_prev) {
                p2[i] = zv;
            } else {
                const V2 pv = p2[i];
                V2 o;
                o.x = add_rn(zv.x, mul_rn(t, pv.x));
                o.y = add_rn(zv.y, mul_rn(t, pv.y));
                p2[i] = o;
            }
        }
        if ((P.n & 1) && tid0 == 0) {
            const int64_t i = P.n - 1;
            P.p[i] = zero_prev ? z[i] : add_rn(z[i], mul_rn(t, P.p[i]));
        }
        return;
    }
    for (int64_t i = tid0; i < P.n; i += step) {
        P.p[i] = zero_prev ? z[i] : add_rn(z[i], mul_rn(t, P.p[i]));
    }
}

// out = sum w[i]*y[i], skipped when the solver has stopped (formats without a fused dot)
template <typename V>
__global__ void __launch_bounds__(256) dot_skip(int64_t n, const V* __restrict__ a, const V* __restrict__ b,
                                                V* out, const int* skip, void* ws)
{
    if (skip && *skip) return;
    V acc[1] = {V(0)};
    const int64_t step = static_cast<int64_t>(gridDim.x) * blockDim.x;
    for (int64_t i = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; i < n; i += step)
        acc[0] += a[i] * b[i];
    grid_reduce<1>(acc, ws_partials<V>(ws), ws_ticket(ws), [out](V(&tot)[1]) { out[0] = tot[0]; });
}

// multi-RHS general path: per-column criterion on device
template <typename V>
__global__ void cg_criterion_k(SolverState* st, int64_t k, const V* tau, const V* orig_tau, V factor,
                               int64_t max_iters, uint8_t* stop_status, V* hist)
{
    if (st->stopped) return;
    criterion_check(st, k, tau, orig_tau, factor, max_iters, true, stop_status, hist, true);
}

template <typename V>
__global__ void init_state(SolverState* st, V* sc, int64_t k, int hist_cap)
{
    st->hist_cap = hist_cap;
    st->stopped = 0;
    st->iter = 0;
    st->final_iter = 0;
    st->one_changed = 0;
    for (int64_t j = 0; j < k; ++j) {
        sc[S_ONE * k + j] = V(1);
        sc[S_NEG_ONE * k + j] = V(-1);
    }
}

template <typename V>
struct CgSolver : gkob200_solver {
    gkob200_matrix A;
    gkob200_precond M;
    gkob200_stop stop;
    int64_t n = 0, k = 1;
    DevBuf vecs, scalars, state, status, hist, ws;
    SolverState* h_state = nullptr;  // pinned, 2 slots
    cudaEvent_t ev[2] = {nullptr, nullptr};
    cudaGraphExec_t graph = nullptr;
    cudaStream_t cap_stream = nullptr;  // capture stream (the legacy default stream cannot be captured)
    void* graph_x = nullptr;
    int64_t graph_xs = 0;
    int chunk = 8;
    int64_t launches_per_iter = 0;
    int64_t ws_blocks = 0;
    DevBuf host_b, host_x;  // device staging for apply_host

    ~CgSolver()
    {
        if (graph) cudaGraphExecDestroy(graph);
        if (cap_stream) cudaStreamDestroy(cap_stream);
        if (h_state) cudaFreeHost(h_state);
        for (auto& e : ev)
            if (e) cudaEventDestroy(e);
    }

    int init()
    {
        n = A.n_rows;
        if (A.n_rows != A.n_cols) return GKOB200_EINVAL;
        // iterations per CUDA graph / per host poll; bounded so that a huge check_every cannot
        // turn into a graph of hundreds of thousands of nodes
        chunk = stop.check_every > 0 ? (stop.check_every < 512 ? stop.check_every : 512) : 8;
        int rc;
        if ((rc = vecs.alloc(static_cast<size_t>(n) * k * 4 * sizeof(V)))) return rc;
        if ((rc = scalars.alloc(static_cast<size_t>(S_COUNT) * k * sizeof(V)))) return rc;
        if ((rc = state.alloc(sizeof(SolverState)))) return rc;
        if ((rc = status.alloc(static_cast<size_t>(k) + 16))) return rc;
        if ((rc = hist.alloc(static_cast<size_t>(history_capacity(stop.max_iters)) * sizeof(V)))) return rc;
        // reduction scratch: room for one partial per CTA of the largest grid any fused
        // kernel launches (the SpMV with the fused dot runs one CTA per 128 rows)
        ws_blocks = ceildiv(n, 128) + 1;
        if (ws_blocks < kReduceMaxBlocks) ws_blocks = kReduceMaxBlocks;
        if ((rc = ws.alloc(reduce_ws_bytes(ws_blocks)))) return rc;
        GKOB200_CUDA(cudaHostAlloc(reinterpret_cast<void**>(&h_state), 2 * sizeof(SolverState), cudaHostAllocDefault));
        for (auto& e : ev) GKOB200_CUDA(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
        GKOB200_CUDA(cudaStreamCreateWithFlags(&cap_stream, cudaStreamNonBlocking));
        stop_status_host.assign(k, 0);
        return 0;
    }

    V* r() { return vecs.as<V>(); }
    V* z() { return vecs.as<V>() + n * k; }
    V* p() { return vecs.as<V>() + 2 * n * k; }
    V* q() { return vecs.as<V>() + 3 * n * k; }
    V* sc(int which) { return scalars.as<V>() + which * k; }

    int precond_apply(cudaStream_t s, const V* in, V* out);  // generic (block Jacobi), defined below

    // 128-bit accesses need every vector the update / direction kernels touch on 16 bytes
    bool wide_ok(const V* x) const
    {
        auto al = [](const void* p) { return reinterpret_cast<uintptr_t>(p) % 16 == 0; };
        return (static_cast<size_t>(n) * sizeof(V)) % 16 == 0 && al(x) && al(vecs.p) &&
               (M.kind != GKOB200_PRECOND_JACOBI_SCALAR || al(M.inv_diag));
    }

    CgParams<V> params(V* x)
    {
        CgParams<V> P;
        P.n = n;
        P.x = x;
        P.r = r();
        P.z = z();
        P.p = p();
        P.q = q();
        P.inv_diag = M.kind == GKOB200_PRECOND_JACOBI_SCALAR ? static_cast<const V*>(M.inv_diag) : nullptr;
        P.sc = scalars.as<V>();
        P.st = state.as<SolverState>();
        P.stop_status = status.as<uint8_t>();
        P.hist = hist.as<V>();
        P.factor = static_cast<V>(stop.reduction_factor);
        P.max_iters = stop.max_iters;
        P.ws = ws.p;
        return P;
    }

    template <bool First>
    int enqueue_update(cudaStream_t s, V* x)
    {
        CgParams<V> P = params(x);
        const bool wide = wide_ok(x);
        const int grid = grid_for(wide ? (n + 1) / 2 : n, 256, 6);
#define GKOB200_UPD(MODE)                                            \
    if (wide) cg_update<V, MODE, First, true><<<grid, 256, 0, s>>>(P); \
    else cg_update<V, MODE, First, false><<<grid, 256, 0, s>>>(P)
        if (M.kind == GKOB200_PRECOND_NONE) {
            GKOB200_UPD(0);
            ++launch_count;
        } else if (M.kind == GKOB200_PRECOND_JACOBI_SCALAR) {
            GKOB200_UPD(1);
            ++launch_count;
        } else {
            if (!First) {
                GKOB200_UPD(2);
                ++launch_count;
            }
#undef GKOB200_UPD
            int rc = precond_apply(s, r(), z());
            if (rc) return rc;
            cg_dots<V, First><<<grid, 256, 0, s>>>(P);
            ++launch_count;
        }
        GKOB200_CHECK_LAUNCH();
        return 0;
    }

    int enqueue_iteration(cudaStream_t s, V* x)
    {
        CgParams<V> P = params(x);
        const bool wide = wide_ok(x);
        const int grid = grid_for(wide ? (n + 1) / 2 : n, 256, 6);
        if (M.kind == GKOB200_PRECOND_NONE) {
            if (wide) cg_direction<V, true, true><<<grid, 256, 0, s>>>(P);
            else cg_direction<V, true, false><<<grid, 256, 0, s>>>(P);
        } else {
            if (wide) cg_direction<V, false, true><<<grid, 256, 0, s>>>(P);
            else cg_direction<V, false, false><<<grid, 256, 0, s>>>(P);
        }
        ++launch_count;
        GKOB200_CHECK_LAUNCH();
        SpmvFusion<V> fu;
        fu.skip = &state.as<SolverState>()->stopped;
        const bool fuses = matrix_apply_fuses_dot(A, 1);
        if (fuses) {
            fu.w = p();
            fu.out = sc(S_BETA);
            fu.ws = ws.p;
            fu.ws_blocks = ws_blocks;
        }
        int rc = matrix_apply<V>(s, A, p(), 1, 1, nullptr, nullptr, q(), 1, &fu);
        if (rc) return rc;
        ++launch_count;
        if (!fuses) {
            dot_skip<V><<<grid_for(n, 256, 4), 256, 0, s>>>(n, p(), q(), sc(S_BETA), fu.skip, ws.p);
            ++launch_count;
            GKOB200_CHECK_LAUNCH();
        }
        return enqueue_update<false>(s, x);
    }

    int build_graph(V* x, int64_t xs)
    {
        cudaStream_t s = cap_stream;
        if (graph && graph_x == x && graph_xs == xs) return 0;
        if (graph) {
            cudaGraphExecDestroy(graph);
            graph = nullptr;
        }
        const int64_t saved = launch_count;
        cudaGraph_t g = nullptr;
        GKOB200_CUDA(cudaStreamBeginCapture(s, cudaStreamCaptureModeThreadLocal));
        int rc = 0;
        for (int i = 0; i < chunk && rc == 0; ++i) rc = enqueue_iteration(s, x);
        cudaError_t e = cudaStreamEndCapture(s, &g);
        launches_per_iter = (launch_count - saved) / chunk;
        launch_count = saved;
        if (rc) {
            if (g) cudaGraphDestroy(g);
            return rc;
        }
        if (e != cudaSuccess) return static_cast<int>(e);
        e = cudaGraphInstantiate(&graph, g, 0);
        cudaGraphDestroy(g);
        if (e != cudaSuccess) return static_cast<int>(e);
        graph_x = x;
        graph_xs = xs;
        return 0;
    }

    int apply_general(cudaStream_t s, const V* b, int64_t bs, V* x, int64_t xs);

    int apply(cudaStream_t s, const void* b_, int64_t bs, void* x_, int64_t xs) override
    {
        const V* b = static_cast<const V*>(b_);
        V* x = static_cast<V*>(x_);
        launch_count = 0;
        num_iterations = 0;
        if (n == 0) return 0;
        if (!b || !x) return GKOB200_EINVAL;
        if (k != 1 || bs != 1 || xs != 1) return apply_general(s, b, bs, x, xs);
        int rc;
        SolverState* st = state.as<SolverState>();
        // ---- prologue: initialize, r = b - A x, criterion baseline ----------
        init_state<V><<<1, 1, 0, s>>>(st, scalars.as<V>(), k, static_cast<int>(hist.bytes / sizeof(V)));
        if ((rc = gkob200_cg_initialize_f(s, b, bs))) return rc;
        launch_count += 3;
        if ((rc = matrix_apply<V>(s, A, x, 1, 1, sc(S_NEG_ONE), sc(S_ONE), r(), 1, nullptr))) return rc;
        ++launch_count;
        if ((rc = baseline_norm(s, b, bs))) return rc;
        if ((rc = enqueue_update<true>(s, x))) return rc;
        // ---- iterations: graph of `chunk` iterations, host polls the stop flag one
        //      chunk behind so the device never idles ------------------------------
        if ((rc = build_graph(x, xs))) return rc;
        int64_t g = 0;
        bool done = false;
        // after ceil(max_iters / chunk) chunks the Iteration criterion has fired for
        // certain: never enqueue beyond that, just wait for the flag
        const int64_t max_chunks = ceildiv(stop.max_iters, chunk);
        while (!done) {
            if (g >= max_chunks) {
                GKOB200_CUDA(cudaStreamSynchronize(s));
                break;
            }
            GKOB200_CUDA(cudaGraphLaunch(graph, s));
            launch_count += launches_per_iter * chunk;
            const int slot = static_cast<int>(g & 1);
            GKOB200_CUDA(cudaMemcpyAsync(&h_state[slot], st, sizeof(SolverState), cudaMemcpyDeviceToHost, s));
            GKOB200_CUDA(cudaEventRecord(ev[slot], s));
            if (g >= 1) {
                GKOB200_CUDA(cudaEventSynchronize(ev[slot ^ 1]));
                if (h_state[slot ^ 1].stopped) done = true;
            }
            ++g;
        }
        GKOB200_CUDA(cudaStreamSynchronize(s));
        return finish(s);
    }

    int gkob200_cg_initialize_f(cudaStream_t s, const V* b, int64_t bs);
    int baseline_norm(cudaStream_t s, const V* b, int64_t bs);

    int finish(cudaStream_t s)
    {
        SolverState hs;
        GKOB200_CUDA(cudaMemcpyAsync(&h_state[0], state.p, sizeof(SolverState), cudaMemcpyDeviceToHost, s));
        GKOB200_CUDA(cudaMemcpyAsync(stop_status_host.data(), status.p, k, cudaMemcpyDeviceToHost, s));
        GKOB200_CUDA(cudaStreamSynchronize(s));
        hs = h_state[0];
        num_iterations = hs.final_iter;
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
        if ((rc = apply(s, host_b.p, k, host_x.p, k))) return rc;
        GKOB200_CUDA(cudaMemcpyAsync(x_host, host_x.p, bytes, cudaMemcpyDeviceToHost, s));
        GKOB200_CUDA(cudaStreamSynchronize(s));
        launch_count += 3;
        return 0;
    }
};

template <>
int CgSolver<double>::gkob200_cg_initialize_f(cudaStream_t s, const double* b, int64_t bs)
{
    return gkob200_cg_initialize_f64(s, n, k, b, bs, r(), z(), p(), q(), k, sc(S_PREV_RHO), sc(S_RHO),
                                     status.as<uint8_t>());
}
template <>
int CgSolver<float>::gkob200_cg_initialize_f(cudaStream_t s, const float* b, int64_t bs)
{
    return gkob200_cg_initialize_f32(s, n, k, b, bs, r(), z(), p(), q(), k, sc(S_PREV_RHO), sc(S_RHO),
                                     status.as<uint8_t>());
}

inline int norm2_dispatch(cudaStream_t s, int64_t n, int64_t k, const double* x, int64_t xs, double* r, void* ws)
{
    return gkob200_dense_compute_norm2_f64(s, n, k, x, xs, r, ws);
}
inline int norm2_dispatch(cudaStream_t s, int64_t n, int64_t k, const float* x, int64_t xs, float* r, void* ws)
{
    return gkob200_dense_compute_norm2_f32(s, n, k, x, xs, r, ws);
}
inline int fill_dispatch(cudaStream_t s, int64_t n, int64_t k, double* x, int64_t xs, double v)
{
    return gkob200_dense_fill_f64(s, n, k, x, xs, v);
}
inline int fill_dispatch(cudaStream_t s, int64_t n, int64_t k, float* x, int64_t xs, float v)
{
    return gkob200_dense_fill_f32(s, n, k, x, xs, v);
}

// starting_tau of ResidualNormBase [ref: core/stop/residual_norm.cpp:129-186]
template <typename V>
int CgSolver<V>::baseline_norm(cudaStream_t s, const V* b, int64_t bs)
{
    ++launch_count;
    switch (stop.baseline) {
    case GKOB200_STOP_RHS_NORM:
        return norm2_dispatch(s, n, k, b, bs, sc(S_ORIG_TAU), ws.p);
    case GKOB200_STOP_INITIAL_RESNORM:
        return norm2_dispatch(s, n, k, r(), k, sc(S_ORIG_TAU), ws.p);
    case GKOB200_STOP_ABSOLUTE:
        return fill_dispatch(s, 1, k, sc(S_ORIG_TAU), k, V(1));
    default:
        return GKOB200_EINVAL;
    }
}

template <typename V>
int CgSolver<V>::precond_apply(cudaStream_t s, const V* in, V* out)
{
    ++launch_count;
    if (M.kind == GKOB200_PRECOND_JACOBI_BLOCK)
        return typed::jacobi_block_simple_apply(static_cast<V*>(nullptr), s, M.num_blocks,
                                                static_cast<const int32_t*>(M.block_pointers),
                                                static_cast<const V*>(M.blocks), M.block_offset, M.group_offset,
                                                static_cast<int>(M.group_power), n, k, in, k, out, k);
    return GKOB200_EUNSUPPORTED;
}

inline int cg_step_1_dispatch(cudaStream_t s, int64_t n, int64_t k, double* p, const double* z, int64_t st,
                              const double* rho, const double* prev, const uint8_t* stop)
{
    return gkob200_cg_step_1_f64(s, n, k, p, z, st, rho, prev, stop);
}
inline int cg_step_1_dispatch(cudaStream_t s, int64_t n, int64_t k, float* p, const float* z, int64_t st,
                              const float* rho, const float* prev, const uint8_t* stop)
{
    return gkob200_cg_step_1_f32(s, n, k, p, z, st, rho, prev, stop);
}
inline int cg_step_2_dispatch(cudaStream_t s, int64_t n, int64_t k, double* x, int64_t xs, double* r,
                              const double* p, const double* q, int64_t st, const double* beta,
                              const double* rho, const uint8_t* stop)
{
    return gkob200_cg_step_2_f64(s, n, k, x, xs, r, p, q, st, beta, rho, stop);
}
inline int cg_step_2_dispatch(cudaStream_t s, int64_t n, int64_t k, float* x, int64_t xs, float* r, const float* p,
                              const float* q, int64_t st, const float* beta, const float* rho,
                              const uint8_t* stop)
{
    return gkob200_cg_step_2_f32(s, n, k, x, xs, r, p, q, st, beta, rho, stop);
}
inline int dot_dispatch(cudaStream_t s, int64_t n, int64_t k, const double* x, int64_t xs, const double* y,
                        int64_t ys, double* r, void* ws)
{
    return gkob200_dense_compute_dot_f64(s, n, k, x, xs, y, ys, r, ws);
}
inline int dot_dispatch(cudaStream_t s, int64_t n, int64_t k, const float* x, int64_t xs, const float* y,
                        int64_t ys, float* r, void* ws)
{
    return gkob200_dense_compute_dot_f32(s, n, k, x, xs, y, ys, r, ws);
}
inline int jac_dispatch(cudaStream_t s, int64_t n, int64_t k, const double* inv, const double* b, int64_t bs,
                        double* x, int64_t xs)
{
    return gkob200_jacobi_simple_scalar_apply_f64(s, n, k, inv, b, bs, x, xs);
}
inline int jac_dispatch(cudaStream_t s, int64_t n, int64_t k, const float* inv, const float* b, int64_t bs,
                        float* x, int64_t xs)
{
    return gkob200_jacobi_simple_scalar_apply_f32(s, n, k, inv, b, bs, x, xs);
}
inline int copy_dispatch(cudaStream_t s, int64_t n, int64_t k, const double* x, int64_t xs, double* y, int64_t ys)
{
    return gkob200_dense_copy_f64(s, n, k, x, xs, y, ys);
}
inline int copy_dispatch(cudaStream_t s, int64_t n, int64_t k, const float* x, int64_t xs, float* y, int64_t ys)
{
    return gkob200_dense_copy_f32(s, n, k, x, xs, y, ys);
}

// General path (multi-RHS and/or strided b, x): the reference sequence kernel by
// kernel (core/solver/cg.cpp:137-193), criterion on device, host polls every chunk.
template <typename V>
int CgSolver<V>::apply_general(cudaStream_t s, const V* b, int64_t bs, V* x, int64_t xs)
{
    int rc;
    SolverState* st = state.as<SolverState>();
    uint8_t* stat = status.as<uint8_t>();
    init_state<V><<<1, 1, 0, s>>>(st, scalars.as<V>(), k, static_cast<int>(hist.bytes / sizeof(V)));
    if ((rc = gkob200_cg_initialize_f(s, b, bs))) return rc;
    if ((rc = matrix_apply<V>(s, A, x, xs, k, sc(S_NEG_ONE), sc(S_ONE), r(), k, nullptr))) return rc;
    if ((rc = baseline_norm(s, b, bs))) return rc;
    launch_count += 4;
    int64_t it = 0;
    while (true) {
        // z = M^-1 r
        if (M.kind == GKOB200_PRECOND_NONE)
            rc = copy_dispatch(s, n, k, r(), k, z(), k);
        else if (M.kind == GKOB200_PRECOND_JACOBI_SCALAR)
            rc = jac_dispatch(s, n, k, static_cast<const V*>(M.inv_diag), r(), k, z(), k);
        else
            rc = precond_apply(s, r(), z());
        if (rc) return rc;
        if ((rc = dot_dispatch(s, n, k, r(), k, z(), k, sc(S_RHO), ws.p))) return rc;
        if ((rc = norm2_dispatch(s, n, k, r(), k, sc(S_TAU), ws.p))) return rc;
        cg_criterion_k<V><<<1, 1, 0, s>>>(st, k, sc(S_TAU), sc(S_ORIG_TAU), static_cast<V>(stop.reduction_factor),
                                          stop.max_iters, stat, hist.as<V>());
        launch_count += 4;
        if (it % chunk == 0 || it >= stop.max_iters) {
            GKOB200_CUDA(cudaMemcpyAsync(&h_state[0], st, sizeof(SolverState), cudaMemcpyDeviceToHost, s));
            GKOB200_CUDA(cudaStreamSynchronize(s));
            if (h_state[0].stopped) break;
        }
        if ((rc = cg_step_1_dispatch(s, n, k, p(), z(), k, sc(S_RHO), sc(S_PREV_RHO), stat))) return rc;
        if ((rc = matrix_apply<V>(s, A, p(), k, k, nullptr, nullptr, q(), k, nullptr))) return rc;
        if ((rc = dot_dispatch(s, n, k, p(), k, q(), k, sc(S_BETA), ws.p))) return rc;
        if ((rc = cg_step_2_dispatch(s, n, k, x, xs, r(), p(), q(), k, sc(S_BETA), sc(S_RHO), stat))) return rc;
        // swap(prev_rho, rho): rho is recomputed at the top, so a copy is equivalent
        if ((rc = copy_dispatch(s, 1, k, sc(S_RHO), k, sc(S_PREV_RHO), k))) return rc;
        launch_count += 5;
        ++it;
    }
    return finish(s);
}

template <typename V>
gkob200_solver* make_cg(const gkob200_matrix* A, const gkob200_precond* M, const gkob200_stop* stop, int64_t nrhs,
                        int* rc)
{
    auto* s = new CgSolver<V>();
    s->A = *A;
    if (M)
        s->M = *M;
    else {
        s->M = gkob200_precond{};
        s->M.kind = GKOB200_PRECOND_NONE;
    }
    s->stop = *stop;
    s->k = nrhs;
    s->nrhs = nrhs;
    *rc = s->init();
    if (*rc) {
        delete s;
        return nullptr;
    }
    return s;
}

}  // namespace

gkob200_solver* make_cg_f64(const gkob200_matrix* A, const gkob200_precond* M, const gkob200_stop* st, int64_t nrhs, int* rc)
{
    return make_cg<double>(A, M, st, nrhs, rc);
}
gkob200_solver* make_cg_f32(const gkob200_matrix* A, const gkob200_precond* M, const gkob200_stop* st, int64_t nrhs, int* rc)
{
    return make_cg<float>(A, M, st, nrhs, rc);
}

}  // namespace gkob200
