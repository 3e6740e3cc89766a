// solver_krylov.cu — BiCGSTAB and GMRES solver objects.
//
// The reference loops (core/solver/bicgstab.cpp:109-241, core/solver/gmres.cpp:139-369)
// issued kernel by kernel on the GPU with the stopping criterion evaluated ON THE DEVICE
// (solver_common.cuh): nothing is copied to the host per iteration (the reference does two
// blocking 1-byte copies per criterion check, i.e. four per BiCGSTAB iteration); the host
// polls the device-side status every `check_every` iterations and — for GMRES — at every
// restart boundary.  Kernels issued after a column has stopped are no-ops through the
// reference's own has_stopped() guards, so the reported iteration count is exact.
// Any number of right-hand sides; BiCGSTAB with one right-hand side and an identity / scalar
// Jacobi preconditioner runs the fused iteration below.
#include "solver_common.cuh"

namespace gkob200 {
namespace {

// ------------------------------------------------------------------------------
// Fused BiCGSTAB iteration for one right-hand side and an identity or scalar-Jacobi
// preconditioner: 5 kernels and 16 vector passes per iteration instead of 18 and 27.
//   bi_step1   p = r + (rho/prev_rho * alpha/omega) (p - omega v)          [+ y = D^-1 p]
//   SpMV       v = A y            with beta = rr.v reduced inside the SpMV kernel
//   bi_step2   alpha = rho/beta; s = r - alpha v; ||s||; criterion check    [+ z = D^-1 s]
//   SpMV       t = A z            with gamma = s.t and t.t reduced inside the SpMV kernel
//   bi_step3   omega = gamma/(t.t); x += alpha y + omega z; r = s - omega t;
//              next rho = rr.r and ||r||; ++iter; criterion check
// Same arithmetic per entry as bicgstab_step_1/2/3 (krylov_kernels.cu); with the identity
// preconditioner y aliases p and z aliases s (the reference copies them).
enum { BI_RHO = 0, BI_PREV_RHO, BI_ALPHA, BI_BETA, BI_GAMMA, BI_OMEGA, BI_TT, BI_COUNT };

template <typename V>
struct BiParams {
    int64_t n;
    V* x;
    int64_t xs;
    V *r, *rr, *p, *v, *s, *t, *y, *z;
    const V* inv_diag;  // scalar Jacobi or nullptr
    V* sc;
    SolverState* st;
    uint8_t* stop_status;
    V* hist;
    V* tau;
    const V* orig_tau;
    V factor;
    int64_t max_iters;
    void* ws;
};

// rho = rr.r, tau = ||r||, first criterion check (start of iteration 0)
template <typename V>
__global__ void __launch_bounds__(256) bi_top(BiParams<V> P)
{
    V acc[2] = {V(0), V(0)};
    const int64_t step = static_cast<int64_t>(gridDim.x) * blockDim.x;
    for (int64_t i = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; i < P.n; i += step) {
        const V ri = P.r[i];
        acc[0] += P.rr[i] * ri;
        acc[1] += ri * ri;
    }
    BiParams<V> Q = P;
    grid_reduce<2>(acc, ws_partials<V>(P.ws), ws_ticket(P.ws), [Q](V(&tot)[2]) {
        Q.sc[BI_RHO] = tot[0];
        Q.tau[0] = sqrt_rn(tot[1]);
        criterion_check(Q.st, 1, Q.tau, Q.orig_tau, Q.factor, Q.max_iters, true, Q.stop_status, Q.hist, true);
    });
}

template <typename V, bool Jacobi>
__global__ void __launch_bounds__(256) bi_step1(BiParams<V> P)
{
    if (P.st->stopped) return;
    const V om = P.sc[BI_OMEGA], prev = P.sc[BI_PREV_RHO];
    const bool upd = mul_rn(prev, om) != V(0);
    const V tmp = upd ? div_rn(mul_rn(div_rn(P.sc[BI_RHO], prev), P.sc[BI_ALPHA]), om) : V(0);
    const int64_t step = static_cast<int64_t>(gridDim.x) * blockDim.x;
    for (int64_t i = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; i < P.n; i += step) {
        const V ri = P.r[i];
        const V pi = upd ? add_rn(ri, mul_rn(tmp, sub_rn(P.p[i], mul_rn(om, P.v[i])))) : ri;
        P.p[i] = pi;
        if (Jacobi) P.y[i] = mul_rn(pi, P.inv_diag[i]);
    }
}

template <typename V, bool Jacobi>
__global__ void __launch_bounds__(256) bi_step2(BiParams<V> P)
{
    if (P.st->stopped) return;
    const V be = P.sc[BI_BETA];
    const V a = be != V(0) ? div_rn(P.sc[BI_RHO], be) : V(0);
    V acc[1] = {V(0)};
    const int64_t step = static_cast<int64_t>(gridDim.x) * blockDim.x;
    for (int64_t i = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; i < P.n; i += step) {
        const V si = be != V(0) ? sub_rn(P.r[i], mul_rn(a, P.v[i])) : P.r[i];
        P.s[i] = si;
        if (Jacobi) P.z[i] = mul_rn(si, P.inv_diag[i]);
        acc[0] += si * si;
    }
    BiParams<V> Q = P;
    grid_reduce<1>(acc, ws_partials<V>(P.ws), ws_ticket(P.ws), [Q, a](V(&tot)[1]) {
        Q.sc[BI_ALPHA] = a;
        Q.tau[0] = sqrt_rn(tot[0]);
        // mid-iteration check: not finalized, x += alpha y follows once the solver has stopped
        criterion_check(Q.st, 1, Q.tau, Q.orig_tau, Q.factor, Q.max_iters, false, Q.stop_status, Q.hist, false);
    });
}

template <typename V>
__global__ void __launch_bounds__(256) bi_step3(BiParams<V> P)
{
    if (P.st->stopped) return;
    const V tt = P.sc[BI_TT];
    const V om = tt != V(0) ? div_rn(P.sc[BI_GAMMA], tt) : V(0);
    const V a = P.sc[BI_ALPHA];
    V acc[2] = {V(0), V(0)};
    const int64_t step = static_cast<int64_t>(gridDim.x) * blockDim.x;
    for (int64_t i = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; i < P.n; i += step) {
        P.x[i * P.xs] = add_rn(P.x[i * P.xs], add_rn(mul_rn(a, P.y[i]), mul_rn(om, P.z[i])));
        const V ri = sub_rn(P.s[i], mul_rn(om, P.t[i]));
        P.r[i] = ri;
        acc[0] += P.rr[i] * ri;
        acc[1] += ri * ri;
    }
    BiParams<V> Q = P;
    grid_reduce<2>(acc, ws_partials<V>(P.ws), ws_ticket(P.ws), [Q, om](V(&tot)[2]) {
        Q.sc[BI_OMEGA] = om;
        Q.sc[BI_PREV_RHO] = Q.sc[BI_RHO];
        Q.sc[BI_RHO] = tot[0];
        Q.tau[0] = sqrt_rn(tot[1]);
        criterion_check(Q.st, 1, Q.tau, Q.orig_tau, Q.factor, Q.max_iters, true, Q.stop_status, Q.hist, true);
    });
}

// formats without an in-kernel reduction: out[0] = a.b, out[1] = b.b (out1 may be null)
template <typename V>
__global__ void __launch_bounds__(256) bi_dots(int64_t n, const V* __restrict__ a, const V* __restrict__ b, V* out0,
                                               V* out1, const int* skip, void* ws)
{
    if (*skip) return;
    V acc[2] = {V(0), V(0)};
    const int64_t step = static_cast<int64_t>(gridDim.x) * blockDim.x;
    for (int64_t i = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; i < n; i += step) {
        const V bi = b[i];
        acc[0] += a[i] * bi;
        acc[1] += bi * bi;
    }
    grid_reduce<2>(acc, ws_partials<V>(ws), ws_ticket(ws), [out0, out1](V(&tot)[2]) {
        out0[0] = tot[0];
        if (out1) out1[0] = tot[1];
    });
}

// ------------------------------------------------------------------------------
template <typename V>
struct BicgstabSolver : SolverBase<V> {
    using B = SolverBase<V>;
    using B::A; using B::M; using B::k; using B::n; using B::stop; using B::ws; using B::launch_count;
    DevBuf vecs, scal;
    V* vec(int i) { return vecs.as<V>() + static_cast<int64_t>(i) * n * k; }
    V* sc(int i) { return scal.as<V>() + static_cast<int64_t>(i) * k; }

    int init()
    {
        int rc = this->init_base();
        if (rc) return rc;
        if ((rc = vecs.alloc(static_cast<size_t>(n) * k * 8 * sizeof(V)))) return rc;
        // fused path: the SpMV kernels leave up to two partials per CTA (128 rows each)
        ws_blocks = 2 * (ceildiv(n, 128) + 1);
        if (ws_blocks < kReduceMaxBlocks) ws_blocks = kReduceMaxBlocks;
        if ((rc = ws.alloc(reduce_ws_bytes(ws_blocks)))) return rc;
        return scal.alloc(static_cast<size_t>(k > BI_COUNT ? k : BI_COUNT) * 6 * sizeof(V));
    }
    int64_t ws_blocks = 0;

    bool fusable() const
    {
        return k == 1 && (M.kind == GKOB200_PRECOND_NONE || M.kind == GKOB200_PRECOND_JACOBI_SCALAR);
    }

    // one SpMV of the fused iteration: out = A in, d0 = w.out and (optionally) d1 = out.out
    int fused_spmv(cudaStream_t s, const V* in, V* out, const V* w, V* d0, V* d1)
    {
        SpmvFusion<V> fu;
        fu.skip = &this->st()->stopped;
        const bool fuses = matrix_apply_fuses_dot(A, 1) && A.format != GKOB200_FMT_CSR_ROWS;
        if (fuses) {
            fu.w = w;
            fu.out = d0;
            fu.out_sq = d1;
            fu.ws = ws.p;
            fu.ws_blocks = ws_blocks;
        }
        int rc = matrix_apply<V>(s, A, in, 1, 1, nullptr, nullptr, out, 1, &fu);
        if (rc) return rc;
        launch_count += fuses ? 2 : 1;
        if (!fuses) {
            bi_dots<V><<<grid_for(n, 256, 4), 256, 0, s>>>(n, w, out, d0, d1, fu.skip, ws.p);
            ++launch_count;
            GKOB200_CHECK_LAUNCH();
        }
        return 0;
    }

    int apply_fused(cudaStream_t s, const V* b, int64_t bs, V* x, int64_t xs)
    {
        V* tag = B::tag();
        const bool jac = M.kind == GKOB200_PRECOND_JACOBI_SCALAR;
        BiParams<V> P;
        P.n = n;
        P.x = x;
        P.xs = xs;
        P.r = vec(0);
        P.z = jac ? vec(1) : vec(4);
        P.y = jac ? vec(2) : vec(6);
        P.v = vec(3);
        P.s = vec(4);
        P.t = vec(5);
        P.p = vec(6);
        P.rr = vec(7);
        P.inv_diag = jac ? static_cast<const V*>(M.inv_diag) : nullptr;
        P.sc = scal.as<V>();
        P.st = this->st();
        P.stop_status = this->stat();
        P.hist = this->hist.template as<V>();
        P.tau = this->tau();
        P.orig_tau = this->orig_tau();
        P.factor = static_cast<V>(stop.reduction_factor);
        P.max_iters = stop.max_iters;
        P.ws = ws.p;
        int rc;
        if ((rc = this->reset_state(s))) return rc;
        // r = b - A x, rr = r, p = v = 0, scalars: rho = prev_rho = alpha = omega = 1 (bicgstab_initialize)
        if ((rc = typed::bicgstab_initialize(tag, s, n, k, b, bs, P.r, P.rr, vec(2), P.s, P.t, vec(1), P.v, P.p, k,
                                             P.sc + BI_PREV_RHO, P.sc + BI_RHO, P.sc + BI_ALPHA, P.sc + BI_BETA,
                                             P.sc + BI_GAMMA, P.sc + BI_OMEGA, this->stat())))
            return rc;
        if ((rc = matrix_apply<V>(s, A, x, xs, k, this->neg_one(), this->one(), P.r, k, nullptr))) return rc;
        if ((rc = this->baseline_norm(s, b, bs, P.r, k))) return rc;
        if ((rc = typed::dense_copy(tag, s, n, k, P.r, k, P.rr, k))) return rc;
        const int grid = grid_for(n, 256, 6);
        bi_top<V><<<grid, 256, 0, s>>>(P);
        GKOB200_CHECK_LAUNCH();
        launch_count += 5;
        bool stopped = false;
        for (int64_t it = 0;; ++it) {
            if (it % this->chunk == 0 || it > stop.max_iters) {
                if ((rc = this->poll(s, &stopped))) return rc;
                if (stopped) break;
            }
            if (jac)
                bi_step1<V, true><<<grid, 256, 0, s>>>(P);
            else
                bi_step1<V, false><<<grid, 256, 0, s>>>(P);
            GKOB200_CHECK_LAUNCH();
            if ((rc = fused_spmv(s, P.y, P.v, P.rr, P.sc + BI_BETA, nullptr))) return rc;
            if (jac)
                bi_step2<V, true><<<grid, 256, 0, s>>>(P);
            else
                bi_step2<V, false><<<grid, 256, 0, s>>>(P);
            GKOB200_CHECK_LAUNCH();
            if ((rc = fused_spmv(s, P.z, P.t, P.s, P.sc + BI_GAMMA, P.sc + BI_TT))) return rc;
            bi_step3<V><<<grid, 256, 0, s>>>(P);
            GKOB200_CHECK_LAUNCH();
            launch_count += 3;
        }
        // stopped by the mid-iteration check: x += alpha y, then mark finalized
        if ((rc = typed::bicgstab_finalize(tag, s, n, k, x, xs, P.y, k, P.sc + BI_ALPHA, this->stat()))) return rc;
        launch_count += 2;
        return this->finish(s);
    }

    int apply(cudaStream_t s, const void* b_, int64_t bs, void* x_, int64_t xs) override
    {
        const V* b = static_cast<const V*>(b_);
        V* x = static_cast<V*>(x_);
        V* tag = B::tag();
        launch_count = 0;
        this->num_iterations = 0;
        if (n == 0) return 0;
        if (!b || !x) return GKOB200_EINVAL;
        if (fusable()) return apply_fused(s, b, bs, x, xs);
        V *r = vec(0), *z = vec(1), *y = vec(2), *v = vec(3), *s_ = vec(4), *t = vec(5), *p = vec(6), *rr = vec(7);
        V *alpha = sc(0), *beta = sc(1), *gamma = sc(2), *prev_rho = sc(3), *rho = sc(4), *omega = sc(5);
        uint8_t* stat = this->stat();
        int rc;
        if ((rc = this->reset_state(s))) return rc;
        if ((rc = typed::bicgstab_initialize(tag, s, n, k, b, bs, r, rr, y, s_, t, z, v, p, k, prev_rho, rho, alpha,
                                             beta, gamma, omega, stat)))
            return rc;
        if ((rc = matrix_apply<V>(s, A, x, xs, k, this->neg_one(), this->one(), r, k, nullptr))) return rc;
        if ((rc = this->baseline_norm(s, b, bs, r, k))) return rc;
        if ((rc = typed::dense_copy(tag, s, n, k, r, k, rr, k))) return rc;
        launch_count += 4;
        int64_t it = 0;
        bool stopped = false;
        while (true) {
            // rho = rr . r ; ++iter ; first check on ||r|| (finalized)
            if ((rc = typed::dense_compute_dot(tag, s, n, k, rr, k, r, k, rho, ws.p))) return rc;
            if ((rc = typed::dense_compute_norm2(tag, s, n, k, r, k, this->tau(), ws.p))) return rc;
            if ((rc = this->check(s, this->tau(), true, true))) return rc;
            launch_count += 2;
            if (it % this->chunk == 0 || it >= stop.max_iters) {
                if ((rc = this->poll(s, &stopped))) return rc;
                if (stopped) break;
            }
            if ((rc = typed::bicgstab_step_1(tag, s, n, k, r, p, v, k, rho, prev_rho, alpha, omega, stat))) return rc;
            if ((rc = this->precond_apply(s, p, k, y, k))) return rc;
            if ((rc = matrix_apply<V>(s, A, y, k, k, nullptr, nullptr, v, k, nullptr))) return rc;
            if ((rc = typed::dense_compute_dot(tag, s, n, k, rr, k, v, k, beta, ws.p))) return rc;
            if ((rc = typed::bicgstab_step_2(tag, s, n, k, r, s_, v, k, rho, alpha, beta, stat))) return rc;
            // second check on ||s||, not finalized; x += alpha y for the columns it stopped
            if ((rc = typed::dense_compute_norm2(tag, s, n, k, s_, k, this->tau(), ws.p))) return rc;
            if ((rc = this->check(s, this->tau(), false, false))) return rc;
            if ((rc = typed::bicgstab_finalize(tag, s, n, k, x, xs, y, k, alpha, stat))) return rc;
            if ((rc = this->precond_apply(s, s_, k, z, k))) return rc;
            if ((rc = matrix_apply<V>(s, A, z, k, k, nullptr, nullptr, t, k, nullptr))) return rc;
            if ((rc = typed::dense_compute_dot(tag, s, n, k, s_, k, t, k, gamma, ws.p))) return rc;
            if ((rc = typed::dense_compute_dot(tag, s, n, k, t, k, t, k, beta, ws.p))) return rc;
            if ((rc = typed::bicgstab_step_3(tag, s, n, k, x, xs, r, s_, t, y, z, k, alpha, beta, gamma, omega, stat)))
                return rc;
            // swap(prev_rho, rho): rho is recomputed at the top, a copy is equivalent
            if ((rc = typed::dense_copy(tag, s, int64_t(1), k, rho, k, prev_rho, k))) return rc;
            launch_count += 12;
            ++it;
        }
        return this->finish(s);
    }
};

// ------------------------------------------------------------------------------
// Fused modified Gram-Schmidt step of the GMRES Arnoldi process (one right-hand side):
//   w -= h_i v_i   and, in the same pass,   h_{i+1} = w . v_{i+1}      (Last: ||w|| instead)
// — 4 vector passes per basis vector instead of the 5 of compute_dot + sub_scaled, with the
// first dot w . v_0 reduced inside the SpMV kernel that produces w.  Same arithmetic and the
// same (sequential) dependency chain as the reference loop (core/solver/gmres.cpp:299-316).
template <typename V, bool Last>
__global__ void __launch_bounds__(256) gmres_mgs_step(int64_t n, V* __restrict__ w, const V* __restrict__ vi,
                                                      const V* __restrict__ vnext, const V* h_i, V* out,
                                                      const int* skip, void* ws)
{
    if (*skip) return;
    const V h = h_i[0];
    V acc[1] = {V(0)};
    const int64_t step = static_cast<int64_t>(gridDim.x) * blockDim.x;
    for (int64_t i = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; i < n; i += step) {
        const V wi = sub_rn(w[i], mul_rn(h, vi[i]));
        w[i] = wi;
        acc[0] += Last ? wi * wi : wi * vnext[i];
    }
    grid_reduce<1>(acc, ws_partials<V>(ws), ws_ticket(ws), [out](V(&tot)[1]) { out[0] = Last ? sqrt_rn(tot[0]) : tot[0]; });
}

// ------------------------------------------------------------------------------
template <typename V>
struct GmresSolver : SolverBase<V> {
    using B = SolverBase<V>;
    using B::A; using B::M; using B::k; using B::n; using B::stop; using B::ws; using B::launch_count;
    int64_t m = 100;  // krylov_dim
    DevBuf bases, vecs, small, fin;

    int init()
    {
        int rc = this->init_base();
        if (rc) return rc;
        if ((rc = bases.alloc(static_cast<size_t>(n) * k * (m + 1) * sizeof(V)))) return rc;
        if ((rc = vecs.alloc(static_cast<size_t>(n) * k * 4 * sizeof(V)))) return rc;
        // hessenberg (m+1) x m k | givens_sin m x k | givens_cos m x k | rnc (m+1) x k | y m x k | residual_norm k
        const size_t cnt = static_cast<size_t>((m + 1) * m * k + 2 * m * k + (m + 1) * k + m * k + k);
        if ((rc = small.alloc(cnt * sizeof(V)))) return rc;
        // fused Arnoldi: the SpMV kernel leaves one partial per CTA (128 rows each)
        ws_blocks = ceildiv(n, 128) + 1;
        if (ws_blocks < kReduceMaxBlocks) ws_blocks = kReduceMaxBlocks;
        if ((rc = ws.alloc(reduce_ws_bytes(ws_blocks)))) return rc;
        return fin.alloc(static_cast<size_t>(k) * sizeof(uint64_t));
    }
    int64_t ws_blocks = 0;

    // next_k = A precvec; modified Gram-Schmidt against basis 0..j; hnorm = ||next_k||  (k == 1)
    int arnoldi_fused(cudaStream_t s, const V* precvec, V* kb, int64_t j, V* hess_iter, int64_t hs)
    {
        V* next_k = kb + (j + 1) * n;
        SpmvFusion<V> fu;
        fu.skip = &this->st()->stopped;
        const bool fuses = matrix_apply_fuses_dot(A, 1) && A.format != GKOB200_FMT_CSR_ROWS;
        if (fuses) {
            fu.w = kb;
            fu.out = hess_iter;
            fu.ws = ws.p;
            fu.ws_blocks = ws_blocks;
        }
        int rc = matrix_apply<V>(s, A, precvec, 1, 1, nullptr, nullptr, next_k, 1, &fu);
        if (rc) return rc;
        launch_count += fuses ? 2 : 1;
        if (!fuses) {
            bi_dots<V><<<grid_for(n, 256, 4), 256, 0, s>>>(n, kb, next_k, hess_iter, static_cast<V*>(nullptr), fu.skip, ws.p);
            ++launch_count;
            GKOB200_CHECK_LAUNCH();
        }
        const int grid = grid_for(n, 256, 6);
        for (int64_t i = 0; i <= j; ++i) {
            const V* h = hess_iter + i * hs;
            if (i < j)
                gmres_mgs_step<V, false><<<grid, 256, 0, s>>>(n, next_k, kb + i * n, kb + (i + 1) * n, h,
                                                              hess_iter + (i + 1) * hs, fu.skip, ws.p);
            else
                gmres_mgs_step<V, true><<<grid, 256, 0, s>>>(n, next_k, kb + i * n, static_cast<const V*>(nullptr), h,
                                                             hess_iter + (j + 1) * hs, fu.skip, ws.p);
            GKOB200_CHECK_LAUNCH();
            ++launch_count;
        }
        return 0;
    }

    int apply(cudaStream_t s, const void* b_, int64_t bs, void* x_, int64_t xs) override
    {
        const V* b = static_cast<const V*>(b_);
        V* x = static_cast<V*>(x_);
        V* tag = B::tag();
        launch_count = 0;
        this->num_iterations = 0;
        if (n == 0) return 0;
        if (!b || !x) return GKOB200_EINVAL;
        V* residual = vecs.as<V>();
        V* precvec = residual + n * k;
        V* before = precvec + n * k;
        V* after = before + n * k;
        V* kb = bases.as<V>();
        V* hess = small.as<V>();
        const int64_t hs = m * k;
        V* gsin = hess + (m + 1) * hs;
        V* gcos = gsin + m * k;
        V* rnc = gcos + m * k;
        V* yv = rnc + (m + 1) * k;
        V* rnorm = yv + m * k;
        uint64_t* fin_it = fin.as<uint64_t>();
        uint8_t* stat = this->stat();
        int rc;
        if ((rc = this->reset_state(s))) return rc;
        if ((rc = typed::gmres_initialize(tag, s, n, k, m, b, bs, residual, k, gsin, gcos, stat))) return rc;
        if ((rc = matrix_apply<V>(s, A, x, xs, k, this->neg_one(), this->one(), residual, k, nullptr))) return rc;
        if ((rc = typed::dense_compute_norm2(tag, s, n, k, residual, k, rnorm, ws.p))) return rc;
        if ((rc = typed::gmres_restart(tag, s, n, k, residual, k, rnorm, rnc, kb, fin_it))) return rc;
        if ((rc = this->baseline_norm(s, b, bs, residual, k))) return rc;
        launch_count += 8;
        int64_t restart_iter = 0, since_poll = 0;
        bool stopped = false;
        auto update_x = [&]() -> int {
            // y = H \ rnc ; before = V y ; x += M^-1 before   (gmres.cpp:245-256, 350-369)
            int r2;
            if ((r2 = typed::gmres_solve_krylov(tag, s, k, rnc, hess, hs, yv, fin_it, stat))) return r2;
            if ((r2 = typed::gmres_multi_axpy(tag, s, n, k, kb, yv, before, k, fin_it, stat))) return r2;
            if ((r2 = this->precond_apply(s, before, k, after, k))) return r2;
            launch_count += 4;
            return typed::dense_add_scaled(tag, s, n, k, this->one(), int64_t(1), after, k, x, xs);
        };
        while (true) {
            // ++total_iter; check on the (implicit) residual norm, NOT finalized
            if ((rc = this->check(s, rnorm, false, true))) return rc;
            ++since_poll;
            if (restart_iter == m || since_poll >= this->chunk) {
                // always poll before a restart: the x-update must run exactly once
                if ((rc = this->poll(s, &stopped))) return rc;
                since_poll = 0;
                if (stopped) break;
            }
            if (restart_iter == m) {
                if ((rc = update_x())) return rc;
                if ((rc = typed::dense_copy(tag, s, n, k, b, bs, residual, k))) return rc;
                if ((rc = matrix_apply<V>(s, A, x, xs, k, this->neg_one(), this->one(), residual, k, nullptr))) return rc;
                if ((rc = typed::dense_compute_norm2(tag, s, n, k, residual, k, rnorm, ws.p))) return rc;
                if ((rc = typed::gmres_restart(tag, s, n, k, residual, k, rnorm, rnc, kb, fin_it))) return rc;
                launch_count += 5;
                restart_iter = 0;
            }
            V* this_k = kb + restart_iter * n * k;
            V* next_k = kb + (restart_iter + 1) * n * k;
            V* hess_iter = hess + restart_iter * k;
            if ((rc = this->precond_apply(s, this_k, k, precvec, k))) return rc;
            V* hnorm = hess_iter + (restart_iter + 1) * hs;
            if (k == 1) {
                if ((rc = arnoldi_fused(s, precvec, kb, restart_iter, hess_iter, hs))) return rc;
            } else {
                if ((rc = matrix_apply<V>(s, A, precvec, k, k, nullptr, nullptr, next_k, k, nullptr))) return rc;
                ++launch_count;
                // modified Gram-Schmidt against all previous basis vectors
                for (int64_t i = 0; i <= restart_iter; ++i) {
                    V* h = hess_iter + i * hs;
                    V* basis = kb + i * n * k;
                    if ((rc = typed::dense_compute_dot(tag, s, n, k, next_k, k, basis, k, h, ws.p))) return rc;
                    if ((rc = typed::dense_sub_scaled(tag, s, n, k, h, k, basis, k, next_k, k))) return rc;
                    launch_count += 2;
                }
                if ((rc = typed::dense_compute_norm2(tag, s, n, k, next_k, k, hnorm, ws.p))) return rc;
                ++launch_count;
            }
            if ((rc = typed::dense_inv_scale(tag, s, n, k, hnorm, k, next_k, k))) return rc;
            if ((rc = typed::gmres_hessenberg_qr(tag, s, k, gsin, gcos, rnorm, rnc, hess_iter, hs, restart_iter, fin_it,
                                                 stat)))
                return rc;
            launch_count += 2;
            ++restart_iter;
        }
        if ((rc = update_x())) return rc;
        return this->finish(s);
    }
};

// ------------------------------------------------------------------------------
// FCG and CGS: the reference loops (core/solver/fcg.cpp:104-196,
// core/solver/cgs.cpp:104-212) kernel by kernel, with the criterion on the device like the
// BiCGSTAB / GMRES general paths.  Any number of right-hand sides.
template <typename V>
struct FcgSolver : SolverBase<V> {
    using B = SolverBase<V>;
    using B::A; using B::M; using B::k; using B::n; using B::stop; using B::ws; using B::launch_count;
    DevBuf vecs, scal;
    V* vec(int i) { return vecs.as<V>() + static_cast<int64_t>(i) * n * k; }
    V* sc(int i) { return scal.as<V>() + static_cast<int64_t>(i) * k; }

    int init()
    {
        int rc = this->init_base();
        if (rc) return rc;
        if ((rc = vecs.alloc(static_cast<size_t>(n) * k * 5 * sizeof(V)))) return rc;
        return scal.alloc(static_cast<size_t>(k) * 4 * sizeof(V));
    }

    int apply(cudaStream_t s, const void* b_, int64_t bs, void* x_, int64_t xs) override
    {
        const V* b = static_cast<const V*>(b_);
        V* x = static_cast<V*>(x_);
        V* tag = B::tag();
        launch_count = 0;
        this->num_iterations = 0;
        if (n == 0) return 0;
        if (!b || !x) return GKOB200_EINVAL;
        V *r = vec(0), *z = vec(1), *p = vec(2), *q = vec(3), *t = vec(4);
        V *beta = sc(0), *prev_rho = sc(1), *rho = sc(2), *rho_t = sc(3);
        uint8_t* stat = this->stat();
        int rc;
        if ((rc = this->reset_state(s))) return rc;
        if ((rc = typed::fcg_initialize(tag, s, n, k, b, bs, r, z, p, q, t, k, prev_rho, rho, rho_t, stat))) return rc;
        if ((rc = matrix_apply<V>(s, A, x, xs, k, this->neg_one(), this->one(), r, k, nullptr))) return rc;
        if ((rc = this->baseline_norm(s, b, bs, r, k))) return rc;
        launch_count += 3;
        int64_t it = 0;
        bool stopped = false;
        while (true) {
            if ((rc = this->precond_apply(s, r, k, z, k))) return rc;
            if ((rc = typed::dense_compute_dot(tag, s, n, k, r, k, z, k, rho, ws.p))) return rc;
            if ((rc = typed::dense_compute_dot(tag, s, n, k, t, k, z, k, rho_t, ws.p))) return rc;
            if ((rc = typed::dense_compute_norm2(tag, s, n, k, r, k, this->tau(), ws.p))) return rc;
            if ((rc = this->check(s, this->tau(), true, true))) return rc;
            launch_count += 3;
            if (it % this->chunk == 0 || it >= stop.max_iters) {
                if ((rc = this->poll(s, &stopped))) return rc;
                if (stopped) break;
            }
            if ((rc = typed::fcg_step_1(tag, s, n, k, p, z, k, rho_t, prev_rho, stat))) return rc;
            if ((rc = matrix_apply<V>(s, A, p, k, k, nullptr, nullptr, q, k, nullptr))) return rc;
            if ((rc = typed::dense_compute_dot(tag, s, n, k, p, k, q, k, beta, ws.p))) return rc;
            if ((rc = typed::fcg_step_2(tag, s, n, k, x, xs, r, t, p, q, k, beta, rho, stat))) return rc;
            // swap(prev_rho, rho): rho is recomputed at the top, a copy is equivalent
            if ((rc = typed::dense_copy(tag, s, int64_t(1), k, rho, k, prev_rho, k))) return rc;
            launch_count += 5;
            ++it;
        }
        return this->finish(s);
    }
};

template <typename V>
struct CgsSolver : SolverBase<V> {
    using B = SolverBase<V>;
    using B::A; using B::M; using B::k; using B::n; using B::stop; using B::ws; using B::launch_count;
    DevBuf vecs, scal;
    V* vec(int i) { return vecs.as<V>() + static_cast<int64_t>(i) * n * k; }
    V* sc(int i) { return scal.as<V>() + static_cast<int64_t>(i) * k; }

    int init()
    {
        int rc = this->init_base();
        if (rc) return rc;
        if ((rc = vecs.alloc(static_cast<size_t>(n) * k * 8 * sizeof(V)))) return rc;
        return scal.alloc(static_cast<size_t>(k) * 5 * sizeof(V));
    }

    int apply(cudaStream_t s, const void* b_, int64_t bs, void* x_, int64_t xs) override
    {
        const V* b = static_cast<const V*>(b_);
        V* x = static_cast<V*>(x_);
        V* tag = B::tag();
        launch_count = 0;
        this->num_iterations = 0;
        if (n == 0) return 0;
        if (!b || !x) return GKOB200_EINVAL;
        V *r = vec(0), *r_tld = vec(1), *p = vec(2), *q = vec(3), *u = vec(4), *u_hat = vec(5), *v_hat = vec(6), *t = vec(7);
        V *alpha = sc(0), *beta = sc(1), *gamma = sc(2), *prev_rho = sc(3), *rho = sc(4);
        uint8_t* stat = this->stat();
        int rc;
        if ((rc = this->reset_state(s))) return rc;
        if ((rc = typed::cgs_initialize(tag, s, n, k, b, bs, r, r_tld, p, q, u, u_hat, v_hat, t, k, alpha, beta, gamma,
                                        prev_rho, rho, stat)))
            return rc;
        if ((rc = matrix_apply<V>(s, A, x, xs, k, this->neg_one(), this->one(), r, k, nullptr))) return rc;
        if ((rc = this->baseline_norm(s, b, bs, r, k))) return rc;
        if ((rc = typed::dense_copy(tag, s, n, k, r, k, r_tld, k))) return rc;
        launch_count += 4;
        int64_t it = 0;
        bool stopped = false;
        while (true) {
            if ((rc = typed::dense_compute_dot(tag, s, n, k, r, k, r_tld, k, rho, ws.p))) return rc;
            if ((rc = typed::dense_compute_norm2(tag, s, n, k, r, k, this->tau(), ws.p))) return rc;
            if ((rc = this->check(s, this->tau(), true, true))) return rc;
            launch_count += 2;
            if (it % this->chunk == 0 || it >= stop.max_iters) {
                if ((rc = this->poll(s, &stopped))) return rc;
                if (stopped) break;
            }
            if ((rc = typed::cgs_step_1(tag, s, n, k, r, u, p, q, k, beta, rho, prev_rho, stat))) return rc;
            if ((rc = this->precond_apply(s, p, k, t, k))) return rc;
            if ((rc = matrix_apply<V>(s, A, t, k, k, nullptr, nullptr, v_hat, k, nullptr))) return rc;
            if ((rc = typed::dense_compute_dot(tag, s, n, k, r_tld, k, v_hat, k, gamma, ws.p))) return rc;
            if ((rc = typed::cgs_step_2(tag, s, n, k, u, v_hat, q, t, k, alpha, rho, gamma, stat))) return rc;
            if ((rc = this->precond_apply(s, t, k, u_hat, k))) return rc;
            if ((rc = matrix_apply<V>(s, A, u_hat, k, k, nullptr, nullptr, t, k, nullptr))) return rc;
            if ((rc = typed::cgs_step_3(tag, s, n, k, t, u_hat, r, k, x, xs, alpha, stat))) return rc;
            if ((rc = typed::dense_copy(tag, s, int64_t(1), k, rho, k, prev_rho, k))) return rc;
            launch_count += 9;
            ++it;
        }
        return this->finish(s);
    }
};

template <typename S>
gkob200_solver* make(const gkob200_matrix* A, const gkob200_precond* M, const gkob200_stop* stop, int64_t nrhs,
                     int64_t krylov_dim, int* rc)
{
    auto* s = new S();
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
    if (krylov_dim > 0) {
        if (auto* g = dynamic_cast<GmresSolver<double>*>(static_cast<gkob200_solver*>(s))) g->m = krylov_dim;
        if (auto* g = dynamic_cast<GmresSolver<float>*>(static_cast<gkob200_solver*>(s))) g->m = krylov_dim;
    }
    *rc = s->init();
    if (*rc) {
        delete s;
        return nullptr;
    }
    return s;
}

}  // namespace

gkob200_solver* make_bicgstab_f64(const gkob200_matrix* A, const gkob200_precond* M, const gkob200_stop* st, int64_t nrhs, int* rc)
{
    return make<BicgstabSolver<double>>(A, M, st, nrhs, 0, rc);
}
gkob200_solver* make_bicgstab_f32(const gkob200_matrix* A, const gkob200_precond* M, const gkob200_stop* st, int64_t nrhs, int* rc)
{
    return make<BicgstabSolver<float>>(A, M, st, nrhs, 0, rc);
}
gkob200_solver* make_fcg(const gkob200_matrix* A, const gkob200_precond* M, const gkob200_stop* st, int64_t nrhs, int* rc)
{
    return A->value_type == GKOB200_F64 ? make<FcgSolver<double>>(A, M, st, nrhs, 0, rc)
                                        : make<FcgSolver<float>>(A, M, st, nrhs, 0, rc);
}
gkob200_solver* make_cgs(const gkob200_matrix* A, const gkob200_precond* M, const gkob200_stop* st, int64_t nrhs, int* rc)
{
    return A->value_type == GKOB200_F64 ? make<CgsSolver<double>>(A, M, st, nrhs, 0, rc)
                                        : make<CgsSolver<float>>(A, M, st, nrhs, 0, rc);
}
gkob200_solver* make_gmres_f64(const gkob200_matrix* A, const gkob200_precond* M, const gkob200_stop* st, int64_t nrhs, int64_t m, int* rc)
{
    return make<GmresSolver<double>>(A, M, st, nrhs, m, rc);
}
gkob200_solver* make_gmres_f32(const gkob200_matrix* A, const gkob200_precond* M, const gkob200_stop* st, int64_t nrhs, int64_t m, int* rc)
{
    return make<GmresSolver<float>>(A, M, st, nrhs, m, rc);
}

}  // namespace gkob200
