// solver_krylov.cu — BiCGSTAB and GMRES solver objects.
//
// The reference loops (core/solver/bicgstab.cpp:109-241, core/solver/gmres.cpp:139-369)
// issued kernel by kernel on the GPU with the stopping criterion evaluated ON THE DEVICE
// (solver_common.cuh): nothing is copied to the host per iteration (the reference does two
// blocking 1-byte copies per criterion check, i.e. four per BiCGSTAB iteration); the host
// polls the device-side status every `check_every` iterations and — for GMRES — at every
// restart boundary.  Kernels issued after a column has stopped are no-ops through the
// reference's own has_stopped() guards, so the reported iteration count is exact.
// Any number of right-hand sides.
#include "solver_common.cuh"

namespace gkob200 {
namespace {

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
        return scal.alloc(static_cast<size_t>(k) * 6 * sizeof(V));
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
        return fin.alloc(static_cast<size_t>(k) * sizeof(uint64_t));
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
            V* hnorm = hess_iter + (restart_iter + 1) * hs;
            if ((rc = typed::dense_compute_norm2(tag, s, n, k, next_k, k, hnorm, ws.p))) return rc;
            if ((rc = typed::dense_inv_scale(tag, s, n, k, hnorm, k, next_k, k))) return rc;
            if ((rc = typed::gmres_hessenberg_qr(tag, s, k, gsin, gcos, rnorm, rnc, hess_iter, hs, restart_iter, fin_it,
                                                 stat)))
                return rc;
            launch_count += 3;
            ++restart_iter;
        }
        if ((rc = update_x())) return rc;
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
gkob200_solver* make_gmres_f64(const gkob200_matrix* A, const gkob200_precond* M, const gkob200_stop* st, int64_t nrhs, int64_t m, int* rc)
{
    return make<GmresSolver<double>>(A, M, st, nrhs, m, rc);
}
gkob200_solver* make_gmres_f32(const gkob200_matrix* A, const gkob200_precond* M, const gkob200_stop* st, int64_t nrhs, int64_t m, int* rc)
{
    return make<GmresSolver<float>>(A, M, st, nrhs, m, rc);
}

}  // namespace gkob200
