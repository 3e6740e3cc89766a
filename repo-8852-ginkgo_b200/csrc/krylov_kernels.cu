// krylov_kernels.cu — BiCGSTAB and GMRES step kernels, 1:1 with the reference set.
//
// [ref] BiCGSTAB: core/solver/bicgstab_kernels.hpp:55-104; oracle
//       reference/solver/bicgstab_kernels.cpp:53-213; replaced
//       common/unified/solver/bicgstab_kernels.cpp:53-213.
//       GMRES: core/solver/{gmres,common_gmres}_kernels.hpp; oracle
//       reference/solver/{gmres,common_gmres}_kernels.cpp; replaced
//       common/unified/solver/{gmres,common_gmres}_kernels.cpp (where hessenberg_qr and
//       solve_krylov run one thread per right-hand side, as they do here — they are
//       O(m) scalar recurrences).
// All vectors n x k row-major with a common `stride`; scalars 1 x k.
#include "launch.cuh"

namespace gkob200 {
namespace {

template <typename V>
__device__ __forceinline__ V vabs(V v) { return v < V(0) ? -v : v; }

// ------------------------------- BiCGSTAB ------------------------------------
template <typename V>
int bicgstab_initialize(void* st, int64_t n, int64_t k, const V* b, int64_t bs, V* r, V* rr, V* y, V* s_, V* t, V* z,
                        V* v, V* p, int64_t s, V* prev_rho, V* rho, V* alpha, V* beta, V* gamma, V* omega,
                        uint8_t* stop)
{
    if (n < 0 || k < 0) return GKOB200_EINVAL;
    if (k == 0) return 0;
    int rc = launch_2d(as_stream(st), 1, k, [=] __device__(int64_t, int64_t j) {
        rho[j] = prev_rho[j] = alpha[j] = beta[j] = gamma[j] = omega[j] = V(1);
        stop[j] = 0;
    });
    if (rc) return rc;
    return launch_2d(as_stream(st), n, k, [=] __device__(int64_t i, int64_t j) {
        r[i * s + j] = b[i * bs + j];
        rr[i * s + j] = z[i * s + j] = v[i * s + j] = s_[i * s + j] = t[i * s + j] = y[i * s + j] = p[i * s + j] = V(0);
    });
}

template <typename V>
int bicgstab_step_1(void* st, int64_t n, int64_t k, const V* r, V* p, const V* v, int64_t s, const V* rho,
                    const V* prev_rho, const V* alpha, const V* omega, const uint8_t* stop)
{
    if (n < 0 || k < 0) return GKOB200_EINVAL;
    return launch_2d(as_stream(st), n, k, [=] __device__(int64_t i, int64_t j) {
        if (status_has_stopped(stop[j])) return;
        const V om = omega[j];
        if (mul_rn(prev_rho[j], om) != V(0)) {
            // rho / prev_rho * alpha / omega, left to right
            const V tmp = div_rn(mul_rn(div_rn(rho[j], prev_rho[j]), alpha[j]), om);
            p[i * s + j] = add_rn(r[i * s + j], mul_rn(tmp, sub_rn(p[i * s + j], mul_rn(om, v[i * s + j]))));
        } else {
            p[i * s + j] = r[i * s + j];
        }
    });
}

template <typename V>
int bicgstab_step_2(void* st, int64_t n, int64_t k, const V* r, V* s_, const V* v, int64_t s, const V* rho, V* alpha,
                    const V* beta, const uint8_t* stop)
{
    if (n < 0 || k < 0) return GKOB200_EINVAL;
    return launch_2d(as_stream(st), n, k, [=] __device__(int64_t i, int64_t j) {
        if (status_has_stopped(stop[j])) return;
        const V be = beta[j];
        if (be != V(0)) {
            const V a = div_rn(rho[j], be);
            if (i == 0) alpha[j] = a;
            s_[i * s + j] = sub_rn(r[i * s + j], mul_rn(a, v[i * s + j]));
        } else {
            if (i == 0) alpha[j] = V(0);
            s_[i * s + j] = r[i * s + j];
        }
    });
}

template <typename V>
int bicgstab_step_3(void* st, int64_t n, int64_t k, V* x, int64_t xs, V* r, const V* s_, const V* t, const V* y,
                    const V* z, int64_t s, const V* alpha, const V* beta, const V* gamma, V* omega,
                    const uint8_t* stop)
{
    if (n < 0 || k < 0) return GKOB200_EINVAL;
    if (n == 0 && k > 0) {
        // the reference updates omega even for an empty system
        return launch_2d(as_stream(st), 1, k, [=] __device__(int64_t, int64_t j) {
            if (status_has_stopped(stop[j])) return;
            omega[j] = beta[j] != V(0) ? div_rn(gamma[j], beta[j]) : V(0);
        });
    }
    return launch_2d(as_stream(st), n, k, [=] __device__(int64_t i, int64_t j) {
        if (status_has_stopped(stop[j])) return;
        const V om = beta[j] != V(0) ? div_rn(gamma[j], beta[j]) : V(0);
        if (i == 0) omega[j] = om;
        x[i * xs + j] = add_rn(x[i * xs + j], add_rn(mul_rn(alpha[j], y[i * s + j]), mul_rn(om, z[i * s + j])));
        r[i * s + j] = sub_rn(s_[i * s + j], mul_rn(om, t[i * s + j]));
    });
}

template <typename V>
int bicgstab_finalize(void* st, int64_t n, int64_t k, V* x, int64_t xs, const V* y, int64_t s, const V* alpha,
                      uint8_t* stop)
{
    if (n < 0 || k < 0) return GKOB200_EINVAL;
    int rc = launch_2d(as_stream(st), n, k, [=] __device__(int64_t i, int64_t j) {
        const uint8_t sj = stop[j];
        if (status_has_stopped(sj) && !status_is_finalized(sj))
            x[i * xs + j] = add_rn(x[i * xs + j], mul_rn(alpha[j], y[i * s + j]));
    });
    if (rc) return rc;
    // second launch: status bytes are only rewritten once every reader above is done
    if (n == 0) return 0;  // reference: finalize() sits inside the row loop
    return launch_2d(as_stream(st), 1, k, [=] __device__(int64_t, int64_t j) {
        if (status_has_stopped(stop[j])) stop[j] |= 0x40;
    });
}

// --------------------------------- FCG / CGS ----------------------------------
// 1:1 with the reference kernels [fcg: common/unified/solver/fcg_kernels.cpp, oracle
// reference/solver/fcg_kernels.cpp:50-128; cgs: common/unified/solver/cgs_kernels.cpp, oracle
// reference/solver/cgs_kernels.cpp:50-167].
template <typename V>
int fcg_initialize(void* st, int64_t n, int64_t k, const V* b, int64_t bs, V* r, V* z, V* p, V* q, V* t, int64_t s,
                   V* prev_rho, V* rho, V* rho_t, uint8_t* stop)
{
    if (n < 0 || k < 0) return GKOB200_EINVAL;
    if (k == 0) return 0;
    int rc = launch_2d(as_stream(st), 1, k, [=] __device__(int64_t, int64_t j) {
        rho[j] = V(0);
        prev_rho[j] = rho_t[j] = V(1);
        stop[j] = 0;
    });
    if (rc) return rc;
    return launch_2d(as_stream(st), n, k, [=] __device__(int64_t i, int64_t j) {
        t[i * s + j] = r[i * s + j] = b[i * bs + j];
        z[i * s + j] = p[i * s + j] = q[i * s + j] = V(0);
    });
}

template <typename V>
int fcg_step_1(void* st, int64_t n, int64_t k, V* p, const V* z, int64_t s, const V* rho_t, const V* prev_rho,
               const uint8_t* stop)
{
    if (n < 0 || k < 0) return GKOB200_EINVAL;
    return launch_2d(as_stream(st), n, k, [=] __device__(int64_t i, int64_t j) {
        if (status_has_stopped(stop[j])) return;
        const V pr = prev_rho[j];
        p[i * s + j] = pr == V(0) ? z[i * s + j] : add_rn(z[i * s + j], mul_rn(div_rn(rho_t[j], pr), p[i * s + j]));
    });
}

template <typename V>
int fcg_step_2(void* st, int64_t n, int64_t k, V* x, int64_t xs, V* r, V* t, const V* p, const V* q, int64_t s,
               const V* beta, const V* rho, const uint8_t* stop)
{
    if (n < 0 || k < 0) return GKOB200_EINVAL;
    return launch_2d(as_stream(st), n, k, [=] __device__(int64_t i, int64_t j) {
        if (status_has_stopped(stop[j])) return;
        const V be = beta[j];
        if (be == V(0)) return;
        const V tmp = div_rn(rho[j], be);
        const V prev_r = r[i * s + j];
        x[i * xs + j] = add_rn(x[i * xs + j], mul_rn(tmp, p[i * s + j]));
        const V nr = sub_rn(prev_r, mul_rn(tmp, q[i * s + j]));
        r[i * s + j] = nr;
        t[i * s + j] = sub_rn(nr, prev_r);
    });
}

// BiCG step kernels [ref: common/unified/solver/bicg_kernels.cpp:53-170; the arithmetic of
// reference/solver/bicg_kernels.cpp].  The transposed system matrix and preconditioner are the
// caller's (core/solver/bicg.cpp:160-185 builds them with csr::transpose).
template <typename V>
int bicg_initialize(void* st, int64_t n, int64_t k, const V* b, int64_t bs, V* r, V* z, V* p, V* q, V* r2, V* z2, V* p2,
                    V* q2, int64_t s, V* prev_rho, V* rho, uint8_t* stop)
{
    if (n < 0 || k < 0) return GKOB200_EINVAL;
    if (k == 0) return 0;
    int rc = launch_2d(as_stream(st), 1, k, [=] __device__(int64_t, int64_t j) {
        rho[j] = V(0);
        prev_rho[j] = V(1);
        stop[j] = 0;
    });
    if (rc) return rc;
    return launch_2d(as_stream(st), n, k, [=] __device__(int64_t i, int64_t j) {
        r[i * s + j] = r2[i * s + j] = b[i * bs + j];
        z[i * s + j] = p[i * s + j] = q[i * s + j] = z2[i * s + j] = p2[i * s + j] = q2[i * s + j] = V(0);
    });
}

template <typename V>
int bicg_step_1(void* st, int64_t n, int64_t k, V* p, const V* z, V* p2, const V* z2, int64_t s, const V* rho,
                const V* prev_rho, const uint8_t* stop)
{
    if (n < 0 || k < 0) return GKOB200_EINVAL;
    return launch_2d(as_stream(st), n, k, [=] __device__(int64_t i, int64_t j) {
        if (status_has_stopped(stop[j])) return;
        const V pr = prev_rho[j];
        const V tmp = pr == V(0) ? V(0) : div_rn(rho[j], pr);   // safe_divide
        p[i * s + j] = add_rn(z[i * s + j], mul_rn(tmp, p[i * s + j]));
        p2[i * s + j] = add_rn(z2[i * s + j], mul_rn(tmp, p2[i * s + j]));
    });
}

template <typename V>
int bicg_step_2(void* st, int64_t n, int64_t k, V* x, int64_t xs, V* r, V* r2, const V* p, const V* q, const V* q2,
                int64_t s, const V* beta, const V* rho, const uint8_t* stop)
{
    if (n < 0 || k < 0) return GKOB200_EINVAL;
    return launch_2d(as_stream(st), n, k, [=] __device__(int64_t i, int64_t j) {
        if (status_has_stopped(stop[j])) return;
        const V be = beta[j];
        const V tmp = be == V(0) ? V(0) : div_rn(rho[j], be);   // safe_divide
        x[i * xs + j] = add_rn(x[i * xs + j], mul_rn(tmp, p[i * s + j]));
        r[i * s + j] = sub_rn(r[i * s + j], mul_rn(tmp, q[i * s + j]));
        r2[i * s + j] = sub_rn(r2[i * s + j], mul_rn(tmp, q2[i * s + j]));
    });
}

template <typename V>
int cgs_initialize(void* st, int64_t n, int64_t k, const V* b, int64_t bs, V* r, V* r_tld, V* p, V* q, V* u, V* u_hat,
                   V* v_hat, V* t, int64_t s, V* alpha, V* beta, V* gamma, V* rho_prev, V* rho, uint8_t* stop)
{
    if (n < 0 || k < 0) return GKOB200_EINVAL;
    if (k == 0) return 0;
    int rc = launch_2d(as_stream(st), 1, k, [=] __device__(int64_t, int64_t j) {
        rho[j] = V(0);
        rho_prev[j] = alpha[j] = beta[j] = gamma[j] = V(1);
        stop[j] = 0;
    });
    if (rc) return rc;
    return launch_2d(as_stream(st), n, k, [=] __device__(int64_t i, int64_t j) {
        r[i * s + j] = r_tld[i * s + j] = b[i * bs + j];
        u[i * s + j] = u_hat[i * s + j] = p[i * s + j] = q[i * s + j] = v_hat[i * s + j] = t[i * s + j] = V(0);
    });
}

// beta is only updated where rho_prev != 0 (it keeps its previous value otherwise): the scalar
// update is its own launch so that every vector entry sees the same beta
template <typename V>
int cgs_step_1(void* st, int64_t n, int64_t k, const V* r, V* u, V* p, const V* q, int64_t s, V* beta, const V* rho,
               const V* rho_prev, const uint8_t* stop)
{
    if (n < 0 || k < 0) return GKOB200_EINVAL;
    if (k == 0) return 0;
    int rc = launch_2d(as_stream(st), 1, k, [=] __device__(int64_t, int64_t j) {
        if (status_has_stopped(stop[j])) return;
        if (rho_prev[j] != V(0)) beta[j] = div_rn(rho[j], rho_prev[j]);
    });
    if (rc) return rc;
    return launch_2d(as_stream(st), n, k, [=] __device__(int64_t i, int64_t j) {
        if (status_has_stopped(stop[j])) return;
        const V be = beta[j];
        const V qi = q[i * s + j];
        const V ui = add_rn(r[i * s + j], mul_rn(be, qi));
        u[i * s + j] = ui;
        p[i * s + j] = add_rn(ui, mul_rn(be, add_rn(qi, mul_rn(be, p[i * s + j]))));
    });
}

template <typename V>
int cgs_step_2(void* st, int64_t n, int64_t k, const V* u, const V* v_hat, V* q, V* t, int64_t s, V* alpha, const V* rho,
               const V* gamma, const uint8_t* stop)
{
    if (n < 0 || k < 0) return GKOB200_EINVAL;
    if (k == 0) return 0;
    int rc = launch_2d(as_stream(st), 1, k, [=] __device__(int64_t, int64_t j) {
        if (status_has_stopped(stop[j])) return;
        if (gamma[j] != V(0)) alpha[j] = div_rn(rho[j], gamma[j]);
    });
    if (rc) return rc;
    return launch_2d(as_stream(st), n, k, [=] __device__(int64_t i, int64_t j) {
        if (status_has_stopped(stop[j])) return;
        const V ui = u[i * s + j];
        const V qi = sub_rn(ui, mul_rn(alpha[j], v_hat[i * s + j]));
        q[i * s + j] = qi;
        t[i * s + j] = add_rn(ui, qi);
    });
}

template <typename V>
int cgs_step_3(void* st, int64_t n, int64_t k, const V* t, const V* u_hat, V* r, int64_t s, V* x, int64_t xs,
               const V* alpha, const uint8_t* stop)
{
    if (n < 0 || k < 0) return GKOB200_EINVAL;
    return launch_2d(as_stream(st), n, k, [=] __device__(int64_t i, int64_t j) {
        if (status_has_stopped(stop[j])) return;
        const V a = alpha[j];
        x[i * xs + j] = add_rn(x[i * xs + j], mul_rn(a, u_hat[i * s + j]));
        r[i * s + j] = sub_rn(r[i * s + j], mul_rn(a, t[i * s + j]));
    });
}

// --------------------------------- GMRES --------------------------------------
template <typename V>
int gmres_initialize(void* st, int64_t n, int64_t k, int64_t krylov_dim, const V* b, int64_t bs, V* residual,
                     int64_t rs, V* givens_sin, V* givens_cos, uint8_t* stop)
{
    if (n < 0 || k < 0 || krylov_dim < 0) return GKOB200_EINVAL;
    int rc = launch_2d(as_stream(st), n, k, [=] __device__(int64_t i, int64_t j) { residual[i * rs + j] = b[i * bs + j]; });
    if (rc) return rc;
    rc = launch_2d(as_stream(st), krylov_dim, k, [=] __device__(int64_t i, int64_t j) {
        givens_sin[i * k + j] = V(0);
        givens_cos[i * k + j] = V(0);
    });
    if (rc) return rc;
    return launch_2d(as_stream(st), 1, k, [=] __device__(int64_t, int64_t j) { stop[j] = 0; });
}

template <typename V>
int gmres_restart(void* st, int64_t n, int64_t k, const V* residual, int64_t rs, const V* residual_norm,
                  V* residual_norm_collection, V* krylov_bases, uint64_t* final_iter_nums)
{
    if (n < 0 || k < 0) return GKOB200_EINVAL;
    int rc = launch_2d(as_stream(st), 1, k, [=] __device__(int64_t, int64_t j) {
        residual_norm_collection[j] = residual_norm[j];
        final_iter_nums[j] = 0;
    });
    if (rc) return rc;
    return launch_2d(as_stream(st), n, k, [=] __device__(int64_t i, int64_t j) {
        krylov_bases[i * k + j] = div_rn(residual[i * rs + j], residual_norm[j]);
    });
}

template <typename V>
__global__ void __launch_bounds__(256)
    multi_axpy_kernel(int64_t n, int64_t k, const V* __restrict__ bases, const V* __restrict__ y, V* __restrict__ out,
                      int64_t os, const uint64_t* __restrict__ final_iter_nums, const uint8_t* __restrict__ stop)
{
    const int64_t total = n * k;
    for (int64_t t = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; t < total;
         t += static_cast<int64_t>(gridDim.x) * blockDim.x) {
        const int64_t i = t / k, c = t % k;
        if (status_is_finalized(stop[c])) continue;
        V acc = V(0);
        const int64_t m = static_cast<int64_t>(final_iter_nums[c]);
        for (int64_t j = 0; j < m; ++j) acc = add_rn(acc, mul_rn(bases[(i + j * n) * k + c], y[j * k + c]));
        out[i * os + c] = acc;
    }
}

template <typename V>
int gmres_multi_axpy(void* st, int64_t n, int64_t k, const V* krylov_bases, const V* y, V* before_preconditioner,
                     int64_t bps, const uint64_t* final_iter_nums, uint8_t* stop)
{
    if (n < 0 || k < 0) return GKOB200_EINVAL;
    if (k == 0) return 0;
    if (n > 0) {
        multi_axpy_kernel<V><<<grid_for(n * k, 256, 8), 256, 0, as_stream(st)>>>(n, k, krylov_bases, y,
                                                                                before_preconditioner, bps,
                                                                                final_iter_nums, stop);
        GKOB200_CHECK_LAUNCH();
    }
    return launch_2d(as_stream(st), 1, k, [=] __device__(int64_t, int64_t j) {
        const uint8_t s = stop[j];
        if (!status_is_finalized(s) && status_has_stopped(s)) stop[j] = s | 0x40;
    });
}

// one thread per right-hand side [reference/solver/common_gmres_kernels.cpp:57-209]
template <typename V>
__global__ void hessenberg_qr_kernel(int64_t k, V* givens_sin, V* givens_cos, V* residual_norm,
                                     V* residual_norm_collection, V* hess, int64_t hs, int64_t iter,
                                     uint64_t* final_iter_nums, const uint8_t* stop, const int* skip)
{
    if (skip && *skip) return;
    const int64_t c = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x;
    if (c >= k) return;
    if (status_has_stopped(stop[c])) return;
    final_iter_nums[c] += 1;
    // givens_rotation
    for (int64_t j = 0; j < iter; ++j) {
        const V hj = hess[j * hs + c], hj1 = hess[(j + 1) * hs + c];
        const V co = givens_cos[j * k + c], si = givens_sin[j * k + c];
        const V temp = add_rn(mul_rn(co, hj), mul_rn(si, hj1));
        hess[(j + 1) * hs + c] = add_rn(mul_rn(-si, hj), mul_rn(co, hj1));
        hess[j * hs + c] = temp;
    }
    // calculate_sin_and_cos
    const V this_h = hess[iter * hs + c], next_h = hess[(iter + 1) * hs + c];
    V co, si;
    if (this_h == V(0)) {
        co = V(0);
        si = V(1);
    } else {
        const V scale = add_rn(vabs(this_h), vabs(next_h));
        const V a = vabs(div_rn(this_h, scale)), b = vabs(div_rn(next_h, scale));
        const V hyp = mul_rn(scale, sqrt_rn(add_rn(mul_rn(a, a), mul_rn(b, b))));
        co = div_rn(this_h, hyp);
        si = div_rn(next_h, hyp);
    }
    givens_cos[iter * k + c] = co;
    givens_sin[iter * k + c] = si;
    hess[iter * hs + c] = add_rn(mul_rn(co, this_h), mul_rn(si, next_h));
    hess[(iter + 1) * hs + c] = V(0);
    // calculate_next_residual_norm
    const V rnc = residual_norm_collection[iter * k + c];
    const V nxt = mul_rn(-si, rnc);
    residual_norm_collection[(iter + 1) * k + c] = nxt;
    residual_norm_collection[iter * k + c] = mul_rn(co, rnc);
    residual_norm[c] = vabs(nxt);
}

template <typename V>
int gmres_hessenberg_qr(void* st, int64_t k, V* givens_sin, V* givens_cos, V* residual_norm,
                        V* residual_norm_collection, V* hessenberg_iter, int64_t hess_stride, int64_t iter,
                        uint64_t* final_iter_nums, const uint8_t* stop)
{
    if (k < 0 || iter < 0) return GKOB200_EINVAL;
    if (k == 0) return 0;
    hessenberg_qr_kernel<V><<<static_cast<unsigned>(ceildiv(k, 128)), 128, 0, as_stream(st)>>>(
        k, givens_sin, givens_cos, residual_norm, residual_norm_collection, hessenberg_iter, hess_stride, iter,
        final_iter_nums, stop, nullptr);
    GKOB200_CHECK_LAUNCH();
    return 0;
}

template <typename V>
__global__ void solve_krylov_kernel(int64_t k, const V* rnc, const V* hess, int64_t hs, V* y,
                                    const uint64_t* final_iter_nums, const uint8_t* stop)
{
    const int64_t c = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x;
    if (c >= k) return;
    if (status_is_finalized(stop[c])) return;
    const int64_t m = static_cast<int64_t>(final_iter_nums[c]);
    for (int64_t i = m - 1; i >= 0; --i) {
        V temp = rnc[i * k + c];
        for (int64_t j = i + 1; j < m; ++j) temp = sub_rn(temp, mul_rn(hess[i * hs + j * k + c], y[j * k + c]));
        y[i * k + c] = div_rn(temp, hess[i * hs + i * k + c]);
    }
}

template <typename V>
int gmres_solve_krylov(void* st, int64_t k, const V* residual_norm_collection, const V* hessenberg,
                       int64_t hess_stride, V* y, const uint64_t* final_iter_nums, const uint8_t* stop)
{
    if (k < 0) return GKOB200_EINVAL;
    if (k == 0) return 0;
    solve_krylov_kernel<V><<<static_cast<unsigned>(ceildiv(k, 128)), 128, 0, as_stream(st)>>>(
        k, residual_norm_collection, hessenberg, hess_stride, y, final_iter_nums, stop);
    GKOB200_CHECK_LAUNCH();
    return 0;
}

}  // namespace
}  // namespace gkob200

using namespace gkob200;

extern "C" {

#define GKOB200_DEF_KRYLOV(V, VT)                                                                                 \
    int gkob200_fcg_initialize_##V(void* st, int64_t n, int64_t k, const VT* b, int64_t bs, VT* r, VT* z, VT* p,   \
                                   VT* q, VT* t, int64_t s, VT* prev_rho, VT* rho, VT* rho_t, uint8_t* stop)       \
    { return fcg_initialize<VT>(st, n, k, b, bs, r, z, p, q, t, s, prev_rho, rho, rho_t, stop); }                  \
    int gkob200_fcg_step_1_##V(void* st, int64_t n, int64_t k, VT* p, const VT* z, int64_t s, const VT* rho_t,     \
                               const VT* prev_rho, const uint8_t* stop)                                           \
    { return fcg_step_1<VT>(st, n, k, p, z, s, rho_t, prev_rho, stop); }                                           \
    int gkob200_fcg_step_2_##V(void* st, int64_t n, int64_t k, VT* x, int64_t xs, VT* r, VT* t, const VT* p,       \
                               const VT* q, int64_t s, const VT* beta, const VT* rho, const uint8_t* stop)         \
    { return fcg_step_2<VT>(st, n, k, x, xs, r, t, p, q, s, beta, rho, stop); }                                    \
    int gkob200_bicg_initialize_##V(void* st, int64_t n, int64_t k, const VT* b, int64_t bs, VT* r, VT* z, VT* p,  \
                                    VT* q, VT* r2, VT* z2, VT* p2, VT* q2, int64_t s, VT* prev_rho, VT* rho,       \
                                    uint8_t* stop)                                                                \
    { return bicg_initialize<VT>(st, n, k, b, bs, r, z, p, q, r2, z2, p2, q2, s, prev_rho, rho, stop); }           \
    int gkob200_bicg_step_1_##V(void* st, int64_t n, int64_t k, VT* p, const VT* z, VT* p2, const VT* z2,          \
                                int64_t s, const VT* rho, const VT* prev_rho, const uint8_t* stop)                 \
    { return bicg_step_1<VT>(st, n, k, p, z, p2, z2, s, rho, prev_rho, stop); }                                    \
    int gkob200_bicg_step_2_##V(void* st, int64_t n, int64_t k, VT* x, int64_t xs, VT* r, VT* r2, const VT* p,     \
                                const VT* q, const VT* q2, int64_t s, const VT* beta, const VT* rho,               \
                                const uint8_t* stop)                                                              \
    { return bicg_step_2<VT>(st, n, k, x, xs, r, r2, p, q, q2, s, beta, rho, stop); }                              \
    int gkob200_cgs_initialize_##V(void* st, int64_t n, int64_t k, const VT* b, int64_t bs, VT* r, VT* r_tld,      \
                                   VT* p, VT* q, VT* u, VT* u_hat, VT* v_hat, VT* t, int64_t s, VT* alpha,         \
                                   VT* beta, VT* gamma, VT* rho_prev, VT* rho, uint8_t* stop)                      \
    { return cgs_initialize<VT>(st, n, k, b, bs, r, r_tld, p, q, u, u_hat, v_hat, t, s, alpha, beta, gamma,        \
                                rho_prev, rho, stop); }                                                           \
    int gkob200_cgs_step_1_##V(void* st, int64_t n, int64_t k, const VT* r, VT* u, VT* p, const VT* q, int64_t s,  \
                               VT* beta, const VT* rho, const VT* rho_prev, const uint8_t* stop)                   \
    { return cgs_step_1<VT>(st, n, k, r, u, p, q, s, beta, rho, rho_prev, stop); }                                 \
    int gkob200_cgs_step_2_##V(void* st, int64_t n, int64_t k, const VT* u, const VT* v_hat, VT* q, VT* t,         \
                               int64_t s, VT* alpha, const VT* rho, const VT* gamma, const uint8_t* stop)          \
    { return cgs_step_2<VT>(st, n, k, u, v_hat, q, t, s, alpha, rho, gamma, stop); }                               \
    int gkob200_cgs_step_3_##V(void* st, int64_t n, int64_t k, const VT* t, const VT* u_hat, VT* r, int64_t s,     \
                               VT* x, int64_t xs, const VT* alpha, const uint8_t* stop)                           \
    { return cgs_step_3<VT>(st, n, k, t, u_hat, r, s, x, xs, alpha, stop); }                                       \
    int gkob200_bicgstab_initialize_##V(void* st, int64_t n, int64_t k, const VT* b, int64_t bs, VT* r, VT* rr,    \
                                        VT* y, VT* s_, VT* t, VT* z, VT* v, VT* p, int64_t s, VT* prev_rho,        \
                                        VT* rho, VT* alpha, VT* beta, VT* gamma, VT* omega, uint8_t* stop)         \
    { return bicgstab_initialize<VT>(st, n, k, b, bs, r, rr, y, s_, t, z, v, p, s, prev_rho, rho, alpha, beta,     \
                                     gamma, omega, stop); }                                                       \
    int gkob200_bicgstab_step_1_##V(void* st, int64_t n, int64_t k, const VT* r, VT* p, const VT* v, int64_t s,    \
                                    const VT* rho, const VT* prev_rho, const VT* alpha, const VT* omega,           \
                                    const uint8_t* stop)                                                          \
    { return bicgstab_step_1<VT>(st, n, k, r, p, v, s, rho, prev_rho, alpha, omega, stop); }                       \
    int gkob200_bicgstab_step_2_##V(void* st, int64_t n, int64_t k, const VT* r, VT* s_, const VT* v, int64_t s,   \
                                    const VT* rho, VT* alpha, const VT* beta, const uint8_t* stop)                 \
    { return bicgstab_step_2<VT>(st, n, k, r, s_, v, s, rho, alpha, beta, stop); }                                 \
    int gkob200_bicgstab_step_3_##V(void* st, int64_t n, int64_t k, VT* x, int64_t xs, VT* r, const VT* s_,        \
                                    const VT* t, const VT* y, const VT* z, int64_t s, const VT* alpha,             \
                                    const VT* beta, const VT* gamma, VT* omega, const uint8_t* stop)               \
    { return bicgstab_step_3<VT>(st, n, k, x, xs, r, s_, t, y, z, s, alpha, beta, gamma, omega, stop); }           \
    int gkob200_bicgstab_finalize_##V(void* st, int64_t n, int64_t k, VT* x, int64_t xs, const VT* y, int64_t s,   \
                                      const VT* alpha, uint8_t* stop)                                             \
    { return bicgstab_finalize<VT>(st, n, k, x, xs, y, s, alpha, stop); }                                          \
    int gkob200_gmres_initialize_##V(void* st, int64_t n, int64_t k, int64_t krylov_dim, const VT* b, int64_t bs,  \
                                     VT* residual, int64_t rs, VT* givens_sin, VT* givens_cos, uint8_t* stop)      \
    { return gmres_initialize<VT>(st, n, k, krylov_dim, b, bs, residual, rs, givens_sin, givens_cos, stop); }      \
    int gkob200_gmres_restart_##V(void* st, int64_t n, int64_t k, const VT* residual, int64_t rs,                  \
                                  const VT* residual_norm, VT* residual_norm_collection, VT* krylov_bases,         \
                                  uint64_t* final_iter_nums)                                                      \
    { return gmres_restart<VT>(st, n, k, residual, rs, residual_norm, residual_norm_collection, krylov_bases,      \
                               final_iter_nums); }                                                                \
    int gkob200_gmres_multi_axpy_##V(void* st, int64_t n, int64_t k, const VT* krylov_bases, const VT* y,          \
                                     VT* before_preconditioner, int64_t bps, const uint64_t* final_iter_nums,      \
                                     uint8_t* stop)                                                               \
    { return gmres_multi_axpy<VT>(st, n, k, krylov_bases, y, before_preconditioner, bps, final_iter_nums, stop); } \
    int gkob200_gmres_hessenberg_qr_##V(void* st, int64_t k, VT* givens_sin, VT* givens_cos, VT* residual_norm,    \
                                        VT* residual_norm_collection, VT* hessenberg_iter, int64_t hess_stride,    \
                                        int64_t iter, uint64_t* final_iter_nums, const uint8_t* stop)              \
    { return gmres_hessenberg_qr<VT>(st, k, givens_sin, givens_cos, residual_norm, residual_norm_collection,       \
                                     hessenberg_iter, hess_stride, iter, final_iter_nums, stop); }                 \
    int gkob200_gmres_solve_krylov_##V(void* st, int64_t k, const VT* residual_norm_collection,                    \
                                       const VT* hessenberg, int64_t hess_stride, VT* y,                           \
                                       const uint64_t* final_iter_nums, const uint8_t* stop)                       \
    { return gmres_solve_krylov<VT>(st, k, residual_norm_collection, hessenberg, hess_stride, y, final_iter_nums,  \
                                    stop); }
GKOB200_DEF_KRYLOV(f64, double)
GKOB200_DEF_KRYLOV(f32, float)

}  // extern "C"
