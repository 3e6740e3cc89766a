// cg_kernels.cu — the CG step kernels, the stopping-criterion kernels and the
// scalar-Jacobi kernels, 1:1 with the reference kernel set so that a Ginkgo shim
// can bind each of them.  The fused, graph-captured CG driver
// that the benchmarks use lives in solver_cg.cu.
//
// [ref] common/unified/solver/cg_kernels.cpp:51-138 (replaced),
//       reference/solver/cg_kernels.cpp:56-133 (oracle),
//       cuda/stop/residual_norm_kernels.cu:61-199, cuda/stop/criterion_kernels.cu:56-83,
//       common/unified/preconditioner/jacobi_kernels.cpp (scalar Jacobi),
//       reference/matrix/csr_kernels.cpp:1016-1034 (extract_diagonal).
#include "launch.cuh"

namespace gkob200 {
namespace {

template <typename V>
int cg_initialize(void* st, int64_t n, int64_t k, const V* b, int64_t bs, V* r, V* z, V* p, V* q, int64_t s,
                  V* prev_rho, V* rho, uint8_t* stop)
{
    if (n < 0 || k < 0) return GKOB200_EINVAL;
    if (k == 0) return 0;
    if (!prev_rho || !rho || !stop || (n > 0 && (!b || !r || !z || !p || !q))) return GKOB200_EINVAL;
    int rc = launch_2d(as_stream(st), 1, k, [=] __device__(int64_t, int64_t j) {
        rho[j] = V(0);
        prev_rho[j] = V(1);
        stop[j] = 0;
    });
    if (rc) return rc;
    return launch_2d(as_stream(st), n, k, [=] __device__(int64_t i, int64_t j) {
        r[i * s + j] = b[i * bs + j];
        z[i * s + j] = p[i * s + j] = q[i * s + j] = V(0);
    });
}

template <typename V>
int cg_step_1(void* st, int64_t n, int64_t k, V* p, const V* z, int64_t s, const V* rho, const V* prev_rho,
              const uint8_t* stop)
{
    if (n < 0 || k < 0) return GKOB200_EINVAL;
    if (n * k == 0) return 0;
    if (!p || !z || !rho || !prev_rho || !stop) return GKOB200_EINVAL;
    return launch_2d(as_stream(st), n, k, [=] __device__(int64_t i, int64_t j) {
        if (status_has_stopped(stop[j])) return;
        const V pr = prev_rho[j];
        // safe_divide(rho, prev_rho): 0 when prev_rho == 0 (math.hpp:1242-1245)
        const V t = pr == V(0) ? V(0) : div_rn(rho[j], pr);
        p[i * s + j] = add_rn(z[i * s + j], mul_rn(t, p[i * s + j]));
    });
}

template <typename V>
int cg_step_2(void* st, int64_t n, int64_t k, V* x, int64_t xs, V* r, const V* p, const V* q, int64_t s,
              const V* beta, const V* rho, const uint8_t* stop)
{
    if (n < 0 || k < 0) return GKOB200_EINVAL;
    if (n * k == 0) return 0;
    if (!x || !r || !p || !q || !beta || !rho || !stop) return GKOB200_EINVAL;
    return launch_2d(as_stream(st), n, k, [=] __device__(int64_t i, int64_t j) {
        if (status_has_stopped(stop[j])) return;
        const V be = beta[j];
        if (be == V(0)) return;
        const V t = div_rn(rho[j], be);
        x[i * xs + j] = add_rn(x[i * xs + j], mul_rn(t, p[i * s + j]));
        r[i * s + j] = sub_rn(r[i * s + j], mul_rn(t, q[i * s + j]));
    });
}

// One block; k is the number of right-hand sides (small).
template <typename V, bool Implicit>
__global__ void residual_norm_kernel(int64_t k, const V* __restrict__ tau, const V* __restrict__ orig_tau,
                                     V goal, uint8_t id, bool set_finalized, uint8_t* __restrict__ stop,
                                     uint8_t* __restrict__ flags)
{
    __shared__ int s_all, s_changed;
    if (threadIdx.x == 0) {
        s_all = 1;
        s_changed = 0;
    }
    __syncthreads();
    for (int64_t j = threadIdx.x; j < k; j += blockDim.x) {
        V t = tau[j];
        if (Implicit) t = sqrt_rn(t < V(0) ? -t : t);
        uint8_t s = stop[j];
        if (t < mul_rn(goal, orig_tau[j])) {
            // stopping_status::converge(id, set_finalized): only if not stopped yet
            if (!status_has_stopped(s)) {
                s |= 0x80 | (id & 0x3f);
                if (set_finalized) s |= 0x40;
                stop[j] = s;
            }
            s_changed = 1;  // reference sets one_changed whenever the test passes
        }
        if (!status_has_stopped(s)) s_all = 0;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        flags[0] = static_cast<uint8_t>(s_all);
        flags[1] = static_cast<uint8_t>(s_changed);
    }
}

template <typename V, bool Implicit>
int residual_norm_impl(void* st, int64_t k, const V* tau, const V* orig_tau, V goal, uint8_t id, int fin,
                       uint8_t* stop, uint8_t* flags)
{
    if (k < 0 || !flags || (k > 0 && (!tau || !orig_tau || !stop))) return GKOB200_EINVAL;
    residual_norm_kernel<V, Implicit><<<1, 256, 0, as_stream(st)>>>(k, tau, orig_tau, goal, id, fin != 0, stop, flags);
    GKOB200_CHECK_LAUNCH();
    return 0;
}

template <typename V, typename I>
__global__ void __launch_bounds__(256)
    extract_diagonal_kernel(int64_t n, const I* __restrict__ row_ptrs, const I* __restrict__ cols,
                            const V* __restrict__ vals, V* __restrict__ diag)
{
    // 8 lanes per row: the diagonal of a sorted row sits near its middle
    const int64_t gid = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x;
    const int64_t row = gid >> 3;
    const int sub = static_cast<int>(gid & 7);
    if (row >= n) return;
    const I b = row_ptrs[row], e = row_ptrs[row + 1];
    I found = e;  // first index with col == row
    for (I k = b + sub; k < e; k += 8) {
        if (static_cast<int64_t>(cols[k]) == row) {
            found = k;
            break;
        }
    }
    const unsigned mask = 0xffu << ((threadIdx.x & 31) & ~7);
    for (int o = 4; o > 0; o >>= 1) {
        const I other = __shfl_down_sync(mask, found, o, 8);
        found = other < found ? other : found;
    }
    if (sub == 0) diag[row] = found < e ? vals[found] : V(0);
}

template <typename V, typename I>
int extract_diagonal_impl(void* st, int64_t n_rows, int64_t n_cols, const I* row_ptrs, const I* cols,
                          const V* vals, V* diag)
{
    const int64_t n = n_rows < n_cols ? n_rows : n_cols;
    if (n < 0) return GKOB200_EINVAL;
    if (n == 0) return 0;
    if (!row_ptrs || !diag) return GKOB200_EINVAL;
    extract_diagonal_kernel<V, I><<<static_cast<unsigned>(ceildiv(n * 8, 256)), 256, 0, as_stream(st)>>>(
        n, row_ptrs, cols, vals, diag);
    GKOB200_CHECK_LAUNCH();
    return 0;
}

template <typename V>
int invert_diagonal_impl(void* st, int64_t n, const V* d, V* inv)
{
    if (n < 0 || (n > 0 && (!d || !inv))) return GKOB200_EINVAL;
    return launch_2d(as_stream(st), n, 1, [=] __device__(int64_t i, int64_t) {
        const V v = d[i];
        inv[i] = div_rn(V(1), v == V(0) ? V(1) : v);
    });
}

template <typename V>
int simple_scalar_apply_impl(void* st, int64_t n, int64_t k, const V* inv, const V* b, int64_t bs, V* x,
                             int64_t xs)
{
    if (n < 0 || k < 0 || (n * k > 0 && (!inv || !b || !x))) return GKOB200_EINVAL;
    return launch_2d(as_stream(st), n, k,
                     [=] __device__(int64_t i, int64_t j) { x[i * xs + j] = mul_rn(b[i * bs + j], inv[i]); });
}

template <typename V>
int scalar_apply_impl(void* st, int64_t n, int64_t k, const V* inv, const V* alpha, const V* b, int64_t bs,
                      const V* beta, V* x, int64_t xs)
{
    if (n < 0 || k < 0 || (n * k > 0 && (!inv || !b || !x || !alpha || !beta))) return GKOB200_EINVAL;
    return launch_2d(as_stream(st), n, k, [=] __device__(int64_t i, int64_t j) {
        x[i * xs + j] = add_rn(mul_rn(beta[0], x[i * xs + j]), mul_rn(mul_rn(alpha[0], b[i * bs + j]), inv[i]));
    });
}

}  // namespace
}  // namespace gkob200

using namespace gkob200;

extern "C" {

#define GKOB200_DEF_CG(V, VT)                                                                              \
    int gkob200_cg_initialize_##V(void* st, int64_t n, int64_t k, const VT* b, int64_t bs, VT* r, VT* z,   \
                                  VT* p, VT* q, int64_t s, VT* prev_rho, VT* rho, uint8_t* stop)           \
    { return cg_initialize<VT>(st, n, k, b, bs, r, z, p, q, s, prev_rho, rho, stop); }                     \
    int gkob200_cg_step_1_##V(void* st, int64_t n, int64_t k, VT* p, const VT* z, int64_t s, const VT* rho, \
                              const VT* prev_rho, const uint8_t* stop)                                     \
    { return cg_step_1<VT>(st, n, k, p, z, s, rho, prev_rho, stop); }                                      \
    int gkob200_cg_step_2_##V(void* st, int64_t n, int64_t k, VT* x, int64_t xs, VT* r, const VT* p,       \
                              const VT* q, int64_t s, const VT* beta, const VT* rho, const uint8_t* stop)  \
    { return cg_step_2<VT>(st, n, k, x, xs, r, p, q, s, beta, rho, stop); }                                \
    int gkob200_residual_norm_##V(void* st, int64_t k, const VT* tau, const VT* orig, VT goal, uint8_t id, \
                                  int fin, uint8_t* stop, uint8_t* flags)                                  \
    { return residual_norm_impl<VT, false>(st, k, tau, orig, goal, id, fin, stop, flags); }                \
    int gkob200_implicit_residual_norm_##V(void* st, int64_t k, const VT* tau, const VT* orig, VT goal,    \
                                           uint8_t id, int fin, uint8_t* stop, uint8_t* flags)             \
    { return residual_norm_impl<VT, true>(st, k, tau, orig, goal, id, fin, stop, flags); }                 \
    int gkob200_csr_extract_diagonal_##V##_i32(void* st, int64_t nr, int64_t nc, const int32_t* rp,        \
                                               const int32_t* ci, const VT* v, VT* d)                      \
    { return extract_diagonal_impl<VT, int32_t>(st, nr, nc, rp, ci, v, d); }                               \
    int gkob200_jacobi_invert_diagonal_##V(void* st, int64_t n, const VT* d, VT* inv)                      \
    { return invert_diagonal_impl<VT>(st, n, d, inv); }                                                    \
    int gkob200_jacobi_simple_scalar_apply_##V(void* st, int64_t n, int64_t k, const VT* inv, const VT* b, \
                                               int64_t bs, VT* x, int64_t xs)                              \
    { return simple_scalar_apply_impl<VT>(st, n, k, inv, b, bs, x, xs); }                                  \
    int gkob200_jacobi_scalar_apply_##V(void* st, int64_t n, int64_t k, const VT* inv, const VT* alpha,    \
                                        const VT* b, int64_t bs, const VT* beta, VT* x, int64_t xs)        \
    { return scalar_apply_impl<VT>(st, n, k, inv, alpha, b, bs, beta, x, xs); }
GKOB200_DEF_CG(f64, double)
GKOB200_DEF_CG(f32, float)

int gkob200_set_all_statuses(void* st, int64_t k, uint8_t id, int fin, uint8_t* stop)
{
    if (k < 0 || (k > 0 && !stop)) return GKOB200_EINVAL;
    const bool f = fin != 0;
    return launch_2d(as_stream(st), 1, k, [=] __device__(int64_t, int64_t j) {
        uint8_t s = stop[j];
        if (!status_has_stopped(s)) {  // stopping_status::stop
            s |= (id & 0x3f);
            if (f) s |= 0x40;
            stop[j] = s;
        }
    });
}

}  // extern "C"
