// dense.cu — Dense BLAS-1 kernels on n x k row-major vectors.
// Replaces the cuda instantiations of common/unified/matrix/dense_kernels.cpp:58-466
// and the cuBLAS dot/nrm2 calls of cuda/matrix/dense_kernels.cu:76-149; arithmetic
// follows the oracle reference/matrix/dense_kernels.cpp:158-378.
//
// Reductions are single-launch (ticketed grid reduction) and bit-reproducible.
// Algorithmic bytes (reference models benchmark/blas/blas.cpp:97-140): copy 2n,
// axpy 3n, scal 2n, dot 2n, norm n values.
#include <cstring>

#include "launch.cuh"

namespace gkob200 {
namespace {

template <typename V>
int fill_impl(void* st, int64_t n, int64_t k, V* x, int64_t xs, V value)
{
    if (n < 0 || k < 0 || (n * k > 0 && !x)) return GKOB200_EINVAL;
    return launch_2d(as_stream(st), n, k, [=] __device__(int64_t i, int64_t j) { x[i * xs + j] = value; });
}

template <typename V>
int copy_impl(void* st, int64_t n, int64_t k, const V* x, int64_t xs, V* y, int64_t ys)
{
    if (n < 0 || k < 0 || (n * k > 0 && (!x || !y))) return GKOB200_EINVAL;
    return launch_2d(as_stream(st), n, k,
                     [=] __device__(int64_t i, int64_t j) { y[i * ys + j] = x[i * xs + j]; });
}

template <typename V, int Op>  // 0 scale, 1 inv_scale
int scale_impl(void* st, int64_t n, int64_t k, const V* alpha, int64_t ac, V* x, int64_t xs)
{
    if (n < 0 || k < 0 || (n * k > 0 && (!x || !alpha)) || (ac != 1 && ac != k)) return GKOB200_EINVAL;
    const int64_t am = ac == 1 ? 0 : 1;
    return launch_2d(as_stream(st), n, k, [=] __device__(int64_t i, int64_t j) {
        const V a = alpha[j * am];
        V& v = x[i * xs + j];
        v = Op == 0 ? mul_rn(v, a) : div_rn(v, a);
    });
}

template <typename V, int Sign>  // +1 add_scaled, -1 sub_scaled
int axpy_impl(void* st, int64_t n, int64_t k, const V* alpha, int64_t ac, const V* x, int64_t xs, V* y,
              int64_t ys)
{
    if (n < 0 || k < 0 || (n * k > 0 && (!x || !y || !alpha)) || (ac != 1 && ac != k)) return GKOB200_EINVAL;
    const int64_t am = ac == 1 ? 0 : 1;
    return launch_2d(as_stream(st), n, k, [=] __device__(int64_t i, int64_t j) {
        const V t = mul_rn(alpha[j * am], x[i * xs + j]);
        V& v = y[i * ys + j];
        v = Sign > 0 ? add_rn(v, t) : sub_rn(v, t);
    });
}

template <typename V>
int dot_impl(void* st, int64_t n, int64_t k, const V* x, int64_t xs, const V* y, int64_t ys, V* result,
             void* ws)
{
    if (n < 0 || k < 0 || (k > 0 && !result) || (n * k > 0 && (!x || !y))) return GKOB200_EINVAL;
    return launch_col_reduce<V>(
        as_stream(st), n, k, ws, [=] __device__(int64_t i, int64_t j) { return x[i * xs + j] * y[i * ys + j]; },
        [=] __device__(int64_t j, V s) { result[j] = s; });
}

template <typename V, int Kind>  // 0 norm2, 1 squared norm2, 2 norm1
int norm_impl(void* st, int64_t n, int64_t k, const V* x, int64_t xs, V* result, void* ws)
{
    if (n < 0 || k < 0 || (k > 0 && !result) || (n * k > 0 && !x)) return GKOB200_EINVAL;
    return launch_col_reduce<V>(
        as_stream(st), n, k, ws,
        [=] __device__(int64_t i, int64_t j) {
            const V v = x[i * xs + j];
            return Kind == 2 ? (v < V(0) ? -v : v) : v * v;
        },
        [=] __device__(int64_t j, V s) { result[j] = Kind == 0 ? sqrt_rn(s) : s; });
}

template <typename V>
int sqrt_impl(void* st, int64_t k, V* x)
{
    if (k < 0 || (k > 0 && !x)) return GKOB200_EINVAL;
    return launch_2d(as_stream(st), 1, k, [=] __device__(int64_t, int64_t j) { x[j] = sqrt_rn(x[j]); });
}

template <typename V, typename I>
int row_gather_impl(void* st, int64_t n_out, int64_t k, const I* rows, const V* src, int64_t ss, V* dst,
                    int64_t ds)
{
    if (n_out < 0 || k < 0 || (n_out * k > 0 && (!rows || !src || !dst))) return GKOB200_EINVAL;
    return launch_2d(as_stream(st), n_out, k,
                     [=] __device__(int64_t i, int64_t j) { dst[i * ds + j] = src[rows[i] * ss + j]; });
}

}  // namespace

namespace {
int g_sm_count = 0;
int g_smem_optin = 0;
}
int sm_count()
{
    if (g_sm_count == 0) {
        int dev = 0, n = 0;
        if (cudaGetDevice(&dev) != cudaSuccess ||
            cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0)
            n = 148;
        g_sm_count = n;
    }
    return g_sm_count;
}
int max_smem_optin()
{
    if (g_smem_optin == 0) {
        int dev = 0, n = 0;
        if (cudaGetDevice(&dev) != cudaSuccess ||
            cudaDeviceGetAttribute(&n, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev) != cudaSuccess || n <= 0)
            n = 227 * 1024;
        g_smem_optin = n;
    }
    return g_smem_optin;
}
}  // namespace gkob200

using namespace gkob200;

extern "C" {

/* components::fill_array for any trivially copyable element of 1/2/4/8 bytes
 * [ref: core/components/fill_array_kernels.hpp; common/unified/components/fill_array_kernels.cpp] */
int gkob200_fill_array(void* stream, void* data, int64_t n, int elem_bytes, const void* value_host)
{
    if (n < 0 || !value_host || (n > 0 && !data)) return GKOB200_EINVAL;
    uint64_t v = 0;
    memcpy(&v, value_host, static_cast<size_t>(elem_bytes));
    switch (elem_bytes) {
    case 1: { auto* p = static_cast<uint8_t*>(data); const uint8_t x = static_cast<uint8_t>(v);
              return launch_2d(as_stream(stream), n, 1, [=] __device__(int64_t i, int64_t) { p[i] = x; }); }
    case 2: { auto* p = static_cast<uint16_t*>(data); const uint16_t x = static_cast<uint16_t>(v);
              return launch_2d(as_stream(stream), n, 1, [=] __device__(int64_t i, int64_t) { p[i] = x; }); }
    case 4: { auto* p = static_cast<uint32_t*>(data); const uint32_t x = static_cast<uint32_t>(v);
              return launch_2d(as_stream(stream), n, 1, [=] __device__(int64_t i, int64_t) { p[i] = x; }); }
    case 8: { auto* p = static_cast<uint64_t*>(data); const uint64_t x = v;
              return launch_2d(as_stream(stream), n, 1, [=] __device__(int64_t i, int64_t) { p[i] = x; }); }
    default: return GKOB200_EUNSUPPORTED;
    }
}

int gkob200_version(void) { return 100; }
int gkob200_sm_count(void) { return sm_count(); }
int gkob200_reduce_ws_init(void* stream, void* ws)
{
    if (!ws) return GKOB200_EINVAL;
    GKOB200_CUDA(cudaMemsetAsync(ws, 0, GKOB200_REDUCE_WS_BYTES, as_stream(stream)));
    return 0;
}

#define GKOB200_DEF_DENSE(V, VT)                                                                          \
    int gkob200_dense_fill_##V(void* s, int64_t n, int64_t k, VT* x, int64_t xs, VT v)                     \
    { return fill_impl<VT>(s, n, k, x, xs, v); }                                                          \
    int gkob200_dense_copy_##V(void* s, int64_t n, int64_t k, const VT* x, int64_t xs, VT* y, int64_t ys)  \
    { return copy_impl<VT>(s, n, k, x, xs, y, ys); }                                                      \
    int gkob200_dense_scale_##V(void* s, int64_t n, int64_t k, const VT* a, int64_t ac, VT* x, int64_t xs) \
    { return scale_impl<VT, 0>(s, n, k, a, ac, x, xs); }                                                  \
    int gkob200_dense_inv_scale_##V(void* s, int64_t n, int64_t k, const VT* a, int64_t ac, VT* x,         \
                                    int64_t xs)                                                           \
    { return scale_impl<VT, 1>(s, n, k, a, ac, x, xs); }                                                  \
    int gkob200_dense_add_scaled_##V(void* s, int64_t n, int64_t k, const VT* a, int64_t ac, const VT* x,  \
                                     int64_t xs, VT* y, int64_t ys)                                       \
    { return axpy_impl<VT, 1>(s, n, k, a, ac, x, xs, y, ys); }                                            \
    int gkob200_dense_sub_scaled_##V(void* s, int64_t n, int64_t k, const VT* a, int64_t ac, const VT* x,  \
                                     int64_t xs, VT* y, int64_t ys)                                       \
    { return axpy_impl<VT, -1>(s, n, k, a, ac, x, xs, y, ys); }                                           \
    int gkob200_dense_compute_dot_##V(void* s, int64_t n, int64_t k, const VT* x, int64_t xs, const VT* y, \
                                      int64_t ys, VT* r, void* ws)                                        \
    { return dot_impl<VT>(s, n, k, x, xs, y, ys, r, ws); }                                                \
    int gkob200_dense_compute_norm2_##V(void* s, int64_t n, int64_t k, const VT* x, int64_t xs, VT* r,     \
                                        void* ws)                                                         \
    { return norm_impl<VT, 0>(s, n, k, x, xs, r, ws); }                                                   \
    int gkob200_dense_compute_squared_norm2_##V(void* s, int64_t n, int64_t k, const VT* x, int64_t xs,    \
                                                VT* r, void* ws)                                          \
    { return norm_impl<VT, 1>(s, n, k, x, xs, r, ws); }                                                   \
    int gkob200_dense_compute_norm1_##V(void* s, int64_t n, int64_t k, const VT* x, int64_t xs, VT* r,     \
                                        void* ws)                                                         \
    { return norm_impl<VT, 2>(s, n, k, x, xs, r, ws); }                                                   \
    int gkob200_dense_compute_sqrt_##V(void* s, int64_t k, VT* x) { return sqrt_impl<VT>(s, k, x); }      \
    int gkob200_dense_row_gather_##V##_i32(void* s, int64_t n, int64_t k, const int32_t* rows,             \
                                           const VT* src, int64_t ss, VT* dst, int64_t ds)                \
    { return row_gather_impl<VT, int32_t>(s, n, k, rows, src, ss, dst, ds); }                             \
    int gkob200_dense_row_gather_##V##_i64(void* s, int64_t n, int64_t k, const int64_t* rows,             \
                                           const VT* src, int64_t ss, VT* dst, int64_t ds)                \
    { return row_gather_impl<VT, int64_t>(s, n, k, rows, src, ss, dst, ds); }
GKOB200_DEF_DENSE(f64, double)
GKOB200_DEF_DENSE(f32, float)

}  // extern "C"
