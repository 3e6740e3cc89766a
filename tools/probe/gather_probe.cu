// gather_probe.cu — measurement only (not part of the product library): how fast can ANY kernel
// stream a CSR matrix's (col, val) arrays and gather b[col] on this GPU?  The merge-path SpMV's
// distance to this number is what its own bookkeeping (heads, segmented reduction, carries) costs;
// the distance of this number to the algorithmic-byte roofline is what the random gather costs.
#include <cstdint>
#include <cuda_runtime.h>

namespace {
constexpr int kThreads = 256;

template <int kItems, int kMinCtas, bool Gather>
__global__ void __launch_bounds__(kThreads, kMinCtas)
    probe(int64_t nnz, const int* __restrict__ cols, const double* __restrict__ vals, const double* __restrict__ b,
          double* __restrict__ out)
{
    const int64_t k0 = static_cast<int64_t>(blockIdx.x) * (kThreads * kItems);
    int col[kItems];
    double v[kItems], x[kItems];
#pragma unroll
    for (int u = 0; u < kItems; ++u) {
        const int64_t k = k0 + threadIdx.x + u * kThreads;
        const bool in = k < nnz;
        col[u] = in ? __ldcs(cols + k) : 0;
        v[u] = in ? __ldcs(vals + k) : 0.0;
    }
    double acc = 0.0;
    if (Gather) {
#pragma unroll
        for (int u = 0; u < kItems; ++u) x[u] = __ldg(b + col[u]);
#pragma unroll
        for (int u = 0; u < kItems; ++u) acc += v[u] * x[u];
    } else {
#pragma unroll
        for (int u = 0; u < kItems; ++u) acc += v[u] * static_cast<double>(col[u]);
    }
    // one value per thread: ~1/9 of a store per entry, like the row results of the SpMV
    out[static_cast<int64_t>(blockIdx.x) * kThreads + threadIdx.x] = acc;
}
}  // namespace

extern "C" int gather_probe(void* stream, int mode, int64_t nnz, const int* cols, const double* vals, const double* b,
                            double* out)
{
    cudaStream_t s = static_cast<cudaStream_t>(stream);
#define LAUNCH(ITEMS, CTAS, G)                                                                            \
    probe<ITEMS, CTAS, G><<<static_cast<unsigned>((nnz + kThreads * ITEMS - 1) / (kThreads * ITEMS)), kThreads, 0, s>>>( \
        nnz, cols, vals, b, out)
    switch (mode) {
    case 0: LAUNCH(9, 4, false); break;
    case 1: LAUNCH(9, 4, true); break;
    case 2: LAUNCH(9, 6, true); break;
    case 3: LAUNCH(9, 8, true); break;
    case 4: LAUNCH(4, 8, true); break;
    case 5: LAUNCH(16, 3, true); break;
    case 6: LAUNCH(9, 2, true); break;
    default: return 1;
    }
    return cudaGetLastError() == cudaSuccess ? 0 : 2;
}
