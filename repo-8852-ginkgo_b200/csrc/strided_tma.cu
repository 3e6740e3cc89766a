// strided_tma.cu — SELL-P and ELL SpMV with bulk-async (TMA) staging: the same design as
// the CSR row-block kernel (csr_spmv.cu).  A CTA of 128 threads owns 128 consecutive rows;
//  * SELL-P: those rows are whole slices (slice_size 32 / 64 / 128), each slice one
//    contiguous, 16-byte aligned block  [slice_sets[s]*slice_size, slice_sets[s+1]*slice_size)
//    -> one bulk copy per slice and array;
//  * ELL: column i of the 128 rows is contiguous (column-major storage) -> one bulk copy
//    per stored column and array (needs stride*sizeof % 16 == 0).
// Thread t then walks row t out of shared memory with stride-1 (conflict-free) accesses,
// gathers kept in flight in batches, sums in storage order with a rounded product and a
// rounded sum (bit-identical to reference/matrix/{sellp,ell}_kernels.cpp), skipping the
// padding entries (col == -1).  An L2 prefetch of the window one resident wave ahead keeps
// more bytes in flight than shared memory alone can hold.
// Falls back to the thread-per-row kernel of formats_spmv.cu when the layout does not allow
// bulk copies (odd slice sizes, unaligned strides, tiles larger than shared memory).
#include <cuda.h>   // CUtensorMap (types only; the encoder is fetched through cudaGetDriverEntryPoint)

#include "internal.h"
#include "tma.cuh"

namespace gkob200 {
namespace {

constexpr int kRows = 128;

// address of b[col * stride]: one IMAD.WIDE.U32 (32-bit column x 32-bit pitch in bytes + pointer)
template <typename V>
__device__ __forceinline__ const V* b_at(const V* b, int32_t col, uint32_t pitch_bytes)
{
    return reinterpret_cast<const V*>(reinterpret_cast<const char*>(b) +
                                      static_cast<uint64_t>(static_cast<uint32_t>(col)) * pitch_bytes);
}
template <typename V>
__device__ __forceinline__ const V* b_at(const V* b, int64_t col, uint32_t pitch_bytes)
{
    return reinterpret_cast<const V*>(reinterpret_cast<const char*>(b) + col * static_cast<int64_t>(pitch_bytes));
}

// One row out of shared memory: entries idx, idx + step, ... (len of them), summed in storage
// order with a rounded product and a rounded sum; padding entries (col < 0) are skipped.
// Full batches of kBatch entries run without bounds predicates: all
// their gathers are issued before the first add.
template <bool Advanced, int kBatch, typename V, typename I>
__device__ __forceinline__ V row_walk(V acc, const I* s_col, const V* s_val, int idx, int step, int len,
                                      const V* b, uint32_t b_pitch, V alpha)
{
    int i = 0;
    for (; i + kBatch <= len; i += kBatch) {
        I col[kBatch];
        V v[kBatch], xv[kBatch];
#pragma unroll
        for (int u = 0; u < kBatch; ++u) {
            col[u] = s_col[idx + u * step];
            v[u] = s_val[idx + u * step];
        }
#pragma unroll
        for (int u = 0; u < kBatch; ++u) xv[u] = ldg(b_at(b, col[u] < I(0) ? I(0) : col[u], b_pitch));
#pragma unroll
        for (int u = 0; u < kBatch; ++u) {
            // a branch, not a select: it keeps ptxas from sinking the gathers between the adds
            if (col[u] >= I(0))
                acc = Advanced ? add_rn(acc, mul_rn(mul_rn(alpha, v[u]), xv[u])) : add_rn(acc, mul_rn(v[u], xv[u]));
        }
        idx += kBatch * step;
    }
    if (i < len) {
        I col[kBatch];
        V v[kBatch], xv[kBatch];
#pragma unroll
        for (int u = 0; u < kBatch; ++u) {
            const bool in = i + u < len;
            col[u] = in ? s_col[idx + u * step] : I(-1);
            v[u] = in ? s_val[idx + u * step] : V(0);
        }
#pragma unroll
        for (int u = 0; u < kBatch; ++u) xv[u] = ldg(b_at(b, col[u] < I(0) ? I(0) : col[u], b_pitch));
#pragma unroll
        for (int u = 0; u < kBatch; ++u) {
            // a branch, not a select: it keeps ptxas from sinking the gathers between the adds
            if (col[u] >= I(0))
                acc = Advanced ? add_rn(acc, mul_rn(mul_rn(alpha, v[u]), xv[u])) : add_rn(acc, mul_rn(v[u], xv[u]));
        }
    }
    return acc;
}

template <typename V, typename I, bool Advanced, bool Fused, int kBatch>
__global__ void __launch_bounds__(kRows, 4)
    sellp_spmv_tma(int64_t n_rows, int slice_size, const uint64_t* __restrict__ slice_sets,
                   const I* __restrict__ cols, const V* __restrict__ vals, const V* __restrict__ b, uint32_t b_pitch,
                   const V* __restrict__ alpha_p, const V* __restrict__ beta_p, V* __restrict__ c, int64_t c_stride,
                   int cap, SpmvFusion<V> fu, int prefetch_tiles, int64_t total_elems)
{
    extern __shared__ __align__(128) unsigned char smem_raw[];
    V* s_val = reinterpret_cast<V*>(smem_raw);
    I* s_col = reinterpret_cast<I*>(smem_raw + align16(static_cast<size_t>(cap) * sizeof(V)));
    __shared__ __align__(8) uint64_t bar;
    __shared__ int64_t s_set[6];  // slice_sets of the (<= 4) slices of this CTA, +1

    const int tid = threadIdx.x;
    const int spc = kRows / slice_size;  // slices per CTA
    const int64_t slice0 = static_cast<int64_t>(blockIdx.x) * spc;
    const int64_t n_slices = (n_rows + slice_size - 1) / slice_size;
    const int ns = static_cast<int>(min(static_cast<int64_t>(spc), n_slices - slice0));
    const int64_t row = slice0 * slice_size + tid;
    int skip = 0;
    V w_row = V(0);
    if (Fused) {
        if (fu.skip) skip = *fu.skip;
        if (fu.out && row < n_rows) w_row = ldg(fu.w + row);
    }
    if (tid <= ns) s_set[tid] = static_cast<int64_t>(slice_sets[slice0 + tid]);
    if (tid == 0) mbar_init(&bar, 1);
    __syncthreads();
    if (Fused && skip) return;
    const int64_t begin = s_set[0] * slice_size, end = s_set[ns] * slice_size;  // element range of the tile
    if (tid == 0) {
        const unsigned n_el = static_cast<unsigned>(end - begin);
        mbar_expect_tx(&bar, n_el * static_cast<unsigned>(sizeof(V) + sizeof(I)));
        if (n_el) {
            bulk_g2s(s_val, vals + begin, n_el * sizeof(V), &bar);
            bulk_g2s(s_col, cols + begin, n_el * sizeof(I), &bar);
            if (prefetch_tiles > 0) {
                const int64_t ahead = begin + static_cast<int64_t>(prefetch_tiles) * n_el;
                if (ahead + n_el <= total_elems) {
                    bulk_prefetch_l2(vals + ahead, n_el * sizeof(V));
                    bulk_prefetch_l2(cols + ahead, n_el * sizeof(I));
                }
            }
        }
    }
    const int sl = tid / slice_size, local = tid - sl * slice_size;
    const bool live = row < n_rows && sl < ns;
    V alpha = V(1), acc = V(0);
    if (Advanced) {
        alpha = *alpha_p;
        if (live) acc = mul_rn(c[row * c_stride], *beta_p);
    }
    mbar_wait(&bar, 0);
    if (live) {
        const int base = static_cast<int>(s_set[sl] - s_set[0]) * slice_size + local;
        const int len = static_cast<int>(s_set[sl + 1] - s_set[sl]);
        acc = row_walk<Advanced, kBatch>(acc, s_col, s_val, base, slice_size, len, b, b_pitch, alpha);
        c[row * c_stride] = acc;
    }
    if (Fused && fu.out) {
        if (fu.out_sq)
            store_block_partial2(live ? acc * w_row : V(0), live ? acc * acc : V(0), ws_partials<V>(fu.ws));
        else
            store_block_partial(live ? acc * w_row : V(0), ws_partials<V>(fu.ws));
    }
}

template <typename V, typename I, bool Advanced, bool Fused, int kBatch>
__global__ void __launch_bounds__(kRows, 4)
    ell_spmv_tma(int64_t n_rows, int64_t stride, int width, const I* __restrict__ cols, const V* __restrict__ vals,
                 const V* __restrict__ b, uint32_t b_pitch, const V* __restrict__ alpha_p, const V* __restrict__ beta_p,
                 V* __restrict__ c, int64_t c_stride, SpmvFusion<V> fu, int prefetch_tiles)
{
    extern __shared__ __align__(128) unsigned char smem_raw[];
    V* s_val = reinterpret_cast<V*>(smem_raw);
    I* s_col = reinterpret_cast<I*>(smem_raw + align16(static_cast<size_t>(width) * kRows * sizeof(V)));
    __shared__ __align__(8) uint64_t bar;

    const int tid = threadIdx.x;
    const int64_t row0 = static_cast<int64_t>(blockIdx.x) * kRows;
    const int nrow = static_cast<int>(min(static_cast<int64_t>(kRows), n_rows - row0));
    const int64_t row = row0 + tid;
    int skip = 0;
    V w_row = V(0);
    if (Fused) {
        if (fu.skip) skip = *fu.skip;
        if (fu.out && tid < nrow) w_row = ldg(fu.w + row);
    }
    if (tid == 0) mbar_init(&bar, 1);
    __syncthreads();
    if (Fused && skip) return;
    // rows of the last, partial tile: round the copy length up to 16 bytes (stays inside the
    // column because stride >= n_rows rounded to the same granularity is checked on the host)
    constexpr int VA = 16 / sizeof(V), IA = 16 / sizeof(I);
    const unsigned nv = static_cast<unsigned>((nrow + VA - 1) / VA * VA), ni = static_cast<unsigned>((nrow + IA - 1) / IA * IA);
    if (tid == 0) {
        mbar_expect_tx(&bar, static_cast<unsigned>(width) * (nv * sizeof(V) + ni * sizeof(I)));
        for (int i = 0; i < width; ++i) {
            bulk_g2s(s_val + i * kRows, vals + row0 + i * stride, nv * sizeof(V), &bar);
            bulk_g2s(s_col + i * kRows, cols + row0 + i * stride, ni * sizeof(I), &bar);
        }
        if (prefetch_tiles > 0) {
            const int64_t ahead = row0 + static_cast<int64_t>(prefetch_tiles) * kRows;
            if (ahead + kRows <= n_rows)
                for (int i = 0; i < width; ++i) {
                    bulk_prefetch_l2(vals + ahead + i * stride, kRows * sizeof(V));
                    bulk_prefetch_l2(cols + ahead + i * stride, kRows * sizeof(I));
                }
        }
    }
    const bool live = tid < nrow;
    V alpha = V(1), acc = V(0);
    if (Advanced) {
        alpha = *alpha_p;
        if (live) acc = mul_rn(c[row * c_stride], *beta_p);
    }
    mbar_wait(&bar, 0);
    if (live) {
        acc = row_walk<Advanced, kBatch>(acc, s_col, s_val, tid, kRows, width, b, b_pitch, alpha);
        c[row * c_stride] = acc;
    }
    if (Fused && fu.out) {
        if (fu.out_sq)
            store_block_partial2(live ? acc * w_row : V(0), live ? acc * acc : V(0), ws_partials<V>(fu.ws));
        else
            store_block_partial(live ? acc * w_row : V(0), ws_partials<V>(fu.ws));
    }
}

// ---------------------------------------------------------------------------
// ELL with ONE 2-D tensor-map copy per array (cp.async.bulk.tensor.2d, SASS UTMALDG): the tile of
// 128 rows x `width` stored columns of the column-major value / index arrays arrives with two
// instructions instead of 2 x width 1 KB bulk copies issued by one thread (the round-1 kernel:
// 54 copies + 54 prefetches per 27-column tile, 534 us on the 27-pt 200^3 matrix against 398 us
// for CSR on the same matrix).  Rows past the end of the matrix are zero-filled by the TMA unit.
// ---------------------------------------------------------------------------
__device__ __forceinline__ void tma_load_2d(void* dst, const CUtensorMap* map, int c0, int c1, uint64_t* bar)
{
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];" ::"r"(
            smem_u32(dst)),
        "l"(map), "r"(c0), "r"(c1), "r"(smem_u32(bar))
        : "memory");
}
__device__ __forceinline__ void tma_prefetch_2d(const CUtensorMap* map, int c0, int c1)
{
    asm volatile("cp.async.bulk.prefetch.tensor.2d.L2.global.tile [%0, {%1, %2}];" ::"l"(map), "r"(c0), "r"(c1) : "memory");
}

template <typename V, typename I, bool Advanced, bool Fused, int kBatch>
__global__ void __launch_bounds__(kRows, 4)
    ell_spmv_tma2d(int64_t n_rows, int width, const __grid_constant__ CUtensorMap tm_val,
                   const __grid_constant__ CUtensorMap tm_col, const V* __restrict__ b, uint32_t b_pitch,
                   const V* __restrict__ alpha_p, const V* __restrict__ beta_p, V* __restrict__ c, int64_t c_stride,
                   SpmvFusion<V> fu, int prefetch_tiles)
{
    extern __shared__ __align__(128) unsigned char smem_raw[];
    V* s_val = reinterpret_cast<V*>(smem_raw);
    I* s_col = reinterpret_cast<I*>(smem_raw + static_cast<size_t>(width) * kRows * sizeof(V));
    __shared__ __align__(8) uint64_t bar;

    const int tid = threadIdx.x;
    const int64_t row0 = static_cast<int64_t>(blockIdx.x) * kRows;
    const int nrow = static_cast<int>(min(static_cast<int64_t>(kRows), n_rows - row0));
    const int64_t row = row0 + tid;
    int skip = 0;
    if (Fused && fu.skip) skip = *fu.skip;
    if (tid == 0) mbar_init(&bar, 1);
    __syncthreads();
    if (Fused && skip) return;
    if (tid == 0) {
        mbar_expect_tx(&bar, static_cast<unsigned>(width) * kRows * static_cast<unsigned>(sizeof(V) + sizeof(I)));
        tma_load_2d(s_val, &tm_val, static_cast<int>(row0), 0, &bar);
        tma_load_2d(s_col, &tm_col, static_cast<int>(row0), 0, &bar);
        if (prefetch_tiles > 0) {
            const int64_t ahead = row0 + static_cast<int64_t>(prefetch_tiles) * kRows;
            if (ahead + kRows <= n_rows) {
                tma_prefetch_2d(&tm_val, static_cast<int>(ahead), 0);
                tma_prefetch_2d(&tm_col, static_cast<int>(ahead), 0);
            }
        }
    }
    const bool live = tid < nrow;
    V alpha = V(1), acc = V(0);
    if (Advanced) {
        alpha = *alpha_p;
        if (live) acc = mul_rn(c[row * c_stride], *beta_p);
    }
    mbar_wait(&bar, 0);
    if (live) {
        acc = row_walk<Advanced, kBatch>(acc, s_col, s_val, tid, kRows, width, b, b_pitch, alpha);
        c[row * c_stride] = acc;
    }
    if (Fused && fu.out) {
        const V w_row = live ? ldg(fu.w + row) : V(0);
        if (fu.out_sq)
            store_block_partial2(live ? acc * w_row : V(0), live ? acc * acc : V(0), ws_partials<V>(fu.ws));
        else
            store_block_partial(live ? acc * w_row : V(0), ws_partials<V>(fu.ws));
    }
}

// Tensor maps over the column-major ELL arrays: dims {stride (rows, contiguous), width}, box
// {128, width}.  Encoded on the host (cuTensorMapEncodeTiled through the runtime's driver entry
// point: no link-time dependency on libcuda) and cached per thread for the last few matrices.
using EncodeFn = CUresult (*)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                              const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                              CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
inline EncodeFn tensor_map_encoder()
{
    static EncodeFn fn = [] {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) != cudaSuccess ||
            q != cudaDriverEntryPointSuccess)
            p = nullptr;
        cudaGetLastError();
        return reinterpret_cast<EncodeFn>(p);
    }();
    return fn;
}
template <typename T>
CUtensorMapDataType tm_type();
template <> inline CUtensorMapDataType tm_type<double>() { return CU_TENSOR_MAP_DATA_TYPE_FLOAT64; }
template <> inline CUtensorMapDataType tm_type<float>() { return CU_TENSOR_MAP_DATA_TYPE_FLOAT32; }
template <> inline CUtensorMapDataType tm_type<int32_t>() { return CU_TENSOR_MAP_DATA_TYPE_INT32; }
template <> inline CUtensorMapDataType tm_type<int64_t>() { return CU_TENSOR_MAP_DATA_TYPE_INT64; }

template <typename T>
bool ell_tensor_map(const T* base, int64_t stride, int64_t width, CUtensorMap* out)
{
    struct Key {
        const void* base;
        int64_t stride, width;
        CUtensorMap map;
    };
    thread_local Key cache[8];
    thread_local int next = 0;
    for (const Key& k : cache)
        if (k.base == base && k.stride == stride && k.width == width) {
            *out = k.map;
            return true;
        }
    EncodeFn enc = tensor_map_encoder();
    if (!enc) return false;
    const cuuint64_t dims[2] = {static_cast<cuuint64_t>(stride), static_cast<cuuint64_t>(width)};
    const cuuint64_t strides[1] = {static_cast<cuuint64_t>(stride) * sizeof(T)};
    const cuuint32_t box[2] = {static_cast<cuuint32_t>(kRows), static_cast<cuuint32_t>(width)};
    const cuuint32_t estr[2] = {1, 1};
    if (enc(out, tm_type<T>(), 2, const_cast<T*>(base), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
            CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS)
        return false;
    cache[next] = Key{base, stride, width, *out};
    next = (next + 1) % 8;
    return true;
}

// gathers kept in flight per thread for rows wider than 18 entries: 14 (like the CSR row-block
// kernel on the 27-pt stencil) or 9; GKOB200_STRIDED_BATCH=9 keeps the round-1 setting (A/B)
inline bool strided_batch14()
{
    static const bool v = [] {
        const char* e = getenv("GKOB200_STRIDED_BATCH");
        return !(e && atoi(e) == 9);
    }();
    return v;
}

inline int resident_ctas(size_t smem)
{
    int r = static_cast<int>((227 * 1024) / (smem + 1024));
    return r > 16 ? 16 : (r < 1 ? 1 : r);
}

constexpr size_t kMaxTileBytes = 100 * 1024;  // beyond this the thread-per-row kernel takes over

}  // namespace

// returns 1 if the launch was done, 0 if the caller must fall back, <0 / >0 on error
template <typename V, typename I>
int sellp_spmv_tma_launch(cudaStream_t s, int64_t n_rows, int64_t slice_size, const uint64_t* slice_sets,
                          int64_t max_slice_len, int64_t total_cols, const I* cols, const V* vals, const V* b,
                          int64_t b_stride, const V* alpha, const V* beta, V* c, int64_t c_stride,
                          const SpmvFusion<V>* fusion)
{
    if (slice_size != 32 && slice_size != 64 && slice_size != 128) return 0;
    if (max_slice_len <= 0 || reinterpret_cast<uintptr_t>(vals) % 16 || reinterpret_cast<uintptr_t>(cols) % 16) return 0;
    if (b_stride * sizeof(V) > 0xffffffffull) return 0;
    const uint32_t pitch = static_cast<uint32_t>(b_stride * sizeof(V));
    const int cap = static_cast<int>(max_slice_len * kRows);
    const size_t smem = align16(static_cast<size_t>(cap) * sizeof(V)) + align16(static_cast<size_t>(cap) * sizeof(I));
    if (smem > kMaxTileBytes) return 0;
    const int spc = kRows / static_cast<int>(slice_size);
    const int64_t n_slices = ceildiv(n_rows, slice_size);
    const unsigned grid = static_cast<unsigned>(ceildiv(n_slices, spc));
    SpmvFusion<V> fu;
    if (fusion) fu = *fusion;
    const bool fused = fusion != nullptr, adv = alpha != nullptr;
    if (fused && fu.out && static_cast<int64_t>(grid) * (fu.out_sq ? 2 : 1) > fu.ws_blocks) return GKOB200_EWORKSPACE;
    const int pf = sm_count() * resident_ctas(smem);
    const int64_t total = total_cols * slice_size;
#define GKOB200_SP(ADV, FUSED, BATCH)                                                                             \
    {                                                                                                             \
        auto kern = sellp_spmv_tma<V, I, ADV, FUSED, BATCH>;                                                       \
        if (smem > 48 * 1024) cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kMaxTileBytes); \
        kern<<<grid, kRows, smem, s>>>(n_rows, static_cast<int>(slice_size), slice_sets, cols, vals, b, pitch,     \
                                       alpha, beta, c, c_stride, cap, fu, pf, total);                              \
    }
    const bool wide = max_slice_len > 8;
    if (max_slice_len > 18 && strided_batch14()) {
        if (adv && fused) GKOB200_SP(true, true, 14) else if (adv) GKOB200_SP(true, false, 14)
        else if (fused) GKOB200_SP(false, true, 14) else GKOB200_SP(false, false, 14)
    } else if (wide) {
        if (adv && fused) GKOB200_SP(true, true, 9) else if (adv) GKOB200_SP(true, false, 9)
        else if (fused) GKOB200_SP(false, true, 9) else GKOB200_SP(false, false, 9)
    } else {
        if (adv && fused) GKOB200_SP(true, true, 7) else if (adv) GKOB200_SP(true, false, 7)
        else if (fused) GKOB200_SP(false, true, 7) else GKOB200_SP(false, false, 7)
    }
#undef GKOB200_SP
    GKOB200_CHECK_LAUNCH();
    if (fused && fu.out) {
        const int frc = launch_finish_partials<V>(s, static_cast<int64_t>(grid), fu);
        if (frc) return frc;
    }
    return 1;
}

template <typename V, typename I>
int ell_spmv_tma_launch(cudaStream_t s, int64_t n_rows, int64_t stride, int64_t width, const I* cols, const V* vals,
                        const V* b, int64_t b_stride, const V* alpha, const V* beta, V* c, int64_t c_stride,
                        const SpmvFusion<V>* fusion)
{
    if (width <= 0 || width > 64 || b_stride * sizeof(V) > 0xffffffffull) return 0;
    const uint32_t pitch = static_cast<uint32_t>(b_stride * sizeof(V));
    if ((stride * sizeof(V)) % 16 || (stride * sizeof(I)) % 16 || reinterpret_cast<uintptr_t>(vals) % 16 ||
        reinterpret_cast<uintptr_t>(cols) % 16)
        return 0;
    // the padded copy of the last tile must stay inside a column
    if (ceildiv(n_rows, 4) * 4 > stride) return 0;
    const size_t smem = align16(static_cast<size_t>(width) * kRows * sizeof(V)) + align16(static_cast<size_t>(width) * kRows * sizeof(I));
    if (smem > kMaxTileBytes) return 0;
    const unsigned grid = static_cast<unsigned>(ceildiv(n_rows, kRows));
    SpmvFusion<V> fu;
    if (fusion) fu = *fusion;
    const bool fused = fusion != nullptr, adv = alpha != nullptr;
    if (fused && fu.out && static_cast<int64_t>(grid) * (fu.out_sq ? 2 : 1) > fu.ws_blocks) return GKOB200_EWORKSPACE;
    const int pf = sm_count() * resident_ctas(smem);
    // one 2-D tensor-map copy per array when the encoder is available (GKOB200_ELL_TMA2D=0: the
    // per-column bulk copies, for A/B on the box)
    static const bool use_2d = [] {
        const char* e = getenv("GKOB200_ELL_TMA2D");
        return !(e && e[0] == '0');
    }();
    CUtensorMap tm_val, tm_col;
    if (use_2d && width <= 256 && n_rows < (int64_t(1) << 31) && (static_cast<size_t>(width) * kRows * sizeof(V)) % 128 == 0 &&
        ell_tensor_map(vals, stride, width, &tm_val) && ell_tensor_map(cols, stride, width, &tm_col)) {
#define GKOB200_EL2(ADV, FUSED, BATCH)                                                                            \
    {                                                                                                             \
        auto kern = ell_spmv_tma2d<V, I, ADV, FUSED, BATCH>;                                                       \
        if (smem > 48 * 1024) cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kMaxTileBytes); \
        kern<<<grid, kRows, smem, s>>>(n_rows, static_cast<int>(width), tm_val, tm_col, b, pitch, alpha, beta, c,   \
                                       c_stride, fu, pf);                                                          \
    }
        if (width > 18 && strided_batch14()) {
            if (adv && fused) GKOB200_EL2(true, true, 14) else if (adv) GKOB200_EL2(true, false, 14)
            else if (fused) GKOB200_EL2(false, true, 14) else GKOB200_EL2(false, false, 14)
        } else if (width > 8) {
            if (adv && fused) GKOB200_EL2(true, true, 9) else if (adv) GKOB200_EL2(true, false, 9)
            else if (fused) GKOB200_EL2(false, true, 9) else GKOB200_EL2(false, false, 9)
        } else {
            if (adv && fused) GKOB200_EL2(true, true, 7) else if (adv) GKOB200_EL2(true, false, 7)
            else if (fused) GKOB200_EL2(false, true, 7) else GKOB200_EL2(false, false, 7)
        }
#undef GKOB200_EL2
        GKOB200_CHECK_LAUNCH();
        if (fused && fu.out) {
            const int frc = launch_finish_partials<V>(s, static_cast<int64_t>(grid), fu);
            if (frc) return frc;
        }
        return 1;
    }
#define GKOB200_EL(ADV, FUSED, BATCH)                                                                             \
    {                                                                                                             \
        auto kern = ell_spmv_tma<V, I, ADV, FUSED, BATCH>;                                                         \
        if (smem > 48 * 1024) cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kMaxTileBytes); \
        kern<<<grid, kRows, smem, s>>>(n_rows, stride, static_cast<int>(width), cols, vals, b, pitch, alpha,       \
                                       beta, c, c_stride, fu, pf);                                                 \
    }
    if (width > 8) {
        if (adv && fused) GKOB200_EL(true, true, 9) else if (adv) GKOB200_EL(true, false, 9)
        else if (fused) GKOB200_EL(false, true, 9) else GKOB200_EL(false, false, 9)
    } else {
        if (adv && fused) GKOB200_EL(true, true, 7) else if (adv) GKOB200_EL(true, false, 7)
        else if (fused) GKOB200_EL(false, true, 7) else GKOB200_EL(false, false, 7)
    }
#undef GKOB200_EL
    GKOB200_CHECK_LAUNCH();
    if (fused && fu.out) {
        const int frc = launch_finish_partials<V>(s, static_cast<int64_t>(grid), fu);
        if (frc) return frc;
    }
    return 1;
}

#define GKOB200_INST(V, I)                                                                                         \
    template int sellp_spmv_tma_launch<V, I>(cudaStream_t, int64_t, int64_t, const uint64_t*, int64_t, int64_t,    \
                                             const I*, const V*, const V*, int64_t, const V*, const V*, V*,        \
                                             int64_t, const SpmvFusion<V>*);                                       \
    template int ell_spmv_tma_launch<V, I>(cudaStream_t, int64_t, int64_t, int64_t, const I*, const V*, const V*,  \
                                           int64_t, const V*, const V*, V*, int64_t, const SpmvFusion<V>*);
GKOB200_INST(double, int32_t)
GKOB200_INST(float, int32_t)
GKOB200_INST(double, int64_t)
GKOB200_INST(float, int64_t)
#undef GKOB200_INST

}  // namespace gkob200
