// formats_spmv.cu — ELL, SELL-P and COO SpMV / SpMM for sm_100a.
//
// Replaces gko::kernels::cuda::{ell,sellp,coo}::{spmv,advanced_spmv,spmv2,advanced_spmv2}
// (reference cuda/matrix/{ell,sellp,coo}_kernels.cu; device code
// common/cuda_hip/matrix/{ell,sellp,coo}_kernels.hpp.inc) and follows the arithmetic of
// the oracle reference/matrix/{ell,sellp,coo}_kernels.cpp.
//
// ELL / SELL-P: one thread per row; both formats store a row's entries strided by the
// (slice) height, so the col/val streams of a warp are fully coalesced 128/256-byte
// lines and, for banded matrices, so are the gathers x[col].  A thread issues a batch
// of independent loads before it starts adding, and adds in storage order with a
// rounded product + rounded sum: bit-identical to the reference executor.  Padding
// entries (col == -1) are skipped like the reference does.
//
// Multi-RHS (SpMM): spmm.cuh (warp tiles walked row by row, lane = RHS column).
//
// COO (accumulating spmv2, row-sorted): nnz are split evenly over CTAs and threads;
// products are staged coalesced into shared memory, each thread reduces its 9 entries in
// registers, row pieces are stitched in thread order inside the CTA, row sums leave the CTA
// as one coalesced read-modify-write of c, per-CTA carries by a small fix-up kernel.  No
// atomics (the reference's coo kernel uses atomic_add:
// common/cuda_hip/matrix/coo_kernels.hpp.inc:54-218), deterministic.
//
// Algorithmic bytes: ELL n*w*(V+I) + (n_cols+n)*k*V;
// SELL-P S*(V+I) + (ns+1)*8 + (n_cols+n)*k*V; COO nnz*(V+2I) + n_cols*k*V + 2*n*k*V.
#include "internal.h"
#include "p2p.cuh"
#include "spmm.cuh"
#include "tma.cuh"

namespace gkob200 {
namespace {

constexpr int kBatch = 9;

// ---------------------------------------------------------------------------
// thread-per-row kernel shared by ELL and SELL-P.  `Fmt` gives, for a row, the
// position of its first stored entry, the distance between consecutive entries
// and the number of stored entries.
// ---------------------------------------------------------------------------
struct EllFmt {
    int64_t stride, width;
    __device__ __forceinline__ void row(int64_t r, int64_t& first, int64_t& step, int64_t& len) const
    {
        first = r;
        step = stride;
        len = width;
    }
};
struct SellpFmt {
    int64_t slice_size;
    const uint64_t* slice_sets;
    const uint64_t* slice_lengths;
    __device__ __forceinline__ void row(int64_t r, int64_t& first, int64_t& step, int64_t& len) const
    {
        const int64_t slice = r / slice_size;
        first = static_cast<int64_t>(slice_sets[slice]) * slice_size + (r - slice * slice_size);
        step = slice_size;
        len = static_cast<int64_t>(slice_lengths[slice]);
    }
};

template <typename V, typename I, typename Fmt, bool Advanced, bool Fused>
__global__ void __launch_bounds__(256)
    strided_spmv(int64_t n_rows, Fmt fmt, const I* __restrict__ cols, const V* __restrict__ vals,
                 const V* __restrict__ b, int64_t b_stride, const V* __restrict__ alpha_p,
                 const V* __restrict__ beta_p, V* __restrict__ c, int64_t c_stride, SpmvFusion<V> fu)
{
    if (Fused && fu.skip && *fu.skip) return;
    const int64_t r = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x;
    const bool live = r < n_rows;
    V acc = V(0), alpha = V(1);
    if (live) {
        int64_t first, step, len;
        fmt.row(r, first, step, len);
        if (Advanced) {
            alpha = *alpha_p;
            acc = mul_rn(c[r * c_stride], *beta_p);
        }
        for (int64_t i = 0; i < len; i += kBatch) {
            V v[kBatch], xv[kBatch];
            I col[kBatch];
#pragma unroll
            for (int u = 0; u < kBatch; ++u) {
                const bool in = i + u < len;
                col[u] = in ? cols[first + (i + u) * step] : I(-1);
                v[u] = in ? vals[first + (i + u) * step] : V(0);
            }
#pragma unroll
            for (int u = 0; u < kBatch; ++u)
                // unconditional gather (padding reads x[0] and is discarded below): keeps the
                // batch free of branches so that all loads are issued before the first add
                xv[u] = ldg(b + static_cast<int64_t>(col[u] < I(0) ? I(0) : col[u]) * b_stride);
#pragma unroll
            for (int u = 0; u < kBatch; ++u) {
                const V nxt = Advanced ? add_rn(acc, mul_rn(mul_rn(alpha, v[u]), xv[u]))
                                       : add_rn(acc, mul_rn(v[u], xv[u]));
                acc = col[u] != I(-1) ? nxt : acc;
            }
        }
        c[r * c_stride] = acc;
    }
    if (Fused && fu.out) {
        if (fu.out_sq)
            store_block_partial2(live ? acc * fu.w[r] : V(0), live ? acc * acc : V(0), ws_partials<V>(fu.ws));
        else
            store_block_partial(live ? acc * fu.w[r] : V(0), ws_partials<V>(fu.ws));
    }
}

template <typename V, typename I, typename Fmt>
int strided_launch(cudaStream_t s, int64_t n_rows, Fmt fmt, const I* cols, const V* vals, const V* b,
                   int64_t b_stride, int64_t nrhs, const V* alpha, const V* beta, V* c, int64_t c_stride,
                   const SpmvFusion<V>* fusion)
{
    const bool adv = alpha != nullptr;
    if (nrhs > 1) {
        if (fusion && fusion->out) return GKOB200_EUNSUPPORTED;
        spmm::StridedStager<V, I, Fmt, spmm::Default> st{fmt, cols, vals, n_rows, 0, 0, 0, 0};
        return spmm::launch_cfg<V, I, decltype(st), spmm::Default>(s, n_rows, st, b, b_stride, nrhs, alpha, beta, c,
                                                                   c_stride);
    }
    SpmvFusion<V> fu;
    if (fusion) fu = *fusion;
    const bool fused = fusion != nullptr;
    const unsigned grid = static_cast<unsigned>(ceildiv(n_rows, 256));
    if (fused && fu.out && static_cast<int64_t>(grid) * (fu.out_sq ? 2 : 1) > fu.ws_blocks) return GKOB200_EWORKSPACE;
#define GKOB200_ST(ADV, FUSED) \
    strided_spmv<V, I, Fmt, ADV, FUSED><<<grid, 256, 0, s>>>(n_rows, fmt, cols, vals, b, b_stride, alpha, beta, c, c_stride, fu)
    if (adv && fused) GKOB200_ST(true, true);
    else if (adv) GKOB200_ST(true, false);
    else if (fused) GKOB200_ST(false, true);
    else GKOB200_ST(false, false);
#undef GKOB200_ST
    GKOB200_CHECK_LAUNCH();
    if (fused && fu.out) {
        const int frc = launch_finish_partials<V>(s, static_cast<int64_t>(grid), fu);
        if (frc) return frc;
    }
    return 0;
}

// ---------------------------------------------------------------------------
// COO  c += [alpha] A b   (row-sorted entries)
// ---------------------------------------------------------------------------
constexpr int kCooThreads = 256;
constexpr int kCooItems = 9;   // odd: the per-thread windows are bank-conflict free
constexpr int kCooTile = kCooThreads * kCooItems;

// Row-sorted COO, accumulating (c += A b): a segmented reduction over equal-size tiles of
// entries, the same scheme as the merge-path CSR kernel (csr_spmv.cu):
//   1. every thread issues all its (row, col, val) loads, then all its gathers (9 independent
//      requests in flight), products and rows go to shared memory;
//   2. thread t reduces its 9 consecutive entries in registers; a segment that starts and
//      ends inside the window is complete;
//   3. segments spanning threads are stitched in thread order (left-to-right association);
//   4. the segment touching the first / last entry of the tile may continue in a neighbouring
//      tile: exported as carries (2 slots per tile, summed in order by coo_fixup);
//   5. all other row sums are collected in shared memory (the tile's rows are a contiguous
//      range when the matrix has no long runs of empty rows) and added to c as one coalesced
//      read-modify-write; tiles spanning more rows than fit update c directly.
// No atomics (the reference's kernel uses atomic_add: coo_kernels.hpp.inc:54-218), deterministic.
// Bulk: the three streams of a FULL tile (16-byte aligned windows) arrive through three bulk
// asynchronous copies (TMA, UBLKCP) issued by one thread, with an L2 prefetch of the tile one
// resident wave ahead — the staging of the CSR row-block kernel; the products then overwrite
// the staged values in place.  The last, partial tile (and unaligned arrays) take the
// register-staged path (`tile0` = index of the first tile of this launch).
template <typename V, typename I, bool Advanced, bool Bulk>
__global__ void __launch_bounds__(kCooThreads)
    coo_spmv2(int64_t nnz, const I* __restrict__ rows, const I* __restrict__ cols, const V* __restrict__ vals,
              const V* __restrict__ b, int64_t b_stride, int64_t j, const V* __restrict__ alpha_p,
              V* __restrict__ c, int64_t c_stride, int64_t* __restrict__ carry_row, V* __restrict__ carry_val,
              int64_t tile0, int prefetch_tiles)
{
    extern __shared__ __align__(128) unsigned char coo_smem[];
    V* s_prod = reinterpret_cast<V*>(coo_smem);
    I* s_row = reinterpret_cast<I*>(coo_smem + kCooTile * sizeof(V));
    I* s_col = s_row + kCooTile;                 // (Bulk only)
    __shared__ unsigned char s_flag[kCooTile];   // row (relative to the tile's first row) received a sum
    __shared__ V s_lead[kCooThreads];            // sum of a thread's entries before its first head
    __shared__ bool s_has[kCooThreads];          // the thread's window contains a head
    __shared__ __align__(8) uint64_t bar;

    const int tid = threadIdx.x;
    const int64_t tile = tile0 + blockIdx.x;
    const int64_t k0 = tile * kCooTile;
    const int len = static_cast<int>(min(static_cast<int64_t>(kCooTile), nnz - k0));
    V alpha = V(1);
    if (Advanced) alpha = *alpha_p;
    if (Bulk) {
        if (tid == 0) mbar_init(&bar, 1);
        __syncthreads();
        if (tid == 0) {
            mbar_expect_tx(&bar, kCooTile * static_cast<unsigned>(sizeof(V) + 2 * sizeof(I)));
            bulk_g2s(s_prod, vals + k0, kCooTile * sizeof(V), &bar);
            bulk_g2s(s_col, cols + k0, kCooTile * sizeof(I), &bar);
            bulk_g2s(s_row, rows + k0, kCooTile * sizeof(I), &bar);
            const int64_t ahead = k0 + static_cast<int64_t>(prefetch_tiles) * kCooTile;
            if (prefetch_tiles > 0 && ahead + kCooTile <= nnz) {
                bulk_prefetch_l2(vals + ahead, kCooTile * sizeof(V));
                bulk_prefetch_l2(cols + ahead, kCooTile * sizeof(I));
                bulk_prefetch_l2(rows + ahead, kCooTile * sizeof(I));
            }
        }
        mbar_wait(&bar, 0);
        V v[kCooItems], xv[kCooItems];
#pragma unroll
        for (int u = 0; u < kCooItems; ++u) {
            const int k = tid + u * kCooThreads;
            v[u] = s_prod[k];
            xv[u] = ldg(b + static_cast<int64_t>(s_col[k]) * b_stride + j);
        }
#pragma unroll
        for (int u = 0; u < kCooItems; ++u) {
            const int k = tid + u * kCooThreads;
            s_prod[k] = Advanced ? mul_rn(mul_rn(alpha, v[u]), xv[u]) : mul_rn(v[u], xv[u]);
            s_flag[k] = 0;
        }
    } else {
        V v[kCooItems], xv[kCooItems];
        I col[kCooItems], row[kCooItems];
#pragma unroll
        for (int u = 0; u < kCooItems; ++u) {
            const int k = tid + u * kCooThreads;
            const bool in = k < len;
            col[u] = in ? cols[k0 + k] : I(0);
            v[u] = in ? vals[k0 + k] : V(0);
            row[u] = in ? rows[k0 + k] : I(-1);
        }
#pragma unroll
        for (int u = 0; u < kCooItems; ++u) xv[u] = ldg(b + static_cast<int64_t>(col[u]) * b_stride + j);
#pragma unroll
        for (int u = 0; u < kCooItems; ++u) {
            const int k = tid + u * kCooThreads;
            s_prod[k] = k < len ? (Advanced ? mul_rn(mul_rn(alpha, v[u]), xv[u]) : mul_rn(v[u], xv[u])) : V(0);
            s_row[k] = row[u];
            s_flag[k] = 0;
        }
    }
    __syncthreads();
    const I row_first = s_row[0], row_last = s_row[len - 1];
    // rows of the tile relative to row_first fit the shared result array?
    const bool dense = static_cast<int64_t>(row_last) - static_cast<int64_t>(row_first) < kCooTile;
    V p[kCooItems];
    I r[kCooItems];
    const int begin = tid * kCooItems;
#pragma unroll
    for (int u = 0; u < kCooItems; ++u) {
        p[u] = s_prod[begin + u];
        r[u] = s_row[begin + u];
    }
    const I prev_row = tid > 0 ? s_row[begin - 1] : I(-1);
    __syncthreads();   // s_prod is reused for the row results from here on
    V* s_out = s_prod;
    if (tid == 0) {
        carry_row[2 * tile] = carry_row[2 * tile + 1] = -1;
        carry_val[2 * tile] = carry_val[2 * tile + 1] = V(0);
    }
    // a finished row sum: exported if it touches the tile's first entry, otherwise collected
    auto deliver = [&](I row, V sum, bool first_seg, bool reaches_end) {
        if (first_seg) {
            carry_row[2 * tile] = row;
            carry_val[2 * tile] = sum;
        } else if (reaches_end) {
            carry_row[2 * tile + 1] = row;
            carry_val[2 * tile + 1] = sum;
        } else if (dense) {
            s_out[row - row_first] = sum;
            s_flag[row - row_first] = 1;
        } else {
            c[static_cast<int64_t>(row) * c_stride + j] = add_rn(c[static_cast<int64_t>(row) * c_stride + j], sum);
        }
    };
    V lead = V(0), acc = V(0);
    bool has = false, seg_first = false;
    I seg_row = I(-1);
    // (the carry defaults above are ordered before every deliver() of another thread by the
    // barrier below; thread 0's own in-window deliveries follow them in program order)
    __syncthreads();
#pragma unroll
    for (int u = 0; u < kCooItems; ++u) {
        const bool valid = begin + u < len;
        const bool head = valid && (u == 0 ? (tid == 0 || r[0] != prev_row) : r[u] != r[u - 1]);
        if (head) {
            if (!has)
                lead = acc;
            else
                deliver(seg_row, acc, seg_first, false);   // starts and ends inside this window
            has = true;
            seg_row = r[u];
            seg_first = tid == 0 && u == 0;
            acc = p[u];
        } else {
            acc = add_rn(acc, p[u]);   // entries past the end of the tile are zeros
        }
    }
    if (!has) lead = acc;
    s_lead[tid] = lead;
    s_has[tid] = has;
    __syncthreads();
    if (has) {
        // leader of the row of my last head: my trailing sum + the leading sums of the following
        // threads up to and including the next thread with a head
        V run = acc;
        int t = tid + 1;
        bool ended = false;
        while (t < kCooThreads && !ended) {
            run = add_rn(run, s_lead[t]);
            ended = s_has[t];
            ++t;
        }
        deliver(seg_row, run, seg_first, !ended);
    }
    if (!dense) return;
    __syncthreads();
    const int span = static_cast<int>(row_last - row_first) + 1;
    for (int q = tid; q < span; q += kCooThreads) {
        if (s_flag[q]) {
            const int64_t at = (static_cast<int64_t>(row_first) + q) * c_stride + j;
            c[at] = add_rn(c[at], s_out[q]);
        }
    }
}

// carries: 2 slots per tile (first-row piece, last-row piece), in tile order; runs of
// equal rows are added in order by their leader.  Deterministic, no atomics.
template <typename V>
__global__ void coo_fixup(int64_t n_slots, const int64_t* __restrict__ carry_row, const V* __restrict__ carry_val,
                          V* __restrict__ c, int64_t c_stride, int64_t j)
{
    const int64_t t = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x;
    if (t >= n_slots) return;
    const int64_t row = carry_row[t];
    if (row < 0) return;
    // previous non-empty slot
    int64_t p = t - 1;
    while (p >= 0 && carry_row[p] < 0) --p;
    if (p >= 0 && carry_row[p] == row) return;
    V run = carry_val[t];
    for (int64_t u = t + 1; u < n_slots; ++u) {
        if (carry_row[u] < 0) continue;
        if (carry_row[u] != row) break;
        run = add_rn(run, carry_val[u]);
    }
    c[row * c_stride + j] = add_rn(c[row * c_stride + j], run);
}

template <typename V, typename I>
int coo_spmv2_launch(cudaStream_t s, int64_t n_rows, int64_t nnz, const I* rows, const I* cols, const V* vals,
                     const V* b, int64_t b_stride, int64_t nrhs, const V* alpha, V* c, int64_t c_stride,
                     void* workspace, size_t workspace_bytes)
{
    if (nnz == 0 || nrhs == 0 || n_rows == 0) return 0;
    const int64_t n_tiles = ceildiv(nnz, kCooTile);
    const size_t need = gkob200_coo_spmv_workspace_bytes(nnz, sizeof(V));
    if (!workspace || workspace_bytes < need) return GKOB200_EWORKSPACE;
    int64_t* carry_row = reinterpret_cast<int64_t*>(workspace);
    V* carry_val = reinterpret_cast<V*>(carry_row + 2 * n_tiles);
    // full tiles of 16-byte aligned arrays are staged by bulk copies (GKOB200_COO_BULK=0: A/B)
    static const bool bulk_on = [] {
        const char* e = getenv("GKOB200_COO_BULK");
        return !(e && e[0] == '0');
    }();
    const bool aligned = reinterpret_cast<uintptr_t>(rows) % 16 == 0 && reinterpret_cast<uintptr_t>(cols) % 16 == 0 &&
                         reinterpret_cast<uintptr_t>(vals) % 16 == 0;
    const int64_t n_bulk = (bulk_on && aligned) ? nnz / kCooTile : 0;
    const size_t smem_plain = kCooTile * (sizeof(V) + sizeof(I)), smem_bulk = kCooTile * (sizeof(V) + 2 * sizeof(I));
    const int pf = sm_count() * static_cast<int>((227 * 1024) / (smem_bulk + 6 * 1024));
    if (n_bulk > 0 && smem_bulk + 5 * 1024 > 48 * 1024) {   // 64-bit indices: 54 KB of staging
        static const cudaError_t attr = [&] {
            cudaError_t e = cudaFuncSetAttribute(coo_spmv2<V, I, true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                                 static_cast<int>(smem_bulk));
            if (e == cudaSuccess)
                e = cudaFuncSetAttribute(coo_spmv2<V, I, false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         static_cast<int>(smem_bulk));
            return e;
        }();
        if (attr != cudaSuccess) return static_cast<int>(attr);
    }
    for (int64_t j = 0; j < nrhs; ++j) {
#define GKOB200_COO(ADV)                                                                                           \
    if (n_bulk > 0)                                                                                                \
        coo_spmv2<V, I, ADV, true><<<static_cast<unsigned>(n_bulk), kCooThreads, smem_bulk, s>>>(                   \
            nnz, rows, cols, vals, b, b_stride, j, alpha, c, c_stride, carry_row, carry_val, int64_t(0), pf);       \
    if (n_tiles > n_bulk)                                                                                          \
        coo_spmv2<V, I, ADV, false><<<static_cast<unsigned>(n_tiles - n_bulk), kCooThreads, smem_plain, s>>>(       \
            nnz, rows, cols, vals, b, b_stride, j, alpha, c, c_stride, carry_row, carry_val, n_bulk, 0)
        if (alpha) {
            GKOB200_COO(true);
        } else {
            GKOB200_COO(false);
        }
#undef GKOB200_COO
        GKOB200_CHECK_LAUNCH();
        coo_fixup<V><<<static_cast<unsigned>(ceildiv(2 * n_tiles, 256)), 256, 0, s>>>(2 * n_tiles, carry_row, carry_val, c,
                                                                                   c_stride, j);
        GKOB200_CHECK_LAUNCH();
    }
    return 0;
}

// c = beta * c (or 0) ahead of the accumulating COO kernel
template <typename V>
__global__ void scale_or_zero(int64_t n, int64_t k, V* c, int64_t cs, const V* beta)
{
    const int64_t total = n * k;
    const V be = beta ? *beta : V(0);
    for (int64_t t = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; t < total;
         t += static_cast<int64_t>(gridDim.x) * blockDim.x) {
        V& v = c[(t / k) * cs + t % k];
        v = beta ? mul_rn(v, be) : V(0);
    }
}

// ---------------------------------------------------------------------------
// CSR over a list of rows (the non-local block of the distributed matrix):
// c[row] = beta * c[row] + alpha * sum_k val_k b[col_k], one thread per listed row, in
// the reference's advanced-apply order (bit-identical on the listed rows; rows that are
// not listed have no entries, where the reference's  c = 1*c + 0  leaves the bits alone).
// Optional fused partial dot: out[0] = sum_rows w[row] * (alpha * sum).
// ---------------------------------------------------------------------------
template <typename V, typename I>
__global__ void __launch_bounds__(256)
    csr_rows_spmv(int64_t n_listed, const int32_t* __restrict__ row_list, const I* __restrict__ row_ptrs,
                  const I* __restrict__ cols, const V* __restrict__ vals, const V* __restrict__ b, int64_t b_stride,
                  int64_t nrhs, const V* __restrict__ alpha_p, const V* __restrict__ beta_p, V* __restrict__ c,
                  int64_t c_stride, SpmvFusion<V> fu)
{
    if (fu.skip && *fu.skip) return;
    const V alpha = *alpha_p, beta = *beta_p;
    V dot = V(0);
    const int64_t total = n_listed * nrhs;
    for (int64_t t = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; t < total;
         t += static_cast<int64_t>(gridDim.x) * blockDim.x) {
        const int64_t i = t / nrhs, j = t % nrhs;
        const int64_t row = row_list[i];
        V acc = mul_rn(c[row * c_stride + j], beta);
        const V before = acc;
        for (I k = row_ptrs[i]; k < row_ptrs[i + 1]; ++k)
            acc = add_rn(acc, mul_rn(mul_rn(alpha, vals[k]), ldg(b + static_cast<int64_t>(cols[k]) * b_stride + j)));
        c[row * c_stride + j] = acc;
        if (fu.out) dot += fu.w[row] * (acc - before);
    }
    if (fu.out) {
        V tt[1] = {dot};
        V* out = fu.out;
        grid_reduce<1>(tt, ws_partials<V>(fu.ws), ws_ticket(fu.ws), [out](V(&tot)[1]) { out[0] = tot[0]; });
    }
}

// single right-hand side: 8 lanes per listed row.  The lanes fetch 8 entries (col, val, gather)
// in parallel; the products are then added one after the other in storage order (every lane
// runs the same sequential sum over shuffled products), so the result is bit-identical to the
// thread-per-row walk while the row costs ~4 dependent memory latencies instead of ~20 —
// this kernel sits on the critical path of every distributed SpMV.
template <typename V, typename I>
__global__ void __launch_bounds__(256)
    csr_rows_spmv_sub8(int64_t n_listed, const int32_t* __restrict__ row_list, const I* __restrict__ row_ptrs,
                       const I* __restrict__ cols, const V* __restrict__ vals, const V* __restrict__ b,
                       int64_t b_stride, const V* __restrict__ alpha_p, const V* __restrict__ beta_p,
                       V* __restrict__ c, int64_t c_stride, SpmvFusion<V> fu)
{
    if (fu.skip && *fu.skip) return;
    const V alpha = *alpha_p, beta = *beta_p;
    const int sub = threadIdx.x & 7;
    const unsigned group_mask = 0xffu << (threadIdx.x & 24);
    V dot = V(0);
    const int64_t groups = (static_cast<int64_t>(gridDim.x) * blockDim.x) >> 3;
    // all 8 lanes of a group share i, so the loop is uniform per group
    for (int64_t i = (blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x) >> 3; i < n_listed; i += groups) {
        const int64_t row = row_list[i];
        const I begin = row_ptrs[i], end = row_ptrs[i + 1];
        V acc = mul_rn(c[row * c_stride], beta);
        const V before = acc;
        for (I k0 = begin; k0 < end; k0 += 8) {
            const I k = k0 + sub;
            V prod = V(0);
            if (k < end) prod = mul_rn(mul_rn(alpha, vals[k]), ldg(b + static_cast<int64_t>(cols[k]) * b_stride));
            const int cnt = static_cast<int>(min(static_cast<I>(8), end - k0));
#pragma unroll
            for (int u = 0; u < 8; ++u) {
                const V pu = __shfl_sync(group_mask, prod, u, 8);
                if (u < cnt) acc = add_rn(acc, pu);
            }
        }
        if (sub == 0) {
            c[row * c_stride] = acc;
            if (fu.out) dot += fu.w[row] * (acc - before);
        }
    }
    if (fu.out) {
        V tt[1] = {dot};
        V* out = fu.out;
        const P2pDev* p2p = static_cast<const P2pDev*>(fu.p2p);
        V* p2p_buf = fu.p2p_buf;
        const int p2p_count = fu.p2p_count;
        int* on_fail = fu.on_fail;
        grid_reduce<1>(tt, ws_partials<V>(fu.ws), ws_ticket(fu.ws), [=](V(&tot)[1]) {
            out[0] = tot[0];
            // the dot is complete on this rank: all-reduce it right here (one launch for
            // SpMV + dot + all-reduce)
            if (p2p && !peer_allreduce(*p2p, p2p_buf, p2p_count) && on_fail) *on_fail = 1;
        });
    }
}

}  // namespace

template <typename V, typename I>
int csr_rows_spmv_launch(cudaStream_t s, int64_t n_listed, const int32_t* row_list, const I* row_ptrs, const I* cols,
                         const V* vals, const V* b, int64_t b_stride, int64_t nrhs, const V* alpha, const V* beta,
                         V* c, int64_t c_stride, const SpmvFusion<V>* fusion)
{
    if (n_listed < 0 || nrhs < 0 || !alpha || !beta) return GKOB200_EINVAL;
    SpmvFusion<V> fu;
    if (fusion) fu = *fusion;
    if (nrhs != 1) fu.out = nullptr;
    if (n_listed == 0 || nrhs == 0) {
        if (fu.out) GKOB200_CUDA(cudaMemsetAsync(fu.out, 0, sizeof(V), s));
        return 0;
    }
    if (nrhs == 1) {
        const int grid = grid_for(n_listed * 8, 256, 4);
        csr_rows_spmv_sub8<V, I><<<grid, 256, 0, s>>>(n_listed, row_list, row_ptrs, cols, vals, b, b_stride, alpha, beta,
                                                      c, c_stride, fu);
        GKOB200_CHECK_LAUNCH();
        return 0;
    }
    const int grid = grid_for(n_listed * nrhs, 256, 4);
    csr_rows_spmv<V, I><<<grid, 256, 0, s>>>(n_listed, row_list, row_ptrs, cols, vals, b, b_stride, nrhs, alpha, beta, c,
                                             c_stride, fu);
    GKOB200_CHECK_LAUNCH();
    return 0;
}
template int csr_rows_spmv_launch<double, int32_t>(cudaStream_t, int64_t, const int32_t*, const int32_t*, const int32_t*,
                                                   const double*, const double*, int64_t, int64_t, const double*,
                                                   const double*, double*, int64_t, const SpmvFusion<double>*);
template int csr_rows_spmv_launch<float, int32_t>(cudaStream_t, int64_t, const int32_t*, const int32_t*, const int32_t*,
                                                  const float*, const float*, int64_t, int64_t, const float*,
                                                  const float*, float*, int64_t, const SpmvFusion<float>*);

namespace {
// (anonymous namespace reopened for nothing: keeps the layout of this file simple)
}  // namespace

// ---- C++ entry points used by matrix_apply.cu --------------------------------
template <typename V, typename I>
int ell_spmv_launch(cudaStream_t s, int64_t n_rows, int64_t stride, int64_t width, const I* cols, const V* vals,
                    const V* b, int64_t b_stride, int64_t nrhs, const V* alpha, const V* beta, V* c,
                    int64_t c_stride, const SpmvFusion<V>* fusion)
{
    if (n_rows < 0 || width < 0 || stride < n_rows || nrhs < 0 || (alpha == nullptr) != (beta == nullptr))
        return GKOB200_EINVAL;
    if (n_rows == 0 || nrhs == 0) return 0;
    if (!c || (width > 0 && (!cols || !vals || !b))) return GKOB200_EINVAL;
    if (nrhs == 1) {
        const int rc = ell_spmv_tma_launch<V, I>(s, n_rows, stride, width, cols, vals, b, b_stride, alpha, beta, c,
                                                 c_stride, fusion);
        if (rc != 0) return rc == 1 ? 0 : rc;
    }
    return strided_launch<V, I>(s, n_rows, EllFmt{stride, width}, cols, vals, b, b_stride, nrhs, alpha, beta, c,
                                c_stride, fusion);
}

template <typename V, typename I>
int sellp_spmv_launch(cudaStream_t s, int64_t n_rows, int64_t slice_size, const uint64_t* slice_sets,
                      const uint64_t* slice_lengths, const I* cols, const V* vals, const V* b, int64_t b_stride,
                      int64_t nrhs, const V* alpha, const V* beta, V* c, int64_t c_stride,
                      const SpmvFusion<V>* fusion)
{
    if (n_rows < 0 || slice_size <= 0 || nrhs < 0 || (alpha == nullptr) != (beta == nullptr)) return GKOB200_EINVAL;
    if (n_rows == 0 || nrhs == 0) return 0;
    if (!c || !slice_sets || !slice_lengths) return GKOB200_EINVAL;
    return strided_launch<V, I>(s, n_rows, SellpFmt{slice_size, slice_sets, slice_lengths}, cols, vals, b, b_stride,
                                nrhs, alpha, beta, c, c_stride, fusion);
}

template <typename V, typename I>
int coo_spmv_launch(cudaStream_t s, int64_t n_rows, int64_t nnz, const I* rows, const I* cols, const V* vals,
                    const V* b, int64_t b_stride, int64_t nrhs, const V* alpha, const V* beta, bool accumulate,
                    V* c, int64_t c_stride, void* workspace, size_t workspace_bytes)
{
    if (n_rows < 0 || nnz < 0 || nrhs < 0) return GKOB200_EINVAL;
    if (n_rows == 0 || nrhs == 0) return 0;
    if (!c || (nnz > 0 && (!rows || !cols || !vals || !b))) return GKOB200_EINVAL;
    if (!accumulate) {
        // coo::spmv = fill(c, 0) + spmv2; advanced_spmv = scale(beta, c) + advanced_spmv2
        // [reference/matrix/coo_kernels.cpp:62-89]
        scale_or_zero<V><<<grid_for(n_rows * nrhs, 256, 8), 256, 0, s>>>(n_rows, nrhs, c, c_stride, beta);
        GKOB200_CHECK_LAUNCH();
    }
    return coo_spmv2_launch<V, I>(s, n_rows, nnz, rows, cols, vals, b, b_stride, nrhs, alpha, c, c_stride, workspace,
                                  workspace_bytes);
}

#define GKOB200_INST(V, I)                                                                                       \
    template int ell_spmv_launch<V, I>(cudaStream_t, int64_t, int64_t, int64_t, const I*, const V*, const V*,    \
                                       int64_t, int64_t, const V*, const V*, V*, int64_t, const SpmvFusion<V>*); \
    template int sellp_spmv_launch<V, I>(cudaStream_t, int64_t, int64_t, const uint64_t*, const uint64_t*,       \
                                         const I*, const V*, const V*, int64_t, int64_t, const V*, const V*, V*, \
                                         int64_t, const SpmvFusion<V>*);                                         \
    template int coo_spmv_launch<V, I>(cudaStream_t, int64_t, int64_t, const I*, const I*, const V*, const V*,   \
                                       int64_t, int64_t, const V*, const V*, bool, V*, int64_t, void*, size_t);
GKOB200_INST(double, int32_t)
GKOB200_INST(float, int32_t)
GKOB200_INST(double, int64_t)
GKOB200_INST(float, int64_t)
#undef GKOB200_INST

}  // namespace gkob200

using namespace gkob200;

extern "C" {

size_t gkob200_coo_spmv_workspace_bytes(int64_t nnz, int value_bytes)
{
    const int64_t n_tiles = ceildiv(nnz, kCooTile) + 1;
    return static_cast<size_t>(2 * n_tiles) * (sizeof(int64_t) + static_cast<size_t>(value_bytes)) + 64;
}

#define GKOB200_DEF_FMT(V, VT, I, IT)                                                                          \
    int gkob200_ell_spmv_##V##_##I(void* st, int64_t n_rows, int64_t n_cols, int64_t stride, int64_t width,     \
                                   const IT* cols, const VT* vals, const VT* b, int64_t bs, int64_t nrhs,       \
                                   const VT* alpha, const VT* beta, VT* c, int64_t cs)                          \
    {                                                                                                          \
        (void)n_cols;                                                                                          \
        return ell_spmv_launch<VT, IT>(as_stream(st), n_rows, stride, width, cols, vals, b, bs, nrhs, alpha,    \
                                       beta, c, cs, nullptr);                                                  \
    }                                                                                                          \
    int gkob200_sellp_spmv_##V##_##I(void* st, int64_t n_rows, int64_t n_cols, int64_t slice_size,              \
                                     const uint64_t* slice_sets, const uint64_t* slice_lengths, const IT* cols, \
                                     const VT* vals, const VT* b, int64_t bs, int64_t nrhs, const VT* alpha,    \
                                     const VT* beta, VT* c, int64_t cs)                                         \
    {                                                                                                          \
        (void)n_cols;                                                                                          \
        return sellp_spmv_launch<VT, IT>(as_stream(st), n_rows, slice_size, slice_sets, slice_lengths, cols,    \
                                         vals, b, bs, nrhs, alpha, beta, c, cs, nullptr);                       \
    }                                                                                                          \
    int gkob200_coo_spmv_##V##_##I(void* st, int64_t n_rows, int64_t n_cols, int64_t nnz, const IT* rows,       \
                                   const IT* cols, const VT* vals, const VT* b, int64_t bs, int64_t nrhs,       \
                                   const VT* alpha, const VT* beta, VT* c, int64_t cs, void* ws, size_t wsb)    \
    {                                                                                                          \
        (void)n_cols;                                                                                          \
        if ((alpha == nullptr) != (beta == nullptr)) return GKOB200_EINVAL;                                    \
        return coo_spmv_launch<VT, IT>(as_stream(st), n_rows, nnz, rows, cols, vals, b, bs, nrhs, alpha, beta,  \
                                       false, c, cs, ws, wsb);                                                 \
    }                                                                                                          \
    int gkob200_coo_spmv2_##V##_##I(void* st, int64_t n_rows, int64_t n_cols, int64_t nnz, const IT* rows,      \
                                    const IT* cols, const VT* vals, const VT* b, int64_t bs, int64_t nrhs,      \
                                    const VT* alpha, VT* c, int64_t cs, void* ws, size_t wsb)                   \
    {                                                                                                          \
        (void)n_cols;                                                                                          \
        return coo_spmv_launch<VT, IT>(as_stream(st), n_rows, nnz, rows, cols, vals, b, bs, nrhs, alpha,        \
                                       nullptr, true, c, cs, ws, wsb);                                         \
    }
GKOB200_DEF_FMT(f64, double, i32, int32_t)
GKOB200_DEF_FMT(f32, float, i32, int32_t)
GKOB200_DEF_FMT(f64, double, i64, int64_t)
GKOB200_DEF_FMT(f32, float, i64, int64_t)

}  // extern "C"
