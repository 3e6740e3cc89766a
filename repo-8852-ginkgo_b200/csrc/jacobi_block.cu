// jacobi_block.cu — block-Jacobi: block detection, generation (explicit inverses) and
// application, for sm_100a.
//
// Replaces gko::kernels::cuda::jacobi::{find_blocks, generate, simple_apply, apply}
// (reference cuda/preconditioner/jacobi_*.cu, common/cuda_hip/preconditioner/*.hpp.inc)
// and follows the oracle reference/preconditioner/jacobi_kernels.cpp:66-562.
//
//  * find_blocks — the reference runs two sequential scans (natural blocks = runs of rows
//    with identical sparsity pattern capped at max_block_size; greedy agglomeration),
//    its CUDA back-end as two <<<1,1>>> kernels.  Here both are data-parallel and produce
//    the SAME block pointers: the cap is applied per run with a scan (row i starts a block
//    iff (i - run_start) % max == 0) and the greedy agglomeration chain 0 -> next(0) ->
//    next(next(0)) ... is marked by pointer doubling in log2(n) rounds.
//  * generate — one warp per block, one lane per row, block in shared memory; Gauss-Jordan
//    with the reference's max-abs column pivoting and its exact operation order (rounded
//    product + rounded sum), so the stored inverses are bit-identical to the oracle's.
//  * apply — one warp per block: the lane's row of the inverse is read once (coalesced in
//    the interleaved storage), b is broadcast by shuffles, the inner sum runs in the
//    reference's order (bit-identical); right-hand sides are looped INSIDE the kernel (the
//    reference launches one kernel per column: cuda/preconditioner/jacobi_simple_apply_kernel.cu:77-89).
//
// Algorithmic bytes of apply: num_blocks * stride * max_block_size * V  (=32 n V for
// max_block_size 32) + 2 n k V.
#include "launch.cuh"

namespace gkob200 {
namespace {

// ------------------------------ find_blocks -----------------------------------
__global__ void run_start_flags(int64_t n, const int32_t* __restrict__ rp, const int32_t* __restrict__ ci,
                                int32_t* __restrict__ f)
{
    const int64_t i = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x;
    if (i > n) return;
    if (i == n) {
        f[i] = 0;
        return;
    }
    bool same = false;
    if (i > 0) {
        const int32_t pb = rp[i - 1], cb = rp[i], ce = rp[i + 1];
        same = (ce - cb) == (cb - pb);
        for (int32_t k = 0; same && k < ce - cb; ++k) same = ci[pb + k] == ci[cb + k];
    }
    f[i] = same ? 0 : 1;
}

// E = exclusive scan of f.  starts[run] = i for run-start rows.
__global__ void scatter_run_starts(int64_t n, const int32_t* __restrict__ f_excl, const int32_t* __restrict__ rp,
                                   const int32_t* __restrict__ ci, int32_t* __restrict__ starts)
{
    const int64_t i = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x;
    if (i >= n) return;
    // is row i a run start?  E[i+1] - E[i] == f[i]
    if (f_excl[i + 1] - f_excl[i] == 1) starts[f_excl[i]] = static_cast<int32_t>(i);
}

__global__ void natural_block_flags(int64_t n, const int32_t* __restrict__ f_excl, const int32_t* __restrict__ starts,
                                    int32_t max_bs, int32_t* __restrict__ g)
{
    const int64_t i = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x;
    if (i > n) return;
    if (i == n) {
        g[i] = 0;
        return;
    }
    const int32_t fi = f_excl[i + 1] - f_excl[i];
    const int32_t run = f_excl[i] + fi - 1;
    g[i] = ((static_cast<int32_t>(i) - starts[run]) % max_bs == 0) ? 1 : 0;
}

// compact: ptrs[excl[i]] = i where flag (excl[i+1]-excl[i]) is set; ptrs[total] = n; *count = total
__global__ void compact_starts(int64_t n, const int32_t* __restrict__ excl, int32_t* __restrict__ ptrs,
                               int32_t* __restrict__ count)
{
    const int64_t i = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x;
    if (i > n) return;
    if (i == n) {
        ptrs[excl[n]] = static_cast<int32_t>(n);
        *count = excl[n];
        return;
    }
    if (excl[i + 1] - excl[i] == 1) ptrs[excl[i]] = static_cast<int32_t>(i);
}

// next(a): first natural block j > a with ptrs[j+1] - ptrs[a] > max  (or num_nat)
__global__ void agglomerate_next(int64_t cap, const int32_t* __restrict__ nat_ptrs, const int32_t* __restrict__ num_nat_p,
                                 int32_t max_bs, int32_t* __restrict__ next, int32_t* __restrict__ mark)
{
    const int64_t a = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x;
    if (a > cap) return;
    const int32_t num_nat = *num_nat_p;
    if (a >= num_nat) {
        next[a] = static_cast<int32_t>(a);  // terminal / unused
        mark[a] = 0;
        return;
    }
    int32_t j = static_cast<int32_t>(a) + 1;
    while (j < num_nat && nat_ptrs[j + 1] - nat_ptrs[a] <= max_bs) ++j;
    next[a] = j;
    mark[a] = a == 0 ? 1 : 0;
}

// round: every marked node marks J(node); then J <- J o J
__global__ void mark_round(int64_t cap, const int32_t* __restrict__ num_nat_p, const int32_t* __restrict__ J,
                           const int32_t* __restrict__ mark_in, int32_t* __restrict__ mark_out,
                           int32_t* __restrict__ J_out)
{
    const int64_t a = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x;
    if (a > cap) return;
    const int32_t num_nat = *num_nat_p;
    if (a >= num_nat) {
        J_out[a] = static_cast<int32_t>(a);
        return;
    }
    const int32_t j = J[a];
    J_out[a] = j < num_nat ? J[j] : j;
    if (mark_in[a] && j < num_nat) mark_out[j] = 1;  // benign race: all writers store 1
}

__global__ void compact_blocks(int64_t cap, const int32_t* __restrict__ num_nat_p, const int32_t* __restrict__ excl,
                               const int32_t* __restrict__ nat_ptrs, int32_t* __restrict__ block_ptrs,
                               int64_t* __restrict__ num_blocks, int64_t n)
{
    const int64_t a = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x;
    if (a > cap) return;
    const int32_t num_nat = *num_nat_p;
    if (a == 0) {
        const int32_t nb = num_nat > 0 ? excl[num_nat] : 0;
        *num_blocks = nb;
        block_ptrs[nb] = static_cast<int32_t>(n);
        if (num_nat == 0) block_ptrs[0] = 0;
    }
    if (a < num_nat && excl[a + 1] - excl[a] == 1) block_ptrs[excl[a]] = nat_ptrs[a];
}

// ------------------------------- generate -------------------------------------
constexpr int kMaxBs = 32;
constexpr int kWarpsPerCta = 4;

template <typename V>
__device__ __forceinline__ V vabs(V v) { return v < V(0) ? -v : v; }

template <typename V>
__global__ void __launch_bounds__(32 * kWarpsPerCta)
    block_generate(int64_t num_blocks, const int32_t* __restrict__ rp, const int32_t* __restrict__ ci,
                   const V* __restrict__ va, const int32_t* __restrict__ bptrs, int64_t block_offset,
                   int64_t group_offset, int group_power, V* __restrict__ blocks)
{
    __shared__ V s_blk[kWarpsPerCta][kMaxBs][kMaxBs + 1];
    __shared__ int s_perm[kWarpsPerCta][kMaxBs];
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    const int64_t blk = static_cast<int64_t>(blockIdx.x) * kWarpsPerCta + w;
    if (blk >= num_blocks) return;
    const int32_t start = bptrs[blk];
    const int bs = bptrs[blk + 1] - start;
    V(*B)[kMaxBs + 1] = s_blk[w];
    int* perm = s_perm[w];
    // extract_block: row `lane` of the diagonal block
    if (lane < bs) {
        for (int j = 0; j < bs; ++j) B[lane][j] = V(0);
        for (int32_t k = rp[start + lane]; k < rp[start + lane + 1]; ++k) {
            const int32_t c = ci[k] - start;
            if (c >= 0 && c < bs) B[lane][c] = va[k];
        }
        perm[lane] = lane;
    }
    __syncwarp();
    // invert_block: Gauss-Jordan, pivot = first max |.| in column k among rows >= k
    bool ok = true;
    for (int k = 0; k < bs && ok; ++k) {
        // choose_pivot (strict '<' keeps the first maximum)
        V best = (lane >= k && lane < bs) ? vabs(B[lane][k]) : V(-1);
        int bi = lane;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            const V ov = __shfl_down_sync(0xffffffffu, best, o);
            const int oi = __shfl_down_sync(0xffffffffu, bi, o);
            if (ov > best || (ov == best && oi < bi)) {
                best = ov;
                bi = oi;
            }
        }
        const int cp = __shfl_sync(0xffffffffu, bi, 0);
        // swap_rows(k, cp) + swap(perm[k], perm[cp])
        if (cp != k) {
            if (lane < bs) {
                const V t = B[k][lane];
                B[k][lane] = B[cp][lane];
                B[cp][lane] = t;
            }
            if (lane == 0) {
                const int t = perm[k];
                perm[k] = perm[cp];
                perm[cp] = t;
            }
        }
        __syncwarp();
        // apply_gauss_jordan_transform(k, k)
        const V d = B[k][k];
        if (d == V(0)) {
            ok = false;
            break;
        }
        __syncwarp();
        if (lane < bs) B[lane][k] = div_rn(B[lane][k], -d);
        __syncwarp();
        if (lane == 0) B[k][k] = V(0);
        __syncwarp();
        if (lane < bs) {
            const V f = B[lane][k];
            for (int j = 0; j < bs; ++j) {
                // row k itself: f == 0 there, the reference performs the same (no-op) update
                B[lane][j] = add_rn(B[lane][j], mul_rn(f, B[k][j]));
            }
        }
        __syncwarp();
        if (lane < bs) B[k][lane] = div_rn(B[k][lane], d);
        __syncwarp();
        if (lane == 0) B[k][k] = div_rn(V(1), d);
        __syncwarp();
    }
    // permute_and_transpose_block: result[i + perm[j]*stride] = B[i][j]
    const int64_t gs_mask = (int64_t(1) << group_power) - 1;
    const int64_t stride = block_offset << group_power;
    V* out = blocks + group_offset * (blk >> group_power) + block_offset * (blk & gs_mask);
    if (lane < bs)
        for (int j = 0; j < bs; ++j) out[lane + static_cast<int64_t>(perm[j]) * stride] = B[lane][j];
}

// --------------------------------- apply ---------------------------------------
template <typename V, bool Advanced>
__global__ void __launch_bounds__(32 * kWarpsPerCta)
    block_apply(int64_t num_blocks, const int32_t* __restrict__ bptrs, const V* __restrict__ blocks,
                int64_t block_offset, int64_t group_offset, int group_power, int64_t k, const V* __restrict__ alpha_p,
                const V* __restrict__ b, int64_t bs_, const V* __restrict__ beta_p, V* __restrict__ x, int64_t xs)
{
    const int lane = threadIdx.x & 31;
    const int64_t blk = (blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x) >> 5;
    if (blk >= num_blocks) return;
    const int32_t start = bptrs[blk];
    const int bs = bptrs[blk + 1] - start;
    const int64_t gs_mask = (int64_t(1) << group_power) - 1;
    const int64_t stride = block_offset << group_power;
    const V* blkp = blocks + group_offset * (blk >> group_power) + block_offset * (blk & gs_mask);
    V inv[kMaxBs];
#pragma unroll
    for (int c = 0; c < kMaxBs; ++c) inv[c] = (lane < bs && c < bs) ? blkp[lane + c * stride] : V(0);
    V alpha = V(1), beta = V(0);
    if (Advanced) {
        alpha = *alpha_p;
        beta = *beta_p;
#pragma unroll
        for (int c = 0; c < kMaxBs; ++c) inv[c] = mul_rn(alpha, inv[c]);
    }
    for (int64_t col = 0; col < k; ++col) {
        const V bv = lane < bs ? b[(start + lane) * bs_ + col] : V(0);
        V acc = V(0);
        if (Advanced && lane < bs && beta != V(0)) acc = mul_rn(x[(start + lane) * xs + col], beta);
#pragma unroll
        for (int c = 0; c < kMaxBs; ++c) {
            const V bc = __shfl_sync(0xffffffffu, bv, c);
            if (c < bs) acc = add_rn(acc, mul_rn(inv[c], bc));
        }
        if (lane < bs) x[(start + lane) * xs + col] = acc;
    }
}

}  // namespace
}  // namespace gkob200

using namespace gkob200;

// out block = transpose of the stored (inverse) block, same storage scheme
// [ref: jacobi::transpose_jacobi / conj_transpose_jacobi, reference/preconditioner/jacobi_kernels.cpp:629-690]
template <typename V>
__global__ void __launch_bounds__(32 * kWarpsPerCta)
    block_transpose(int64_t num_blocks, const int32_t* __restrict__ bptrs, const V* __restrict__ blocks,
                    int64_t block_offset, int64_t group_offset, int group_power, V* __restrict__ out)
{
    const int64_t blk = static_cast<int64_t>(blockIdx.x) * kWarpsPerCta + (threadIdx.x >> 5);
    if (blk >= num_blocks) return;
    const int lane = threadIdx.x & 31;
    const int bs = bptrs[blk + 1] - bptrs[blk];
    const int64_t gs_mask = (int64_t(1) << group_power) - 1;
    const int64_t stride = block_offset << group_power;
    const int64_t base = group_offset * (blk >> group_power) + block_offset * (blk & gs_mask);
    for (int c = 0; c < bs; ++c)
        if (lane < bs) out[base + c + lane * stride] = blocks[base + lane + c * stride];
}

extern "C" {

size_t gkob200_jacobi_find_blocks_workspace_bytes(int64_t n_rows)
{
    return static_cast<size_t>(7 * (n_rows + 4)) * sizeof(int32_t) + gkob200_prefix_sum_workspace_bytes(n_rows + 2) + 256;
}

int gkob200_jacobi_find_blocks_i32(void* stream, int64_t n, const int32_t* row_ptrs, const int32_t* col_idxs,
                                   int32_t max_block_size, int64_t* num_blocks, int32_t* block_pointers, void* ws,
                                   size_t ws_bytes)
{
    if (n < 0 || max_block_size < 1 || max_block_size > kMaxBs || !num_blocks || !block_pointers) return GKOB200_EINVAL;
    cudaStream_t s = as_stream(stream);
    if (n == 0) {
        GKOB200_CUDA(cudaMemsetAsync(num_blocks, 0, sizeof(int64_t), s));
        GKOB200_CUDA(cudaMemsetAsync(block_pointers, 0, sizeof(int32_t), s));
        return 0;
    }
    if (!ws || ws_bytes < gkob200_jacobi_find_blocks_workspace_bytes(n)) return GKOB200_EWORKSPACE;
    const int64_t len = n + 4;
    int32_t* f = reinterpret_cast<int32_t*>(ws);  // run-start flags -> exclusive scan
    int32_t* starts = f + len;
    int32_t* g = starts + len;  // natural-block flags -> exclusive scan
    int32_t* nat_ptrs = g + len;
    int32_t* J0 = nat_ptrs + len;
    int32_t* J1 = J0 + len;
    int32_t* mark = J1 + len;
    void* scan_ws = mark + len;
    const size_t scan_wsb = ws_bytes - static_cast<size_t>(7 * len) * sizeof(int32_t);
    int32_t* num_nat = starts + n + 2;  // spare slot
    const unsigned grid = static_cast<unsigned>(ceildiv(n + 1, 256));
    int rc;
    run_start_flags<<<grid, 256, 0, s>>>(n, row_ptrs, col_idxs, f);
    if ((rc = gkob200_prefix_sum_i32(stream, f, n + 1, scan_ws, scan_wsb))) return rc;
    scatter_run_starts<<<grid, 256, 0, s>>>(n, f, row_ptrs, col_idxs, starts);
    natural_block_flags<<<grid, 256, 0, s>>>(n, f, starts, max_block_size, g);
    if ((rc = gkob200_prefix_sum_i32(stream, g, n + 1, scan_ws, scan_wsb))) return rc;
    compact_starts<<<grid, 256, 0, s>>>(n, g, nat_ptrs, num_nat);
    // agglomeration: chain of next() from natural block 0, marked by pointer doubling
    agglomerate_next<<<grid, 256, 0, s>>>(n, nat_ptrs, num_nat, max_block_size, J0, mark);
    int rounds = 1;
    while ((int64_t(1) << rounds) < n + 1) ++rounds;
    int32_t *Jin = J0, *Jout = J1;
    for (int r = 0; r <= rounds; ++r) {
        mark_round<<<grid, 256, 0, s>>>(n, num_nat, Jin, mark, mark, Jout);
        int32_t* t = Jin;
        Jin = Jout;
        Jout = t;
    }
    // (mark_round reads and writes `mark` in place: a node marked early in a round only makes
    //  the doubling progress faster; the fixed point — every node of the chain marked — is the same)
    GKOB200_CUDA(cudaMemsetAsync(mark + n, 0, 4 * sizeof(int32_t), s));
    if ((rc = gkob200_prefix_sum_i32(stream, mark, n + 1, scan_ws, scan_wsb))) return rc;
    compact_blocks<<<grid, 256, 0, s>>>(n, num_nat, mark, nat_ptrs, block_pointers, num_blocks, n);
    GKOB200_CHECK_LAUNCH();
    return 0;
}

#define GKOB200_DEF_JB(V, VT)                                                                                    \
    int gkob200_jacobi_block_generate_##V(void* stream, int64_t n_rows, const int32_t* row_ptrs,                  \
                                          const int32_t* col_idxs, const VT* values, int64_t num_blocks,          \
                                          const int32_t* block_pointers, int64_t block_offset,                    \
                                          int64_t group_offset, int group_power, VT* blocks)                      \
    {                                                                                                            \
        (void)n_rows;                                                                                            \
        if (num_blocks < 0) return GKOB200_EINVAL;                                                               \
        if (num_blocks == 0) return 0;                                                                           \
        block_generate<VT><<<static_cast<unsigned>(ceildiv(num_blocks, kWarpsPerCta)), 32 * kWarpsPerCta, 0,      \
                             as_stream(stream)>>>(num_blocks, row_ptrs, col_idxs, values, block_pointers,         \
                                                  block_offset, group_offset, group_power, blocks);               \
        GKOB200_CHECK_LAUNCH();                                                                                  \
        return 0;                                                                                                \
    }                                                                                                            \
    int gkob200_jacobi_block_transpose_##V(void* stream, int64_t num_blocks, const int32_t* block_pointers,       \
                                           const VT* blocks, int64_t block_offset, int64_t group_offset,          \
                                           int group_power, VT* out_blocks)                                      \
    {                                                                                                            \
        if (num_blocks < 0) return GKOB200_EINVAL;                                                               \
        if (num_blocks == 0) return 0;                                                                           \
        block_transpose<VT><<<static_cast<unsigned>(ceildiv(num_blocks, kWarpsPerCta)), 32 * kWarpsPerCta, 0,     \
                              as_stream(stream)>>>(num_blocks, block_pointers, blocks, block_offset, group_offset, \
                                                   group_power, out_blocks);                                     \
        GKOB200_CHECK_LAUNCH();                                                                                  \
        return 0;                                                                                                \
    }                                                                                                            \
    int gkob200_jacobi_block_simple_apply_##V(void* stream, int64_t num_blocks, const int32_t* block_pointers,    \
                                              const VT* blocks, int64_t block_offset, int64_t group_offset,       \
                                              int group_power, int64_t n, int64_t k, const VT* b, int64_t bs,     \
                                              VT* x, int64_t xs)                                                  \
    {                                                                                                            \
        (void)n;                                                                                                 \
        if (num_blocks < 0 || k < 0) return GKOB200_EINVAL;                                                      \
        if (num_blocks == 0 || k == 0) return 0;                                                                 \
        block_apply<VT, false><<<static_cast<unsigned>(ceildiv(num_blocks, kWarpsPerCta)), 32 * kWarpsPerCta, 0,  \
                                 as_stream(stream)>>>(num_blocks, block_pointers, blocks, block_offset,           \
                                                      group_offset, group_power, k, nullptr, b, bs, nullptr, x,   \
                                                      xs);                                                        \
        GKOB200_CHECK_LAUNCH();                                                                                  \
        return 0;                                                                                                \
    }                                                                                                            \
    int gkob200_jacobi_block_apply_##V(void* stream, int64_t num_blocks, const int32_t* block_pointers,           \
                                       const VT* blocks, int64_t block_offset, int64_t group_offset,              \
                                       int group_power, int64_t n, int64_t k, const VT* alpha, const VT* b,       \
                                       int64_t bs, const VT* beta, VT* x, int64_t xs)                             \
    {                                                                                                            \
        (void)n;                                                                                                 \
        if (num_blocks < 0 || k < 0 || !alpha || !beta) return GKOB200_EINVAL;                                   \
        if (num_blocks == 0 || k == 0) return 0;                                                                 \
        block_apply<VT, true><<<static_cast<unsigned>(ceildiv(num_blocks, kWarpsPerCta)), 32 * kWarpsPerCta, 0,   \
                                as_stream(stream)>>>(num_blocks, block_pointers, blocks, block_offset,            \
                                                     group_offset, group_power, k, alpha, b, bs, beta, x, xs);    \
        GKOB200_CHECK_LAUNCH();                                                                                  \
        return 0;                                                                                                \
    }
GKOB200_DEF_JB(f64, double)
GKOB200_DEF_JB(f32, float)

}  // extern "C"
