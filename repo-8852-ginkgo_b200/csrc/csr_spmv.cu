// csr_spmv.cu — CSR SpMV / SpMM for sm_100a.
//
// Replaces gko::kernels::cuda::csr::{spmv, advanced_spmv}
// (reference cuda/matrix/csr_kernels.cu:430-542; device code
// common/cuda_hip/matrix/csr_kernels.hpp.inc:148-483) and follows the arithmetic
// of the oracle reference/matrix/csr_kernels.cpp:75-129.
//
// Two kernels, chosen once per matrix from row-length statistics:
//
//  * row-block ("classical"): a CTA owns 128 consecutive rows.  Their col/val
//    streams are contiguous in memory, so the CTA copies them HBM -> shared memory
//    with fully coalesced loads, then each thread walks ITS row out of shared
//    memory.  Lanes of a warp are consecutive rows, so for banded / stencil
//    matrices the gathers x[col] of a warp fall into 2-3 cache lines (the x
//    access pattern of ELL / SELL-P, obtained on plain CSR).  Each row is summed
//    left to right with a rounded product and a rounded add: the result is
//    bit-identical to the reference executor.
//
//  * merge-path: rows+nnz are split evenly over CTAs (Merrill & Garland's merge
//    path, split rows planned once per matrix), products are formed during the
//    coalesced staging pass, one thread per tile row sums that row's products in
//    storage order behind a single barrier (long rows: the whole warp), and the
//    per-CTA carries of rows that cross tiles are applied by a small
//    deterministic fix-up kernel (no atomics, no zero-fill pass, no allocation —
//    the reference's merge_path allocates two arrays per call and its
//    load_balance kernel needs a fill pass plus fp64 atomics).
//
// Algorithmic bytes per launch: nnz*(V+I) + (n+1)*I + n_cols*k*V + n*k*V.
#include <cstdlib>
#include <cstdio>

#include "internal.h"
#include "spmm.cuh"
#include "tma.cuh"

namespace gkob200 {
namespace {

constexpr int kRowsPerCta = 128;  // == blockDim.x of the row-block kernel

// Shared-memory index with one padding element per 32: keeps the per-thread row
// walks (stride = row length) conflict-free also for even row lengths.
__device__ __forceinline__ int pad(int k) { return k + (k >> 5); }
__host__ __device__ inline int padded_size(int cap) { return cap + (cap >> 5) + 1; }

// ---------------------------------------------------------------------------
// row-block kernel, single right-hand side
// ---------------------------------------------------------------------------
template <typename V, typename I, bool Advanced, bool Fused>
__global__ void __launch_bounds__(kRowsPerCta)
    csr_spmv_rowblock(int64_t n_rows, const I* __restrict__ row_ptrs, const I* __restrict__ col_idxs,
                      const V* __restrict__ values, const V* __restrict__ b, int64_t b_stride,
                      const V* __restrict__ alpha_p, const V* __restrict__ beta_p, V* __restrict__ c,
                      int64_t c_stride, int cap, SpmvFusion<V> fu)
{
    if (Fused && fu.skip && *fu.skip) return;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    V* s_val = reinterpret_cast<V*>(smem_raw);
    I* s_col = reinterpret_cast<I*>(smem_raw + align16(static_cast<size_t>(padded_size(cap)) * sizeof(V)));
    __shared__ I s_ptr[kRowsPerCta + 1];

    const int tid = threadIdx.x;
    const int64_t row0 = static_cast<int64_t>(blockIdx.x) * kRowsPerCta;
    const int nrow = static_cast<int>(min(static_cast<int64_t>(kRowsPerCta), n_rows - row0));

    if (tid < nrow) s_ptr[tid] = row_ptrs[row0 + tid];
    if (tid == 0) s_ptr[nrow] = row_ptrs[row0 + nrow];
    __syncthreads();
    const I tile_begin = s_ptr[0];
    const I tile_end = s_ptr[nrow];
    const I my_begin = tid < nrow ? s_ptr[tid] : tile_end;
    const I my_end = tid < nrow ? s_ptr[tid + 1] : tile_end;

    V alpha = V(1), acc = V(0);
    if (Advanced) {
        alpha = *alpha_p;
        if (tid < nrow) acc = mul_rn(c[(row0 + tid) * c_stride], *beta_p);
    }

    // The tile is processed in chunks of `cap` entries; normally (cap chosen from
    // the matrix statistics) there is exactly one chunk.
    for (I chunk = tile_begin; chunk < tile_end; chunk += cap) {
        const int len = static_cast<int>(min(static_cast<I>(cap), tile_end - chunk));
        if (chunk != tile_begin) __syncthreads();
        // coalesced staging: consecutive threads, consecutive entries
#pragma unroll 4
        for (int k = tid; k < len; k += kRowsPerCta) {
            s_col[pad(k)] = col_idxs[chunk + k];
            s_val[pad(k)] = values[chunk + k];
        }
        __syncthreads();
        // each thread walks the part of its row that lies in this chunk
        const int lo = static_cast<int>(max(my_begin, chunk) - chunk);
        const int hi = static_cast<int>(min(my_end, chunk + static_cast<I>(len)) - chunk);
        for (int k = lo; k < hi; ++k) {
            const V v = s_val[pad(k)];
            const I col = s_col[pad(k)];
            const V xv = ldg(b + static_cast<int64_t>(col) * b_stride);
            acc = Advanced ? add_rn(acc, mul_rn(mul_rn(alpha, v), xv)) : add_rn(acc, mul_rn(v, xv));
        }
    }
    if (tid < nrow) c[(row0 + tid) * c_stride] = acc;
    if (Fused && fu.out) {
        V t[2] = {tid < nrow ? acc * fu.w[row0 + tid] : V(0), tid < nrow ? acc * acc : V(0)};
        V* out = fu.out;
        V* out_sq = fu.out_sq;
        grid_reduce<2>(t, ws_partials<V>(fu.ws), ws_ticket(fu.ws), [out, out_sq](V(&tot)[2]) {
            out[0] = tot[0];
            if (out_sq) out_sq[0] = tot[1];
        });
    }
}

// ---------------------------------------------------------------------------
// row-block kernel, bulk-async staging: the CTA's col/val streams are two
// contiguous byte ranges, so ONE thread hands them to the TMA engine
// (cp.async.bulk) and the whole tile is in flight at once — no register staging,
// no per-thread load loop; the other threads meanwhile fetch the row pointers and
// the (at most 3) tail elements the 16-byte granularity leaves over.
// Requires 16-byte aligned `values` / `col_idxs` base pointers.
// ---------------------------------------------------------------------------
// ---------------------------------------------------------------------------
// Halo exchange inside the SpMV launch (distributed matrix, p2p.cuh).  The first
// H->n_push_ctas CTAs of the grid do not own rows: they copy the entries of b that the
// neighbours need straight into the neighbours' receive windows (peer memory over NVLink),
// and the last of them to finish publishes the epoch flag.  They never wait for this launch of
// a peer — only for a peer to have ENTERED the previous epoch (flow control of the
// double-buffered window) — so every rank's pushes complete no matter how its row CTAs are
// scheduled.
// ---------------------------------------------------------------------------
template <typename V>
__device__ __forceinline__ void halo_push_cta(const HaloDev* __restrict__ H, const V* __restrict__ b, int64_t b_stride,
                                              int* on_fail)
{
    const int tid = threadIdx.x;
    unsigned char* win = H->window;
    const unsigned long long e = *reinterpret_cast<const unsigned long long*>(win + kHaloEpochOff);
    __shared__ int s_fail;
    if (tid == 0) s_fail = 0;
    __syncthreads();
    // (1) every neighbour learns that this rank entered epoch e (so it has finished reading
    //     the receive buffers of all earlier epochs: those reads were earlier launches)
    if (blockIdx.x == 0 && tid < H->n_neighbours) st_relaxed_sys(H->nb_started[tid], e);
    // (2) flow control: buffer e&1 of a receiver still holds epoch e-2 until it entered e-1
    if (tid < H->n_send_peers) {
        const unsigned long long* f =
            reinterpret_cast<const unsigned long long*>(win + kHaloStartedOff) + H->send_peer[tid];
        if (!wait_flag_ge(f, e - 1, H->timeout_ns)) s_fail = 1;
    }
    __syncthreads();
    if (s_fail) {
        // never publish `arrived`: the receivers time out as well and every rank reports the error
        if (tid == 0) {
            *reinterpret_cast<volatile int*>(win + kHaloErrorOff) = 1;
            if (on_fail) *on_fail = 1;
        }
        return;
    }
    // (3) this CTA's share of the send list
    const int n_push = H->n_push_ctas;
    const long long total = H->send_total;
    long long per = (total + n_push - 1) / n_push;
    per = (per + kRowsPerCta - 1) / kRowsPerCta * kRowsPerCta;
    const long long begin = per * blockIdx.x;
    const long long end = begin + per < total ? begin + per : total;
    const int nsp = H->n_send_peers;
    const size_t parity_off = static_cast<size_t>(e & 1);
#pragma unroll 4
    for (long long t = begin + tid; t < end; t += kRowsPerCta) {
        int i = 0;
        while (i + 1 < nsp && t >= H->send_begin[i + 1]) ++i;
        const V v = b[static_cast<int64_t>(H->gather[t]) * b_stride];
        V* dst = reinterpret_cast<V*>(H->dst_data[i] + parity_off * static_cast<size_t>(H->dst_stride[i])) +
                 (t - H->send_begin[i]);
        *dst = v;
    }
    // (4) all entries of all push CTAs before the flags
    __threadfence_system();
    __syncthreads();
    if (tid == 0) {
        unsigned* ticket = reinterpret_cast<unsigned*>(win + kHaloTicketOff);
        if (atomicAdd(ticket, 1u) == static_cast<unsigned>(n_push) - 1u) {
            *ticket = 0u;
            __threadfence_system();
            for (int i = 0; i < nsp; ++i) st_relaxed_sys(H->dst_arrived[i], e);
        }
    }
}

// Resident CTAs per SM the register allocation must allow.  Short-row matrices (5-pt / 7-pt:
// ~10 KB tiles) are limited by registers, not shared memory, and the kernel needs every resident
// tile it can get to cover HBM latency: 12 CTAs (<= 40 registers) instead of the 10 that the
// fused / halo variants got at 44-48 registers cost 20 % of the bandwidth (7-pt); the 5-entry
// batch of the 5-pt variants fits 32 registers = 16 CTAs.
constexpr int rowblock_min_ctas(int batch, size_t) { return batch <= 5 ? 16 : batch <= 7 ? 12 : batch <= 9 ? 9 : 1; }

template <typename V, typename I, bool Advanced, bool Fused, int kBatch, bool Halo = false>
__global__ void __launch_bounds__(kRowsPerCta, rowblock_min_ctas(kBatch, sizeof(V)))
    csr_spmv_rowblock_tma(int64_t n_rows, const I* __restrict__ row_ptrs, const I* __restrict__ col_idxs,
                          const V* __restrict__ values, const V* __restrict__ b, int64_t b_stride,
                          const V* __restrict__ alpha_p, const V* __restrict__ beta_p, V* __restrict__ c,
                          int64_t c_stride, int cap, SpmvFusion<V> fu, int prefetch_tiles, int64_t nnz)
{
    constexpr int VA = 16 / sizeof(V);  // elements per 16 bytes
    constexpr int IA = 16 / sizeof(I);
    extern __shared__ __align__(128) unsigned char smem_raw[];
    V* s_val = reinterpret_cast<V*>(smem_raw);
    I* s_col = reinterpret_cast<I*>(smem_raw + align16(static_cast<size_t>(cap + VA) * sizeof(V)));
    __shared__ I s_ptr[kRowsPerCta + 1];
    __shared__ __align__(8) uint64_t bar;

    const int tid = threadIdx.x;
    int64_t blk = blockIdx.x;
    int nl_slot = -1;   // >= 0: this CTA owns rows with non-local entries (index into nl_slot_begin)
    if (Halo) {
        const int n_push = fu.halo_push_ctas;
        if (static_cast<int>(blockIdx.x) < n_push) {
            if (fu.skip && *fu.skip) return;   // no rank pushes after the stop (the flag is global)
            if (!(fu.halo_debug & 4)) halo_push_cta<V>(fu.halo, b, b_stride, fu.on_fail);
            if (Fused && fu.out && tid < kRowsPerCta / 32) {   // the deferred reduction reads every slot
                constexpr int W = kRowsPerCta / 32;
                ws_partials<V>(fu.ws)[blockIdx.x * W + tid] = V(0);
                if (fu.out_sq) ws_partials<V>(fu.ws)[(gridDim.x + blockIdx.x) * W + tid] = V(0);
            }
            return;
        }
        // Interior row blocks first, blocks with non-local rows last: by the time they run the
        // neighbours' entries have normally arrived and nobody spins.  The interior mapping is
        // arithmetic on kernel parameters: nothing is loaded in front of the bulk copy.
        const int slot = static_cast<int>(blockIdx.x) - n_push;
        if (slot >= fu.halo_n_interior) {
            if (!(fu.halo_debug & 8)) nl_slot = slot - fu.halo_n_interior;
            blk = fu.halo->order[slot];
        } else if (fu.halo_runs > 0) {
            int r = 0;
#pragma unroll
            for (int i = 1; i < kHaloRuns; ++i)
                if (i < fu.halo_runs && slot >= fu.halo_run_slot[i]) r = i;
            blk = fu.halo_run_block[r] + (slot - fu.halo_run_slot[r]);
        } else {
            blk = fu.halo->order[slot];
        }
    }
    const int64_t row0 = blk * kRowsPerCta;
    const int nrow = static_cast<int>(min(static_cast<int64_t>(kRowsPerCta), n_rows - row0));
    if (tid == 0) mbar_init(&bar, 1);
    // Independent loads issued back to back so that their latencies overlap: the solver's
    // "stopped" flag, this thread's two row pointers (coalesced, overlapping by one) and —
    // for the fused dot — w[row], which is only needed in the epilogue.
    int skip = 0;
    if (Fused) {
        if (fu.skip) skip = *fu.skip;
    }
    const I my_begin = row_ptrs[row0 + min(tid, nrow)];
    const I my_end = row_ptrs[row0 + min(tid + 1, nrow)];
    if (tid == 0) s_ptr[0] = my_begin;
    if (tid == nrow - 1) s_ptr[1] = my_end;
    __syncthreads();
    if (Fused && skip) return;  // uniform: every thread read the same flag
    const I tile_begin = s_ptr[0];
    const I tile_end = s_ptr[1];

    V alpha = V(1), acc = V(0);
    if (Advanced) {
        alpha = *alpha_p;
        if (tid < nrow) acc = mul_rn(c[(row0 + tid) * c_stride], *beta_p);
    }
    unsigned parity = 0;
    for (I chunk = tile_begin; chunk < tile_end; chunk += cap) {
        const I chunk_end = min(chunk + static_cast<I>(cap), tile_end);
        // 16-byte aligned windows [vb, vf) / [cb, cf) go through the bulk copy; smem
        // index of global entry g is g - vb (values) / g - cb (columns)
        const I vb = chunk & ~static_cast<I>(VA - 1), vf = chunk_end & ~static_cast<I>(VA - 1);
        const I cb = chunk & ~static_cast<I>(IA - 1), cf = chunk_end & ~static_cast<I>(IA - 1);
        if (chunk != tile_begin) __syncthreads();  // everyone is done with the previous chunk
        if (tid == 0) {
            const unsigned vbytes = vf > vb ? static_cast<unsigned>((vf - vb) * sizeof(V)) : 0u;
            const unsigned cbytes = cf > cb ? static_cast<unsigned>((cf - cb) * sizeof(I)) : 0u;
            mbar_expect_tx(&bar, vbytes + cbytes);
            if (vbytes) bulk_g2s(s_val, values + vb, vbytes, &bar);
            if (cbytes) bulk_g2s(s_col, col_idxs + cb, cbytes, &bar);
            if (prefetch_tiles > 0) {
                // stream-ahead hint: the same-sized window `prefetch_tiles` tiles further on
                const int64_t ahead = static_cast<int64_t>(prefetch_tiles) * (tile_end - tile_begin);
                const int64_t pv = vb + ahead, pc = cb + ahead;
                if (vbytes && pv + (vf - vb) <= nnz) bulk_prefetch_l2(values + pv - (pv & (VA - 1)), vbytes);
                if (cbytes && pc + (cf - cb) <= nnz) bulk_prefetch_l2(col_idxs + pc - (pc & (IA - 1)), cbytes);
            }
        }
        // tails (< 16 bytes each) with plain loads
        {
            const I vt = vf > vb ? vf : vb;
            if (tid < VA && vt + tid < chunk_end) s_val[vt + tid - vb] = values[vt + tid];
            const I ct = cf > cb ? cf : cb;
            if (tid >= 32 && tid < 32 + IA && ct + (tid - 32) < chunk_end)
                s_col[ct + (tid - 32) - cb] = col_idxs[ct + (tid - 32)];
        }
        if (Halo && nl_slot >= 0 && chunk == tile_begin && (tid & 15) == 0) {
            // Boundary CTA: the per-thread entry ranges of the non-local tail are pulled into L2
            // while the bulk copy of the local tile is in flight (no register lives across the
            // row walk: the tail's first load then is an L2 hit)
            asm volatile("prefetch.global.L2 [%0];" ::"l"(fu.halo->nl_thread_range + nl_slot * kRowsPerCta + tid));
        }
        mbar_wait(&bar, parity);
        parity ^= 1u;
        __syncthreads();  // tails visible
        const I lo = max(my_begin, chunk), hi = min(my_end, chunk_end);
        const V* sv = s_val - vb;  // index with the global entry number
        const I* sc = s_col - cb;
        // batches of kBatch entries: all gathers of a batch are issued before the first
        // add, so a thread keeps kBatch L2/L1 requests in flight; the adds stay in
        // storage order (bit-identical to the oracle)
        for (I k = lo; k < hi; k += kBatch) {
            V v[kBatch], xv[kBatch];
#pragma unroll
            for (int u = 0; u < kBatch; ++u) {
                const bool in = k + u < hi;
                v[u] = in ? sv[k + u] : V(0);
                const I col = in ? sc[k + u] : I(0);
                xv[u] = in ? ldg(b + static_cast<int64_t>(col) * b_stride) : V(0);
            }
#pragma unroll
            for (int u = 0; u < kBatch; ++u) {
                if (k + u < hi)
                    acc = Advanced ? add_rn(acc, mul_rn(mul_rn(alpha, v[u]), xv[u]))
                                   : add_rn(acc, mul_rn(v[u], xv[u]));
            }
        }
    }
    if (Halo && nl_slot >= 0) {
        // Non-local entries of this block's rows: c = 1*c + A_nl * ghost continues the row sums
        // in storage order (reference: non_local_mtx_->apply(one, recv, one, x) after the local
        // apply, core/distributed/matrix.cpp:312-333; advanced apply: alpha in front) — the
        // result is bit-identical to two separate applies.
        const HaloDev* __restrict__ H = fu.halo;
        unsigned char* win = H->window;
        __shared__ int s_wait_fail;
        if (tid == 0) s_wait_fail = 0;
        const int2 nl_range = H->nl_thread_range[nl_slot * kRowsPerCta + tid];
        const int nl_k0 = nl_range.x, nl_k1 = nl_range.y;
        const unsigned long long halo_epoch = *reinterpret_cast<const unsigned long long*>(win + kHaloEpochOff);
        __syncthreads();
        if (tid < H->n_recv_peers) {
            // (acquire even when the peek already saw the flag: the entries are read next)
            const unsigned long long* f =
                reinterpret_cast<const unsigned long long*>(win + kHaloArrivedOff) + H->recv_peer[tid];
            if (!wait_flag_ge(f, halo_epoch, H->timeout_ns)) {
                s_wait_fail = 1;
                *reinterpret_cast<volatile int*>(win + kHaloErrorOff) = 1;
                if (fu.on_fail) *fu.on_fail = 1;
            }
        }
        __syncthreads();
        if (nl_k1 > nl_k0 && !s_wait_fail) {
            // (L2 loads: the window is written by the peers, never through this SM's L1)
            const V* recv = reinterpret_cast<const V*>(win + kHaloDataOff) + (halo_epoch & 1) * H->recv_stride;
            const V* nl_vals = static_cast<const V*>(H->nl_vals);
            for (int k = nl_k0; k < nl_k1; ++k) {
                const V v = Advanced ? mul_rn(alpha, nl_vals[k]) : nl_vals[k];
                acc = add_rn(acc, mul_rn(v, __ldcg(recv + H->nl_cols[k])));
            }
        }
    }
    if (tid < nrow) c[(row0 + tid) * c_stride] = acc;
    if (Fused && fu.out) {
        // w[row] is fetched here, not in the prologue: in the solvers w is the vector this CTA just
        // gathered from (the diagonal entry), so the load hits L1, and two registers less live
        // across the row walk keep the short-row variants at 12 resident CTAs without spills
        const V w_row = tid < nrow ? ldg(fu.w + row0 + tid) : V(0);
        // one partial per warp (no barrier), summed by finish_partials right after this launch
        if (fu.out_sq)
            store_warp_partial2(tid < nrow ? acc * w_row : V(0), tid < nrow ? acc * acc : V(0), ws_partials<V>(fu.ws),
                                kRowsPerCta / 32);
        else
            store_warp_partial(tid < nrow ? acc * w_row : V(0), ws_partials<V>(fu.ws), kRowsPerCta / 32);
    }
}

// ---------------------------------------------------------------------------
// merge-path kernel (single right-hand side)
// ---------------------------------------------------------------------------
constexpr int kMpThreads = 128;
constexpr int kMpItems = 9;                            // merge items per thread (odd: conflict-free)
constexpr int kMpTile = kMpThreads * kMpItems;         // merge items per CTA

// Merge-path diagonal search: how many of the first `diag` merge items are row
// ends.  List A = row end offsets row_ptrs[1..n], list B = 0..nnz-1; a row end is
// consumed when row_end[r] <= current nnz index.
template <typename I, typename P>
__device__ __forceinline__ int64_t merge_path_search(int64_t diag, const P row_end, int64_t n_rows, int64_t nnz)
{
    int64_t lo = diag > nnz ? diag - nnz : 0;
    int64_t hi = diag < n_rows ? diag : n_rows;
    while (lo < hi) {
        const int64_t mid = (lo + hi) >> 1;
        if (static_cast<int64_t>(row_end[mid]) <= diag - mid - 1) {
            lo = mid + 1;
        } else {
            hi = mid;
        }
    }
    return lo;
}

// One tile = kMpTile consecutive merge items (rows + entries), cut by the merge path: tiles are
// equal amounts of work whatever the row lengths.  Inside a tile:
//   1. every thread issues its kMpItems (col, val) loads, the loads of the row boundaries it will
//      stage, then its kMpItems gathers — all independent, all in flight together (the gathers of
//      a skewed matrix are random 32-byte sectors: what has to be hidden is latency);
//   2. products and tile-relative row starts go to shared memory; ONE barrier;
//   3. one thread per tile row walks that row's products front to back (storage order: a row that
//      lies inside one tile gets exactly the reference's sequential sum) and stores the result;
//      a row with more than kMpLongRow entries in the tile is summed by its whole warp instead
//      (lane-strided partial sums + shuffle tree — a fixed association, deterministic);
//   4. the row left open at the end of the tile leaves a carry for the fix-up kernel.
// With a plan (tile split rows computed once per matrix) there is no search and no second
// barrier.  History (profiles/r02_merge_source_phases.txt): the previous kernel flagged row
// heads per entry, reduced kMpItems consecutive products per thread and stitched the pieces in
// thread order — 8 barriers per tile, 37 % of all warp time in the stitch and the barrier
// behind it; 833 us on the 10 M-row power-law matrix.
// Measured at 10 M rows (tools/c3_sweep.py): 128 threads x 8 resident CTAs (58 registers) 659 us;
// 10 CTAs (48 registers) 770 us, 12 CTAs (40 registers, spills) 1141 us; 256-thread tiles 4 % slower;
// look-ahead L2 prefetch + streaming stores of c: 685 -> 676 us (8 % at 4-7 M rows); the
// lane-per-tile fix-up: 676 -> 662 us.  A kernel that ONLY streams (col, val) and gathers b[col]
// (tools/probe/gather_probe.cu) takes 465 us on this matrix: that, not the algorithmic-byte
// roofline, is the bound a CSR SpMV can approach here.
constexpr int kMpLongRow = 48;

template <typename V, typename I, bool Advanced>
__global__ void __launch_bounds__(kMpThreads, 8)
    csr_spmv_merge(int64_t n_rows, int64_t nnz, const I* __restrict__ row_ptrs, const I* __restrict__ col_idxs,
                   const V* __restrict__ values, const V* __restrict__ b, int64_t b_stride,
                   const V* __restrict__ alpha_p, const V* __restrict__ beta_p, V* __restrict__ c,
                   int64_t c_stride, int64_t* __restrict__ carry_row, V* __restrict__ carry_val,
                   const int64_t* __restrict__ plan, int prefetch_tiles)
{
    __shared__ __align__(16) V s_prod[kMpTile];
    __shared__ int s_start[kMpTile + 2];   // [r]: first product of tile row r; [n_tile_rows + 1] = n_tile_nnz
    __shared__ int64_t s_range[2];

    const int tid = threadIdx.x;
    const int64_t total = n_rows + nnz;
    const int64_t diag0 = min(static_cast<int64_t>(blockIdx.x) * kMpTile, total);
    const int64_t diag1 = min(diag0 + kMpTile, total);
    const I* row_end = row_ptrs + 1;
    int64_t r_begin, r_end, pf_row = -1;
    if (plan) {
        // same two words for the whole CTA: one broadcast request, no barrier
        r_begin = plan[blockIdx.x];
        r_end = plan[blockIdx.x + 1];
        if (tid == 0 && prefetch_tiles > 0 && blockIdx.x + prefetch_tiles < gridDim.x) pf_row = plan[blockIdx.x + prefetch_tiles];
    } else {
        if (tid < 2) s_range[tid] = merge_path_search<I>(tid == 0 ? diag0 : diag1, row_end, n_rows, nnz);
        __syncthreads();
        r_begin = s_range[0];
        r_end = s_range[1];
    }
    const int64_t k_begin = diag0 - r_begin;
    const int n_tile_rows = static_cast<int>(r_end - r_begin);                 // rows that END in this tile
    const int n_tile_nnz = static_cast<int>((diag1 - r_end) - k_begin);
    const I* tile_cols = col_idxs + k_begin;
    const V* tile_vals = values + k_begin;
    const I* tile_row_end = row_end + r_begin - 1;   // [r]: end of tile row r-1 = start of tile row r (r >= 1)
    V alpha = V(1);
    if (Advanced) alpha = *alpha_p;
    {
        V v[kMpItems], xv[kMpItems];
        I col[kMpItems];
#pragma unroll
        for (int u = 0; u < kMpItems; ++u) {
            const int k = tid + u * kMpThreads;
            const bool in = k < n_tile_nnz;
            col[u] = in ? __ldcs(tile_cols + k) : I(0);
            v[u] = in ? __ldcs(tile_vals + k) : V(0);
        }
        // row boundaries of the first kMpThreads tile rows: requested before the gathers queue up
        I first_end = I(0);
        if (tid >= 1 && tid <= n_tile_rows) first_end = __ldcs(tile_row_end + tid);
#pragma unroll
        for (int u = 0; u < kMpItems; ++u) xv[u] = __ldg(b + static_cast<int64_t>(col[u]) * b_stride);
        if (pf_row >= 0) {
            // L2 prefetch of the (col, val) ranges of the tile one resident wave ahead: its CTA then
            // starts its gathers an L2 hit, not a DRAM access, after it was launched
            const int64_t kp = (min((static_cast<int64_t>(blockIdx.x) + prefetch_tiles) * kMpTile, total) - pf_row) &
                               ~static_cast<int64_t>(3);
            if (kp + kMpTile <= nnz) {
                bulk_prefetch_l2(values + kp, kMpTile * sizeof(V));
                bulk_prefetch_l2(col_idxs + kp, kMpTile * sizeof(I));
            }
        }
        // tile row r (r == n_tile_rows: the row left open at the end of the tile) starts where row
        // r-1 ends; tile row 0 is open when the tile starts
        s_start[tid] = tid == 0 ? 0
                                : (tid <= n_tile_rows ? static_cast<int>(static_cast<int64_t>(first_end) - k_begin) : n_tile_nnz);
        for (int r = tid + kMpThreads; r <= n_tile_rows + 1; r += kMpThreads)
            s_start[r] = r <= n_tile_rows ? static_cast<int>(static_cast<int64_t>(__ldcs(tile_row_end + r)) - k_begin) : n_tile_nnz;
#pragma unroll
        for (int u = 0; u < kMpItems; ++u) {
            const int k = tid + u * kMpThreads;
            if (k < n_tile_nnz) s_prod[k] = Advanced ? mul_rn(mul_rn(alpha, v[u]), xv[u]) : mul_rn(v[u], xv[u]);
        }
    }
    __syncthreads();

    const int lane = tid & 31;
    for (int base = 0; base <= n_tile_rows; base += kMpThreads) {   // warp-uniform trip count
        const int r = base + tid;
        const bool valid = r <= n_tile_rows;
        const int start = valid ? s_start[r] : 0, end = valid ? s_start[r + 1] : 0;
        const bool is_long = end - start > kMpLongRow;
        V acc = V(0);
        if (!is_long) {
            int k = start;
            for (; k + 4 <= end; k += 4) {
                const V a0 = s_prod[k], a1 = s_prod[k + 1], a2 = s_prod[k + 2], a3 = s_prod[k + 3];
                acc = add_rn(add_rn(add_rn(add_rn(acc, a0), a1), a2), a3);
            }
            for (; k < end; ++k) acc = add_rn(acc, s_prod[k]);
        }
        for (unsigned m = __ballot_sync(0xffffffffu, is_long); m; m &= m - 1) {
            const int src = __ffs(m) - 1;
            const int ls = __shfl_sync(0xffffffffu, start, src), le = __shfl_sync(0xffffffffu, end, src);
            V part = V(0);
            for (int k = ls + lane; k < le; k += 32) part = add_rn(part, s_prod[k]);
            part = __shfl_sync(0xffffffffu, warp_sum(part), 0);   // (the tree leaves the total in lane 0)
            if (lane == src) acc = part;
        }
        if (!valid) continue;
        if (r < n_tile_rows) {
            const int64_t row = r_begin + r;
            // beta*c is applied exactly once, by the tile in which the row ends
            // (streaming store: the result must not push the gathered vector out of L2)
            __stcs(c + row * c_stride, Advanced ? add_rn(mul_rn(c[row * c_stride], *beta_p), acc) : acc);
        } else {
            // the row continues into the next tile: per-CTA carry (-1: no entry of it in this tile)
            carry_row[blockIdx.x] = end > start ? r_begin + r : int64_t(-1);
            carry_val[blockIdx.x] = acc;
        }
    }
}

// split row of every tile diagonal, computed once per matrix
template <typename I>
__global__ void csr_merge_plan(int64_t n_tiles, int64_t n_rows, int64_t nnz, const I* __restrict__ row_ptrs,
                               int64_t* __restrict__ plan, int tile)
{
    const int64_t t = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x;
    if (t > n_tiles) return;
    const int64_t total = n_rows + nnz;
    const int64_t d = min(t * tile, total);
    plan[t] = merge_path_search<I>(d, row_ptrs + 1, n_rows, nnz);
}

// Fix-up: tile carries belonging to the same row are consecutive.  One LANE per tile, a warp
// owns a window of 32 tiles: the carries of the window arrive in three coalesced, independent
// loads, runs of equal rows are added by a segmented shuffle scan, and the lane at the start of
// a run applies the sum to c.  A run that leaves the window (a row of 10^5 entries spans ~90
// tiles) is continued by the whole warp, 32 tiles per step.  Fixed association for a given tile
// layout: deterministic.  (One warp per tile, the previous scheme, was a chain of 4-5 dependent
// memory latencies for each of ~10^5 warps: 24 us at 10 M rows.)
template <typename V>
__global__ void __launch_bounds__(256) csr_spmv_merge_fixup(int n_tiles, const int64_t* __restrict__ carry_row,
                                                            const V* __restrict__ carry_val, int64_t n_rows,
                                                            V* __restrict__ c, int64_t c_stride)
{
    constexpr unsigned kFull = 0xffffffffu;
    const int lane = threadIdx.x & 31;
    const int base = ((blockIdx.x * blockDim.x + threadIdx.x) >> 5) * 32;
    if (base >= n_tiles) return;   // warp-uniform
    const int t = base + lane;
    int64_t row = t < n_tiles ? carry_row[t] : int64_t(-1);
    int64_t before = lane == 0 && base > 0 ? carry_row[base - 1] : int64_t(-1);
    V sum = t < n_tiles ? carry_val[t] : V(0);
    int64_t beyond = base + 32 + lane < n_tiles ? carry_row[base + 32 + lane] : int64_t(-1);
    if (row >= n_rows) row = -1;
    const int64_t up = __shfl_up_sync(kFull, row, 1);
    const bool run_start = row >= 0 && (lane == 0 ? before : up) != row;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const V vo = __shfl_down_sync(kFull, sum, o);
        const int64_t ro = __shfl_down_sync(kFull, row, o);
        if (lane + o < 32 && row >= 0 && ro == row) sum = add_rn(sum, vo);
    }
    const int64_t last = __shfl_sync(kFull, row, 31);
    if (last >= 0) {   // warp-uniform: the window's last run may go on in the tiles behind it
        V extra = V(0);
        for (int u0 = base + 32;; u0 += 32) {
            const int u = u0 + lane;
            const bool same = u < n_tiles && (u0 == base + 32 ? beyond : carry_row[u]) == last;
            const unsigned m = __ballot_sync(kFull, same);
            const int n_same = m == kFull ? 32 : __ffs(~m) - 1;
            if (lane < n_same) extra = add_rn(extra, carry_val[u]);
            if (n_same < 32) break;
        }
        extra = __shfl_sync(kFull, warp_sum(extra), 0);
        if (run_start && row == last) sum = add_rn(sum, extra);
    }
    if (run_start) c[row * c_stride] = add_rn(sum, c[row * c_stride]);
}

// ---------------------------------------------------------------------------
// row statistics
// ---------------------------------------------------------------------------
template <typename I>
__global__ void csr_row_stats(int64_t n_rows, const I* __restrict__ row_ptrs, unsigned long long* stats)
{
    unsigned long long max_row = 0, max_blk = 0, empty = 0;
    for (int64_t r = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; r < n_rows;
         r += static_cast<int64_t>(gridDim.x) * blockDim.x) {
        const unsigned long long len = static_cast<unsigned long long>(row_ptrs[r + 1] - row_ptrs[r]);
        max_row = max(max_row, len);
        empty += (len == 0);
        if (r % kRowsPerCta == 0) {
            const int64_t e = min(r + kRowsPerCta, n_rows);
            max_blk = max(max_blk, static_cast<unsigned long long>(row_ptrs[e] - row_ptrs[r]));
        }
    }
    // integer max/sum are order independent: atomics keep this exact
    for (int o = 16; o > 0; o >>= 1) {
        max_row = max(max_row, __shfl_down_sync(0xffffffffu, max_row, o));
        max_blk = max(max_blk, __shfl_down_sync(0xffffffffu, max_blk, o));
        empty += __shfl_down_sync(0xffffffffu, empty, o);
    }
    if ((threadIdx.x & 31) == 0) {
        atomicMax(stats + 0, max_row);
        atomicMax(stats + 1, max_blk);
        atomicAdd(stats + 2, empty);
    }
    if (blockIdx.x == 0 && threadIdx.x == 0) stats[3] = static_cast<unsigned long long>(row_ptrs[n_rows]);
}

template <typename I>
int row_stats_impl(void* stream, int64_t n_rows, const I* row_ptrs, int64_t* stats)
{
    if (n_rows < 0 || !stats) return GKOB200_EINVAL;
    cudaStream_t s = as_stream(stream);
    GKOB200_CUDA(cudaMemsetAsync(stats, 0, 4 * sizeof(int64_t), s));
    if (n_rows == 0) return 0;
    if (!row_ptrs) return GKOB200_EINVAL;
    csr_row_stats<I><<<grid_for(n_rows, 256, 8), 256, 0, s>>>(n_rows, row_ptrs,
                                                             reinterpret_cast<unsigned long long*>(stats));
    GKOB200_CHECK_LAUNCH();
    return 0;
}

// 1: bulk-async (TMA) staging [default], 0: register-staged loads.  The environment
// variable only exists to A/B the two on the GPU box (profiles/).
inline int rowblock_variant()
{
    static int v = -1;
    if (v < 0) {
        const char* e = getenv("GKOB200_CSR_ROWBLOCK");
        v = (e && e[0] == 'p') ? 0 : 1;
    }
    return v;
}

// tuning knobs of the row-block kernel (environment overrides exist only to A/B on the box)
inline int env_int(const char* name, int dflt)
{
    const char* e = getenv(name);
    return e ? atoi(e) : dflt;
}
inline int rowblock_batch()
{
    static int v = env_int("GKOB200_CSR_BATCH", 0);  // 0: from the mean row length
    return v;
}
inline int rowblock_prefetch()
{
    static int v = env_int("GKOB200_CSR_PREFETCH", -1);  // -1: one resident wave
    return v;
}

inline int rowblock_cap(int64_t max_block_nnz, size_t elem_bytes)
{
    // shared-memory budget per CTA for the staged chunk: stays under the 48 KB
    // static limit (>= 4 CTAs per SM); heavier blocks are processed in chunks
    const int64_t budget = 46 * 1024 / static_cast<int64_t>(elem_bytes);
    int64_t cap = max_block_nnz > 0 ? max_block_nnz : 2048;
    cap = (cap + 255) / 256 * 256;
    if (cap < 256) cap = 256;
    if (cap > budget) cap = budget / 256 * 256;
    return static_cast<int>(cap);
}

}  // namespace

template <typename V, typename I>
int csr_spmv_launch(cudaStream_t s, int64_t n_rows, int64_t n_cols, int64_t nnz, const I* row_ptrs,
                    const I* col_idxs, const V* values, const V* b, int64_t b_stride, int64_t nrhs,
                    const V* alpha, const V* beta, V* c, int64_t c_stride, int strategy,
                    int64_t max_block_nnz, void* workspace, size_t workspace_bytes,
                    const SpmvFusion<V>* fusion)
{
    if (n_rows < 0 || n_cols < 0 || nnz < 0 || nrhs < 0) return GKOB200_EINVAL;
    if ((alpha == nullptr) != (beta == nullptr)) return GKOB200_EINVAL;
    if (n_rows == 0 || nrhs == 0) return 0;
    if (!row_ptrs || !c || (nnz > 0 && (!col_idxs || !values || !b))) return GKOB200_EINVAL;
    if (b_stride < nrhs || c_stride < nrhs) return GKOB200_EINVAL;
    const bool adv = alpha != nullptr;
    SpmvFusion<V> fu;
    if (fusion) fu = *fusion;
    const bool fused = fusion != nullptr;

    if (strategy == GKOB200_CSR_AUTO) {
        strategy = max_block_nnz > 0
                       ? gkob200_csr_pick_strategy(n_rows, nnz, -1, max_block_nnz)
                       : GKOB200_CSR_MERGE_PATH;
    }
    if (nrhs > 1) {
        if (fused) return GKOB200_EUNSUPPORTED;
        // SpMM: warp-tile kernel, lanes over right-hand sides (spmm.cuh)
        spmm::CsrStager<V, I, spmm::Default> st{row_ptrs, col_idxs, values, n_rows, 0, 0};
        return spmm::launch_cfg<V, I, decltype(st), spmm::Default>(s, n_rows, st, b, b_stride, nrhs, alpha, beta, c,
                                                                   c_stride);
    }
    if (strategy == GKOB200_CSR_CLASSICAL) {
        const int cap = rowblock_cap(max_block_nnz, sizeof(V) + sizeof(I));
        const size_t smem = align16(static_cast<size_t>(padded_size(cap)) * sizeof(V)) +
                            align16(static_cast<size_t>(padded_size(cap)) * sizeof(I)) + 64;
        const bool halo = fused && fu.halo != nullptr;
        const unsigned grid = static_cast<unsigned>(ceildiv(n_rows, kRowsPerCta)) + (halo ? fu.halo_push_ctas : 0);
        const bool aligned = (reinterpret_cast<uintptr_t>(values) % 16 == 0) &&
                             (reinterpret_cast<uintptr_t>(col_idxs) % 16 == 0);
        // the halo exchange only exists in the bulk-async kernel with 32-bit indices
        if (halo && !(aligned && rowblock_variant() == 1 && sizeof(I) == 4)) return GKOB200_EUNSUPPORTED;
        if (aligned && rowblock_variant() == 1) {
            // the deferred reduction leaves one partial per CTA (and array) in fu.ws
            constexpr int kWarps = kRowsPerCta / 32;
            if (fused && fu.out &&
                static_cast<int64_t>(grid) * kWarps * (fu.out_sq ? 2 : 1) + 2 * kFinishMaxCtas >
                    (fu.ws_blocks + 256) * kReduceMaxVals)
                return GKOB200_EWORKSPACE;
            // stream-ahead distance: one wave of resident CTAs (shared-memory or thread bound)
            int resident = static_cast<int>((227 * 1024) / (smem + 1024));
            if (resident > 16) resident = 16;
            if (resident < 1) resident = 1;
            int pf = rowblock_prefetch();
            if (pf < 0) pf = sm_count() * resident;
            // gathers kept in flight per thread ~ the mean row length
            int batch = rowblock_batch();
            if (batch <= 0) {
                const double mean = static_cast<double>(nnz) / static_cast<double>(n_rows);
                batch = mean <= 5.5 ? 5 : mean <= 7.5 ? 7 : mean <= 10.5 ? 9 : 14;
            }
#define GKOB200_RBT(ADV, FUSED, BATCH, HALO)                                                          \
    csr_spmv_rowblock_tma<V, I, ADV, FUSED, BATCH, HALO><<<grid, kRowsPerCta, smem, s>>>(             \
        n_rows, row_ptrs, col_idxs, values, b, b_stride, alpha, beta, c, c_stride, cap, fu, pf, nnz)
#define GKOB200_RBT_B(BATCH)                                                        \
    if (halo) {                                                                     \
        if constexpr (sizeof(I) == 4) {                                             \
            if (adv) GKOB200_RBT(true, true, BATCH, true);                          \
            else GKOB200_RBT(false, true, BATCH, true);                             \
        }                                                                           \
    }                                                                               \
    else if (adv && fused) GKOB200_RBT(true, true, BATCH, false);                   \
    else if (adv) GKOB200_RBT(true, false, BATCH, false);                           \
    else if (fused) GKOB200_RBT(false, true, BATCH, false);                         \
    else GKOB200_RBT(false, false, BATCH, false)
            if (batch >= 14) { GKOB200_RBT_B(14); }
            else if (batch >= 9) { GKOB200_RBT_B(9); }
            else if (batch >= 7) { GKOB200_RBT_B(7); }
            else { GKOB200_RBT_B(5); }
#undef GKOB200_RBT_B
#undef GKOB200_RBT
            GKOB200_CHECK_LAUNCH();
            if (fused && fu.out) return launch_finish_partials<V>(s, static_cast<int64_t>(grid) * kWarps, fu);
            return 0;
        }
        if (fused && fu.out && static_cast<int64_t>(grid) * (fu.out_sq ? 2 : 1) > fu.ws_blocks) return GKOB200_EWORKSPACE;
#define GKOB200_RB(ADV, FUSED)                                                                        \
    csr_spmv_rowblock<V, I, ADV, FUSED><<<grid, kRowsPerCta, smem, s>>>(                              \
        n_rows, row_ptrs, col_idxs, values, b, b_stride, alpha, beta, c, c_stride, cap, fu)
        if (adv && fused) GKOB200_RB(true, true);
        else if (adv) GKOB200_RB(true, false);
        else if (fused) GKOB200_RB(false, true);
        else GKOB200_RB(false, false);
#undef GKOB200_RB
        GKOB200_CHECK_LAUNCH();
        return 0;
    }
    if (strategy != GKOB200_CSR_MERGE_PATH && strategy != GKOB200_CSR_MERGE_PATH_PLANNED) return GKOB200_EINVAL;
    if (fused && (fu.out || fu.skip)) return GKOB200_EUNSUPPORTED;
    const int64_t n_tiles = ceildiv(n_rows + nnz, kMpTile);
    const size_t need = gkob200_csr_spmv_workspace_bytes(n_rows, nnz, nrhs, sizeof(V));
    if (!workspace || workspace_bytes < need) return GKOB200_EWORKSPACE;
    int64_t* carry_row = reinterpret_cast<int64_t*>(workspace);
    const int64_t* plan = strategy == GKOB200_CSR_MERGE_PATH_PLANNED ? carry_row + n_tiles + 1 : nullptr;
    V* carry_val = reinterpret_cast<V*>(carry_row + 2 * (n_tiles + 1) + 1);
    // tiles of look-ahead for the L2 prefetch (needs 16-byte aligned arrays); GKOB200_MP_PREFETCH=0: off
    static const int mp_prefetch = [] {
        const char* e = getenv("GKOB200_MP_PREFETCH");
        return e ? atoi(e) : sm_count() * 4;
    }();
    const int pf = (reinterpret_cast<uintptr_t>(values) % 16 == 0 && reinterpret_cast<uintptr_t>(col_idxs) % 16 == 0) ? mp_prefetch : 0;
#define GKOB200_MP(ADV)                                                                                    \
    csr_spmv_merge<V, I, ADV><<<static_cast<unsigned>(n_tiles), kMpThreads, 0, s>>>(                       \
        n_rows, nnz, row_ptrs, col_idxs, values, b, b_stride, alpha, beta, c, c_stride, carry_row, carry_val, plan, pf)
    if (adv) GKOB200_MP(true);
    else GKOB200_MP(false);
#undef GKOB200_MP
    GKOB200_CHECK_LAUNCH();
    csr_spmv_merge_fixup<V><<<static_cast<unsigned>(ceildiv(n_tiles, 256)), 256, 0, s>>>(
        static_cast<int>(n_tiles), carry_row, carry_val, n_rows, c, c_stride);
    GKOB200_CHECK_LAUNCH();
    return 0;
}

#define GKOB200_INST(V, I)                                                                              \
    template int csr_spmv_launch<V, I>(cudaStream_t, int64_t, int64_t, int64_t, const I*, const I*,     \
                                       const V*, const V*, int64_t, int64_t, const V*, const V*, V*,    \
                                       int64_t, int, int64_t, void*, size_t, const SpmvFusion<V>*);
GKOB200_INST(double, int32_t)
GKOB200_INST(float, int32_t)
GKOB200_INST(double, int64_t)
GKOB200_INST(float, int64_t)
#undef GKOB200_INST

}  // namespace gkob200

using namespace gkob200;

template <typename I>
static int merge_plan_impl(void* stream, int64_t n_rows, int64_t nnz, const I* row_ptrs, void* workspace,
                           size_t workspace_bytes)
{
    if (n_rows < 0 || nnz < 0) return GKOB200_EINVAL;
    if (n_rows == 0) return 0;
    if (!row_ptrs) return GKOB200_EINVAL;
    if (!workspace || workspace_bytes < gkob200_csr_spmv_workspace_bytes(n_rows, nnz, 1, 8)) return GKOB200_EWORKSPACE;
    const int64_t n_tiles = ceildiv(n_rows + nnz, kMpTile);
    int64_t* plan = reinterpret_cast<int64_t*>(workspace) + n_tiles + 1;
    csr_merge_plan<I><<<static_cast<unsigned>(ceildiv(n_tiles + 1, 256)), 256, 0, as_stream(stream)>>>(n_tiles, n_rows, nnz,
                                                                                                   row_ptrs, plan, kMpTile);
    GKOB200_CHECK_LAUNCH();
    return 0;
}
extern "C" {

size_t gkob200_csr_spmv_workspace_bytes(int64_t n_rows, int64_t nnz, int64_t nrhs, int value_bytes)
{
    (void)nrhs;
    // [carry_row: n_tiles+1][plan (tile split rows): n_tiles+2][carry_val: n_tiles+1]
    const int64_t n_tiles = ceildiv(n_rows + nnz, kMpTile) + 2;
    return static_cast<size_t>(n_tiles) * (2 * sizeof(int64_t) + static_cast<size_t>(value_bytes)) + 64;
}

int gkob200_csr_pick_strategy(int64_t n_rows, int64_t nnz, int64_t max_row_nnz, int64_t max_block_nnz)
{
    // The row-block kernel is the right choice while a block of 128 rows is close
    // to the average block (regular matrices: stencils, FEM, banded).  A block much
    // heavier than average means skewed rows: thread-per-row would serialise on the
    // long rows, so the nnz-balanced merge-path kernel takes over.
    if (n_rows <= 0 || nnz <= 0) return GKOB200_CSR_CLASSICAL;
    const double mean_block = static_cast<double>(nnz) / static_cast<double>(ceildiv(n_rows, kRowsPerCta));
    const double mean_row = static_cast<double>(nnz) / static_cast<double>(n_rows);
    if (max_block_nnz > 0 && static_cast<double>(max_block_nnz) > 4.0 * mean_block + 1024.0)
        return GKOB200_CSR_MERGE_PATH;
    if (max_row_nnz > 0 && static_cast<double>(max_row_nnz) > 8.0 * mean_row + 256.0)
        return GKOB200_CSR_MERGE_PATH;
    return GKOB200_CSR_CLASSICAL;
}

int gkob200_csr_merge_plan_i32(void* stream, int64_t n_rows, int64_t nnz, const int32_t* row_ptrs, void* ws, size_t wsb)
{
    return merge_plan_impl<int32_t>(stream, n_rows, nnz, row_ptrs, ws, wsb);
}
int gkob200_csr_merge_plan_i64(void* stream, int64_t n_rows, int64_t nnz, const int64_t* row_ptrs, void* ws, size_t wsb)
{
    return merge_plan_impl<int64_t>(stream, n_rows, nnz, row_ptrs, ws, wsb);
}

int gkob200_csr_row_stats_i32(void* stream, int64_t n_rows, const int32_t* row_ptrs, int64_t* stats)
{
    return row_stats_impl<int32_t>(stream, n_rows, row_ptrs, stats);
}
int gkob200_csr_row_stats_i64(void* stream, int64_t n_rows, const int64_t* row_ptrs, int64_t* stats)
{
    return row_stats_impl<int64_t>(stream, n_rows, row_ptrs, stats);
}

#define GKOB200_DEF_CSR_SPMV(V, VT, I, IT)                                                            \
    int gkob200_csr_spmv_##V##_##I(void* stream, int64_t n_rows, int64_t n_cols, int64_t nnz,         \
                                   const IT* row_ptrs, const IT* col_idxs, const VT* values,          \
                                   const VT* b, int64_t b_stride, int64_t nrhs, const VT* alpha,      \
                                   const VT* beta, VT* c, int64_t c_stride, int strategy,             \
                                   int64_t max_block_nnz, void* workspace, size_t workspace_bytes)    \
    {                                                                                                 \
        return csr_spmv_launch<VT, IT>(as_stream(stream), n_rows, n_cols, nnz, row_ptrs, col_idxs,    \
                                       values, b, b_stride, nrhs, alpha, beta, c, c_stride, strategy, \
                                       max_block_nnz, workspace, workspace_bytes, nullptr);           \
    }
GKOB200_DEF_CSR_SPMV(f64, double, i32, int32_t)
GKOB200_DEF_CSR_SPMV(f32, float, i32, int32_t)
GKOB200_DEF_CSR_SPMV(f64, double, i64, int64_t)
GKOB200_DEF_CSR_SPMV(f32, float, i64, int64_t)

}  // extern "C"
