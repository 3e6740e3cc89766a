// spmm.cuh — multi-right-hand-side SpMV (SpMM) for CSR, ELL and SELL-P on sm_100a.
//
// Replaces the nrhs > 1 branch of gko::kernels::cuda::{csr,ell,sellp}::spmv/advanced_spmv
// (reference common/cuda_hip/matrix/csr_kernels.hpp.inc abstract_classical_spmv,
// ell_kernels.hpp.inc:42-190, sellp_kernels.hpp.inc:47-134) and follows the arithmetic of
// reference/matrix/{csr,ell,sellp}_kernels.cpp: per (row, column) the products are added in
// storage order with a rounded multiply and a rounded add, so results are bit-identical.
//
// Dense operands are row-major, so a row of b with 32 fp32 columns is exactly one 128-byte
// line and every stored entry needs one full line of b: 58 GB of line requests for the
// 27-pt 256^3 matrix against 7.9 GB of algorithmic DRAM bytes.  The kernel is bound by the
// L2 -> SM fabric (about 7 TB/s measured on B200, profiles/), i.e. by the L1 hit rate of
// those requests, and ncu shows L1 only retains a line over a reuse distance of a few
// hundred lines per SM.  Hence:
//   * lane = right-hand-side column; a warp owns a tile of kTileRows consecutive rows and
//     walks it ENTRY-MAJOR (entry k of all rows, then entry k+1) with one accumulator per
//     row: for banded matrices entry k+1 of row r is the line entry k of row r+1 just
//     fetched, so the reuse distance is kTileRows lines per warp;
//   * the (col,val) pairs of a tile are staged coalesced into shared memory in that
//     entry-major order and read back as 128-bit broadcasts (4 rows per LDS);
//   * a tile whose rows all have the same length and no padding runs without predicates;
//   * address = one IMAD.WIDE (32-bit column x row pitch in bytes + pointer);
//   * when the rows of b and c are 16-byte aligned a lane owns 16 / sizeof(V) columns and the
//     warp's lane groups work on different rows of the tile (entry_step_vec below): 128-bit loads
//     of b, 2.75 instead of 4.5 warp instructions per stored entry;
//   * regular grid operators get a lattice-aware tile order (Lattice below).
#pragma once
#include <cstdlib>

#include "common.cuh"

namespace gkob200 {
namespace spmm {

constexpr int kMaxLen = 36;                       // longest staged row; longer rows take the unstaged path
constexpr int kMaxPasses = 16;                    // consecutive passes (kWarps tiles each) per CTA

// kWarps warps per CTA, kTileRows rows per warp tile, kMinCtas resident CTAs per SM (register
// budget).  Measured on B200 (27-pt 256^3, 32 fp32 RHS, one column per lane): 16 x 8 x 2 = 4.95 ms;
// 4-row tiles 5.8 ms, 16-row tiles 5.8 ms, half the warps 6.8-7.5 ms.  With several columns per
// lane the same configuration runs at 3.51 ms.
template <int W, int T, int M>
struct Cfg {
    static constexpr int kWarps = W, kTileRows = T, kMinCtas = M;
    static constexpr int kTileCap = T * kMaxLen;     // staged entries per warp
};

// Both stagers produce the same layout: entry k of row r of the tile is s_col / s_val
// [k * kTileRows + r] for k < max_len; slots past the end of a row hold col = -1, val = 0.
// stage() returns true when the tile is "clean": no such slot and no padding entry.
//
// CSR: lane (r = lane % kTileRows, q = lane / kTileRows) copies entries q, q + 32/kTileRows,
// ... of row r: 16-byte pieces of kTileRows neighbouring rows per load instruction, the other
// half of each sector is picked up from L1 by the next one; the stores are conflict-free.
template <typename V, typename I, typename C>
struct CsrStager {
    static constexpr int kTileRows = C::kTileRows;
    const I* row_ptrs;
    const I* cols;
    const V* vals;
    int64_t n_rows;
    // per-lane state
    int64_t ptr;    // row_ptrs[row0 + lane] for lane <= nr
    int max_len;

    __device__ __forceinline__ void begin_tile(int64_t row0, int nr, int lane)
    {
        ptr = row_ptrs[row0 + min(lane, nr)];
        max_len = static_cast<int>(min(__shfl_down_sync(0xffffffffu, ptr, 1) - ptr, static_cast<int64_t>(1) << 30));
        if (lane >= nr) max_len = 0;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) max_len = max(max_len, __shfl_xor_sync(0xffffffffu, max_len, o));
    }
    __device__ __forceinline__ bool single_chunk() const { return max_len <= kMaxLen; }
    __device__ __forceinline__ bool stage(I* s_col, V* s_val, int lane) const
    {
        const int r = lane % kTileRows;
        const int64_t first = __shfl_sync(0xffffffffu, ptr, r);
        const int len = static_cast<int>(__shfl_sync(0xffffffffu, ptr, r + 1) - first);
        for (int k = lane / kTileRows; k < max_len; k += 32 / kTileRows) {
            const bool in = k < len;
            s_col[k * kTileRows + r] = in ? cols[first + k] : I(-1);
            s_val[k * kTileRows + r] = in ? vals[first + k] : V(0);
        }
        return __all_sync(0xffffffffu, len == max_len);
    }
    __device__ __forceinline__ void row_entries(int r, int64_t& first, int64_t& step, int64_t& len) const
    {
        first = __shfl_sync(0xffffffffu, ptr, r);
        step = 1;
        len = __shfl_sync(0xffffffffu, ptr, r + 1) - first;
    }
    // column of entry k of `row`, or -1 (lattice detection)
    __device__ __forceinline__ int64_t sample_col(int64_t row, int k) const
    {
        const int64_t first = row_ptrs[row], len = row_ptrs[row + 1] - first;
        return k < len ? static_cast<int64_t>(cols[first + k]) : int64_t(-1);
    }
};

// ELL / SELL-P: a row's entries are `step` apart; lanes (r = lane % kTileRows, kk = lane / kTileRows)
// load kTileRows consecutive rows x 32/kTileRows stored columns per instruction.
template <typename V, typename I, typename Fmt, typename C>
struct StridedStager {
    static constexpr int kTileRows = C::kTileRows;
    Fmt fmt;
    const I* cols;
    const V* vals;
    int64_t n_rows;
    int64_t first, step, len;   // of row (row0 + (lane & 15))
    int64_t max_len;

    __device__ __forceinline__ void begin_tile(int64_t row0, int nr, int lane)
    {
        const int r = lane & (kTileRows - 1);
        first = step = len = 0;
        if (r < nr) fmt.row(row0 + r, first, step, len);
        max_len = len;
#pragma unroll
        for (int o = kTileRows / 2; o > 0; o >>= 1) max_len = max(max_len, __shfl_xor_sync(0xffffffffu, max_len, o));
    }
    __device__ __forceinline__ bool single_chunk() const { return max_len <= kMaxLen; }
    __device__ __forceinline__ bool stage(I* s_col, V* s_val, int lane) const
    {
        const int r = lane & (kTileRows - 1);
        const int n = static_cast<int>(max_len);
        bool pad = false;
        for (int k = lane / kTileRows; k < n; k += 32 / kTileRows) {
            const bool in = k < len;
            const I col = in ? cols[first + k * step] : I(-1);
            s_col[k * kTileRows + r] = col;
            s_val[k * kTileRows + r] = in ? vals[first + k * step] : V(0);
            pad |= col < I(0);
        }
        return !__any_sync(0xffffffffu, pad);
    }
    __device__ __forceinline__ void row_entries(int r, int64_t& first_r, int64_t& step_r, int64_t& len_r) const
    {
        first_r = __shfl_sync(0xffffffffu, first, r);
        step_r = __shfl_sync(0xffffffffu, step, r);
        len_r = __shfl_sync(0xffffffffu, len, r);
    }
    __device__ __forceinline__ int64_t sample_col(int64_t row, int k) const
    {
        int64_t f = 0, st = 0, ln = 0;
        fmt.row(row, f, st, ln);
        return k < ln ? static_cast<int64_t>(cols[f + k * st]) : int64_t(-1);
    }
};

// address of b[col, j]: one IMAD.WIDE.U32 (32-bit column x 32-bit row pitch in bytes + pointer)
template <typename V>
__device__ __forceinline__ const V* b_row(const V* b_j, int32_t col, uint32_t pitch_bytes)
{
    return reinterpret_cast<const V*>(reinterpret_cast<const char*>(b_j) +
                                      static_cast<uint64_t>(static_cast<uint32_t>(col)) * pitch_bytes);
}
template <typename V>
__device__ __forceinline__ const V* b_row(const V* b_j, int64_t col, uint32_t pitch_bytes)
{
    return reinterpret_cast<const V*>(reinterpret_cast<const char*>(b_j) + col * static_cast<int64_t>(pitch_bytes));
}

// kN consecutive shared-memory elements (16-byte aligned) with 128-bit loads
template <int kN, typename T>
__device__ __forceinline__ void lds_vec(const T* p, T (&out)[kN])
{
    constexpr int per = 16 / sizeof(T);
    static_assert(kN % per == 0, "batch must be a whole number of 16-byte vectors");
#pragma unroll
    for (int q = 0; q < kN / per; ++q) {
        union {
            uint4 raw;
            T t[per];
        } u;
        u.raw = reinterpret_cast<const uint4*>(p)[q];
#pragma unroll
        for (int i = 0; i < per; ++i) out[q * per + i] = u.t[i];
    }
}

// One staged entry slot of all kTileRows rows: acc[r] += val[r] * b[col[r], j].
// Checked = false: every slot is a real entry — no predicates.
template <bool Checked, bool Advanced, int kTileRows, typename V, typename I>
__device__ __forceinline__ void entry_step(V (&acc)[kTileRows], const I* s_col, const V* s_val, const V* b_j,
                                           uint32_t b_pitch, V alpha)
{
    I col[kTileRows];
    V v[kTileRows], xv[kTileRows];
    lds_vec(s_col, col);
    lds_vec(s_val, v);
#pragma unroll
    for (int r = 0; r < kTileRows; ++r)
        xv[r] = ldg(b_row(b_j, Checked && col[r] < I(0) ? I(0) : col[r], b_pitch));
#pragma unroll
    for (int r = 0; r < kTileRows; ++r) {
        const V p = Advanced ? mul_rn(mul_rn(alpha, v[r]), xv[r]) : mul_rn(v[r], xv[r]);
        acc[r] = (!Checked || col[r] >= I(0)) ? add_rn(acc[r], p) : acc[r];
    }
}

// ---- several columns per lane ---------------------------------------------------------------
// With kVec = 16 / sizeof(V) columns per lane a row of b with 32 columns is covered by 32 / kVec
// lanes, and the warp's kVec lane groups work on different rows of the tile at the same time:
// per staged slot of the 8-row tile a lane issues kTileRows / kVec 128-bit loads instead of 8
// 32-bit ones, 2 instead of 4 shared-memory loads and kTileRows / kVec instead of 8 address
// computations — 2.75 instead of 4.5 warp instructions per stored entry for fp32, the rounded
// multiplies and adds (2 per entry, bit parity forbids the FMA) being the floor.  Per (row, column)
// the operations and their order are unchanged: results stay bit-identical.
template <typename V>
struct alignas(16) Vec16 {
    V e[16 / sizeof(V)];
};
template <typename V>
__device__ __forceinline__ Vec16<V> ldg16(const V* p)
{
    union {
        uint4 raw;
        Vec16<V> v;
    } u;
    u.raw = __ldg(reinterpret_cast<const uint4*>(p));
    return u.v;
}
// kN consecutive shared-memory elements, kN * sizeof(T) in {8, 16, 32}, aligned to min(that, 16)
template <int kN, typename T>
__device__ __forceinline__ void lds_small(const T* p, T (&out)[kN])
{
    constexpr int bytes = kN * sizeof(T);
    static_assert(bytes == 8 || bytes % 16 == 0, "unsupported group size");
    if constexpr (bytes == 8) {
        union {
            uint2 raw;
            T t[kN];
        } u;
        u.raw = *reinterpret_cast<const uint2*>(p);
#pragma unroll
        for (int i = 0; i < kN; ++i) out[i] = u.t[i];
    } else {
        lds_vec(p, out);
    }
}

template <bool Checked, bool Advanced, int kRows, typename V, typename I>
__device__ __forceinline__ void entry_step_vec(V (&acc)[kRows][16 / sizeof(V)], const I* s_col, const V* s_val,
                                               const V* b_j, uint32_t b_pitch, V alpha)
{
    constexpr int kVec = 16 / sizeof(V);
    I col[kRows];
    V v[kRows];
    Vec16<V> xv[kRows];
    lds_small(s_col, col);
    lds_small(s_val, v);
#pragma unroll
    for (int r = 0; r < kRows; ++r) xv[r] = ldg16(b_row(b_j, Checked && col[r] < I(0) ? I(0) : col[r], b_pitch));
#pragma unroll
    for (int r = 0; r < kRows; ++r) {
        const V av = Advanced ? mul_rn(alpha, v[r]) : v[r];
#pragma unroll
        for (int e = 0; e < kVec; ++e) {
            const V p = mul_rn(av, xv[r].e[e]);
            acc[r][e] = (!Checked || col[r] >= I(0)) ? add_rn(acc[r][e], p) : acc[r][e];
        }
    }
}

// ---- lattice-aware tile order ------------------------------------------------------------
// Every stored entry needs one full row of b; a CTA whose warps own CONSECUTIVE row tiles can
// only reuse fetched rows along the fastest grid direction (at best 9 of the 27 rows a 27-pt
// stencil row touches come from L2: >= 19 GB of L2 -> L1 traffic for the 256^3 matrix against
// 7.9 GB of algorithmic bytes).  When the matrix is a regular grid operator the 16 warps of a
// CTA take row tiles that are neighbours ACROSS grid lines instead — 4 lines x 4 planes (3-D) or
// 16 lines (2-D) at the same position along the line, consecutive passes sliding along it — so
// that their rows share one ~46 KB window of b in L1.  The grid strides are read off the column
// offsets of one row in the middle of the matrix by every CTA (the same row, hence the same
// answer); anything that does not look like a regular grid keeps the consecutive order.  The
// arithmetic per (row, column) is untouched: results stay bit-identical.
struct Lattice {
    int64_t line_tiles;   // tiles per grid line (0: consecutive order)
    int64_t lines;        // lines per plane
    int64_t planes;       // planes (1: two-dimensional)
};

// sample_cols: the columns of row n_rows / 2 (-1 past its end), loaded by kMaxLen threads at once
template <int kTileRows, int kWarps>
__device__ __forceinline__ Lattice detect_lattice(const int64_t* sample_cols, int64_t n_rows, int64_t n_tiles)
{
    static_assert(kWarps == 16, "the lattice patch is 4 x 4 (or 16 x 1) warps");
    Lattice none{0, 0, 1};
    if (n_rows < 4096) return none;
    const int64_t row = n_rows / 2;
    // cluster centres of the positive column offsets of the sample row (clusters: gaps <= 2)
    int64_t centre[16];
    int n_c = 0;
    int64_t lo = -1, hi = -1;
    for (int k = 0; k <= kMaxLen; ++k) {
        const int64_t col = k < kMaxLen ? sample_cols[k] : int64_t(-1);
        const int64_t off = col >= 0 ? col - row : int64_t(-1);
        if (col >= 0 && off <= 0) continue;
        if (lo >= 0 && (col < 0 || off - hi > 2)) {
            if (n_c < 16) centre[n_c++] = (lo + hi) / 2;
            lo = hi = -1;
        }
        if (col < 0) break;
        if (lo < 0) lo = off;
        hi = off;
    }
    int64_t s1 = 0, s2 = 0;
    for (int i = 0; i < n_c && !s1; ++i)
        if (centre[i] >= kTileRows) s1 = centre[i];
    if (!s1 || s1 % kTileRows) return none;
    for (int i = 0; i < n_c && !s2; ++i) {
        const int64_t c = centre[i];
        if (c <= s1 || c % s1) continue;
        bool below = false, above = false, any_neighbour = false;
        for (int j = 0; j < n_c; ++j) {
            below |= centre[j] == c - s1;
            above |= centre[j] == c + s1;
            any_neighbour |= centre[j] > s1 && centre[j] != c && (centre[j] == c - s1 || centre[j] == c + s1);
        }
        // the plane stride is the middle of a (c - s1, c, c + s1) triple, or stands alone (7-pt)
        if ((below && above) || !any_neighbour) s2 = c;
    }
    Lattice L{s1 / kTileRows, 0, 1};
    if (s2) {
        if (n_rows % s2) return none;
        L.lines = s2 / s1;
        L.planes = n_rows / s2;
        if (L.lines % 4 || L.planes % 4) return none;
    } else {
        if (n_rows % s1) return none;
        L.lines = n_rows / s1;
        if (L.lines % 16) return none;
    }
    if (L.line_tiles * L.lines * L.planes != n_tiles) return none;
    return L;
}

// logical (pass, warp) -> row tile.  The passes of a CTA are consecutive, so the position is
// split into (along the line, line block, plane block) once per CTA and then advanced like an
// odometer: 64-bit divisions per pass and warp were a quarter of the instructions outside the
// inner loop.
struct LatticeCursor {
    int64_t xt, yb, zb;   // tile along the line, line block (of 4, or of 16 in 2-D), plane block (of 4)
    __device__ __forceinline__ void seek(const Lattice& L, int64_t pass)
    {
        xt = pass % L.line_tiles;
        const int64_t q = pass / L.line_tiles;
        if (L.planes > 1) {
            yb = q % (L.lines / 4);
            zb = q / (L.lines / 4);
        } else {
            yb = q;
            zb = 0;
        }
    }
    __device__ __forceinline__ void advance(const Lattice& L)
    {
        if (++xt < L.line_tiles) return;
        xt = 0;
        ++yb;
        if (L.planes > 1 && yb == L.lines / 4) {
            yb = 0;
            ++zb;
        }
    }
    __device__ __forceinline__ int64_t tile(const Lattice& L, int wid) const
    {
        if (L.planes > 1) return ((zb * 4 + (wid >> 2)) * L.lines + yb * 4 + (wid & 3)) * L.line_tiles + xt;
        return (yb * 16 + wid) * L.line_tiles + xt;
    }
};

template <typename V, typename I, typename Stager, typename C, bool Advanced, bool Vector>
__global__ void __launch_bounds__(C::kWarps * 32, C::kMinCtas)
    spmm_tiles(int64_t n_rows, Stager st, const V* __restrict__ b, uint32_t b_pitch, int64_t nrhs,
               const V* __restrict__ alpha_p, const V* __restrict__ beta_p, V* __restrict__ c,
               int64_t c_stride, int passes, int lattice_mode)
{
    constexpr int kWarps = C::kWarps, kTileRows = C::kTileRows, kTileCap = C::kTileCap;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    V* s_val = reinterpret_cast<V*>(smem_raw) + wid * kTileCap;
    I* s_col = reinterpret_cast<I*>(reinterpret_cast<V*>(smem_raw) + kWarps * kTileCap) + wid * kTileCap;
    V alpha = V(1), beta = V(0);
    if (Advanced) {
        alpha = *alpha_p;
        beta = *beta_p;
    }
    const int64_t n_tiles = (n_rows + kTileRows - 1) / kTileRows;
    const int64_t cta_tile0 = static_cast<int64_t>(blockIdx.x) * kWarps * passes;
    __shared__ Lattice s_lattice;
    __shared__ int64_t s_sample[kMaxLen];
    if (lattice_mode) {
        if (threadIdx.x < kMaxLen) s_sample[threadIdx.x] = n_rows >= 4096 ? st.sample_col(n_rows / 2, threadIdx.x) : -1;
        __syncthreads();
        if (threadIdx.x == 0) s_lattice = detect_lattice<kTileRows, kWarps>(s_sample, n_rows, n_tiles);
        __syncthreads();
    }
    const Lattice lat = lattice_mode ? s_lattice : Lattice{0, 0, 1};
    LatticeCursor cur{0, 0, 0};
    if (lat.line_tiles > 0) cur.seek(lat, static_cast<int64_t>(blockIdx.x) * passes);
    for (int pass = 0; pass < passes; ++pass) {
        int64_t tile = cta_tile0 + static_cast<int64_t>(pass) * kWarps + wid;
        if (lat.line_tiles > 0) {
            if ((static_cast<int64_t>(blockIdx.x) * passes + pass) * kWarps >= n_tiles) break;
            tile = cur.tile(lat, wid);
            cur.advance(lat);
        }
        if (tile >= n_tiles) break;
        const int64_t row0 = tile * kTileRows;
        const int nr = static_cast<int>(min(static_cast<int64_t>(kTileRows), n_rows - row0));
        st.begin_tile(row0, nr, lane);
        if (st.single_chunk()) {
            // the whole tile fits the staging buffer (the usual case)
            __syncwarp();
            const bool clean = st.stage(s_col, s_val, lane);
            __syncwarp();
            const int n_k = static_cast<int>(st.max_len);
            if constexpr (Vector) {
                constexpr int kVec = 16 / sizeof(V), kLanesPerRow = 32 / kVec, kRows = kTileRows / kVec;
                static_assert(kTileRows % kVec == 0, "tile rows must split over the lane groups");
                const int g = lane / kLanesPerRow;          // lane group: rows g*kRows .. of the tile
                const int jl_lane = (lane % kLanesPerRow) * kVec;
                for (int64_t j0 = 0; j0 < nrhs; j0 += 32) {
                    const bool jl = j0 + jl_lane < nrhs;    // (nrhs is a multiple of kVec)
                    const int64_t j = jl ? j0 + jl_lane : 0;
                    const V* b_j = b + j;
                    V acc[kRows][kVec];
#pragma unroll
                    for (int r = 0; r < kRows; ++r) {
                        const int tr = g * kRows + r;
                        Vec16<V> c0;
                        if (Advanced && tr < nr) c0 = *reinterpret_cast<const Vec16<V>*>(c + (row0 + tr) * c_stride + j);
#pragma unroll
                        for (int e = 0; e < kVec; ++e) acc[r][e] = (Advanced && tr < nr) ? mul_rn(c0.e[e], beta) : V(0);
                    }
                    const I* g_col = s_col + g * kRows;
                    const V* g_val = s_val + g * kRows;
                    if (clean) {
#pragma unroll 2
                        for (int k = 0; k < n_k; ++k)
                            entry_step_vec<false, Advanced>(acc, g_col + k * kTileRows, g_val + k * kTileRows, b_j, b_pitch, alpha);
                    } else {
                        for (int k = 0; k < n_k; ++k)
                            entry_step_vec<true, Advanced>(acc, g_col + k * kTileRows, g_val + k * kTileRows, b_j, b_pitch, alpha);
                    }
#pragma unroll
                    for (int r = 0; r < kRows; ++r) {
                        const int tr = g * kRows + r;
                        if (jl && tr < nr) {
                            Vec16<V> out;
#pragma unroll
                            for (int e = 0; e < kVec; ++e) out.e[e] = acc[r][e];
                            *reinterpret_cast<Vec16<V>*>(c + (row0 + tr) * c_stride + j) = out;
                        }
                    }
                }
                continue;
            }
            for (int64_t j0 = 0; j0 < nrhs; j0 += 32) {
                const bool jl = j0 + lane < nrhs;
                const int64_t j = jl ? j0 + lane : 0;
                const V* b_j = b + j;
                V acc[kTileRows];
#pragma unroll
                for (int r = 0; r < kTileRows; ++r)
                    acc[r] = (Advanced && r < nr) ? mul_rn(c[(row0 + r) * c_stride + j], beta) : V(0);
                if (clean) {
#pragma unroll 2
                    for (int k = 0; k < n_k; ++k)
                        entry_step<false, Advanced>(acc, s_col + k * kTileRows, s_val + k * kTileRows, b_j,
                                                              b_pitch, alpha);
                } else {
                    for (int k = 0; k < n_k; ++k)
                        entry_step<true, Advanced>(acc, s_col + k * kTileRows, s_val + k * kTileRows, b_j,
                                                             b_pitch, alpha);
                }
#pragma unroll
                for (int r = 0; r < kTileRows; ++r)
                    if (jl && r < nr) c[(row0 + r) * c_stride + j] = acc[r];
            }
            continue;
        }
        // long rows: no staging; per row the lanes fetch 32 entries at a time and broadcast them
        for (int64_t j0 = 0; j0 < nrhs; j0 += 32) {
            const bool jl = j0 + lane < nrhs;
            const int64_t j = jl ? j0 + lane : 0;
            const V* b_j = b + j;
            for (int r = 0; r < nr; ++r) {
                int64_t first, step, len;
                st.row_entries(r, first, step, len);
                V* c_rj = c + (row0 + r) * c_stride + j;
                V acc = Advanced ? mul_rn(*c_rj, beta) : V(0);
                for (int64_t k0 = 0; k0 < len; k0 += 32) {
                    const bool in = k0 + lane < len;
                    const I mycol = in ? st.cols[first + (k0 + lane) * step] : I(-1);
                    const V myval = in ? st.vals[first + (k0 + lane) * step] : V(0);
                    const int n = static_cast<int>(min(static_cast<int64_t>(32), len - k0));
                    for (int u = 0; u < n; ++u) {
                        const I col = __shfl_sync(0xffffffffu, mycol, u);
                        const V v = __shfl_sync(0xffffffffu, myval, u);
                        if (col >= I(0)) {
                            const V xv = ldg(b_row(b_j, col, b_pitch));
                            acc = add_rn(acc, Advanced ? mul_rn(mul_rn(alpha, v), xv) : mul_rn(v, xv));
                        }
                    }
                }
                if (jl) *c_rj = acc;
            }
        }
    }
}

// one-time attributes of one kernel instantiation (keyed by the kernel itself)
template <auto Kernel>
cudaError_t prepare(size_t smem, int carveout_pct)
{
    static const cudaError_t rc = [&] {
        cudaError_t e = cudaFuncSetAttribute(Kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem));
        if (e == cudaSuccess) e = cudaFuncSetAttribute(Kernel, cudaFuncAttributePreferredSharedMemoryCarveout, carveout_pct);
        return e;
    }();
    return rc;
}

template <typename V, typename I, typename Stager, typename C>
int launch_cfg(cudaStream_t s, int64_t n_rows, const Stager& st, const V* b, int64_t b_stride, int64_t nrhs,
               const V* alpha, const V* beta, V* c, int64_t c_stride)
{
    constexpr int kWarps = C::kWarps, kTileRows = C::kTileRows, kTileCap = C::kTileCap;
    if (b_stride * sizeof(V) > 0xffffffffull) return GKOB200_EUNSUPPORTED;
    const int64_t n_tiles = ceildiv(n_rows, static_cast<int64_t>(kTileRows));
    // a CTA sweeps `passes` consecutive groups of kWarps tiles; fewer passes on small matrices
    // so that every SM still gets a CTA
    int passes = static_cast<int>(n_tiles / (static_cast<int64_t>(kWarps) * 2 * C::kMinCtas * sm_count()));
    passes = passes < 1 ? 1 : passes > kMaxPasses ? kMaxPasses : passes;
    if (const char* e = getenv("GKOB200_SPMM_PASSES")) passes = atoi(e) > 0 ? atoi(e) : passes;
    const unsigned grid = static_cast<unsigned>(ceildiv(n_tiles, static_cast<int64_t>(kWarps) * passes));
    const size_t smem = static_cast<size_t>(kWarps) * kTileCap * (sizeof(V) + sizeof(I));
    const uint32_t pitch = static_cast<uint32_t>(b_stride * sizeof(V));
    // lattice-aware tile order (GKOB200_SPMM_LATTICE=0 keeps the consecutive order: A/B on the box)
    static const int lattice_mode = [] {
        const char* e = getenv("GKOB200_SPMM_LATTICE");
        return (e && e[0] == '0') ? 0 : 1;
    }();
    // leave the rest of the 256 KB to L1: the kernel lives on L1 hits for b
    const int carve = static_cast<int>((smem + 1024) * C::kMinCtas * 100 / (228 * 1024)) + 1;
    // several columns per lane need 16-byte aligned rows of b and c (GKOB200_SPMM_VECTOR=0: one column
    // per lane, for A/B on the box)
    static const bool vector_ok = [] {
        const char* e = getenv("GKOB200_SPMM_VECTOR");
        return !(e && e[0] == '0');
    }();
    constexpr int kVec = 16 / sizeof(V);
    const bool vec = vector_ok && nrhs % kVec == 0 && reinterpret_cast<uintptr_t>(b) % 16 == 0 &&
                     reinterpret_cast<uintptr_t>(c) % 16 == 0 && b_stride % kVec == 0 && c_stride % kVec == 0;
#define GKOB200_SPMM(ADV, VEC)                                                                             \
    {                                                                                                      \
        const cudaError_t attr = prepare<spmm_tiles<V, I, Stager, C, ADV, VEC>>(smem, carve);              \
        if (attr != cudaSuccess) return static_cast<int>(attr);                                            \
        spmm_tiles<V, I, Stager, C, ADV, VEC><<<grid, kWarps * 32, smem, s>>>(n_rows, st, b, pitch, nrhs, alpha, beta, c, \
                                                                               c_stride, passes, lattice_mode);  \
    }
    if (alpha && vec) GKOB200_SPMM(true, true)
    else if (alpha) GKOB200_SPMM(true, false)
    else if (vec) GKOB200_SPMM(false, true)
    else GKOB200_SPMM(false, false)
#undef GKOB200_SPMM
    GKOB200_CHECK_LAUNCH();
    return 0;
}

using Default = Cfg<16, 8, 2>;

}  // namespace spmm
}  // namespace gkob200
