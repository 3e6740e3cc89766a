// generators.cu — deterministic synthetic matrices of the BASELINE shapes, built on
// the HOST straight into caller buffers (numpy / pinned memory), rows
// [row_begin,row_end) of the global matrix so that each rank of a row-partitioned
// run only ever materialises its own slab.  The same arrays feed the oracle and
// the GPU path.  (The reference builds its stencils on the host the same way:
// examples/distributed-solver/distributed-solver.cpp:176-186 for the 1-D case.)
//
//  kind 0: 2D 5-pt  Laplacian on nx*ny          diag 4,  off -1
//  kind 1: 3D 7-pt  Laplacian on nx*ny*nz       diag 6,  off -1
//  kind 2: 3D 27-pt stencil   on nx*ny*nz       diag 26, off -1
// Dirichlet truncation: neighbours outside the grid are dropped.  Row index
// = x + nx*(y + ny*z); entries of a row are sorted by column.
#include <omp.h>

#include <algorithm>
#include <cmath>
#include <vector>

#include "common.cuh"

namespace gkob200 {
namespace {

struct Stencil {
    int kind;
    int64_t nx, ny, nz;
    int64_t n() const { return nx * ny * (kind == 0 ? 1 : nz); }
    // number of entries of row i
    template <typename F>
    inline void for_each(int64_t row, F f) const
    {
        const int64_t x = row % nx, y = (row / nx) % ny, z = row / (nx * ny);
        if (kind == 2) {
            for (int dz = -1; dz <= 1; ++dz) {
                if (z + dz < 0 || z + dz >= nz) continue;
                for (int dy = -1; dy <= 1; ++dy) {
                    if (y + dy < 0 || y + dy >= ny) continue;
                    for (int dx = -1; dx <= 1; ++dx) {
                        if (x + dx < 0 || x + dx >= nx) continue;
                        const bool diag = dx == 0 && dy == 0 && dz == 0;
                        f(row + dx + nx * (dy + ny * dz), diag ? 26.0 : -1.0);
                    }
                }
            }
        } else {
            const double d = kind == 0 ? 4.0 : 6.0;
            if (kind == 1 && z > 0) f(row - nx * ny, -1.0);
            if (y > 0) f(row - nx, -1.0);
            if (x > 0) f(row - 1, -1.0);
            f(row, d);
            if (x < nx - 1) f(row + 1, -1.0);
            if (y < ny - 1) f(row + nx, -1.0);
            if (kind == 1 && z < nz - 1) f(row + nx * ny, -1.0);
        }
    }
    inline int64_t row_len(int64_t row) const
    {
        const int64_t x = row % nx, y = (row / nx) % ny, z = row / (nx * ny);
        const int64_t cx = 1 + (x > 0) + (x < nx - 1), cy = 1 + (y > 0) + (y < ny - 1);
        if (kind == 2) return cx * cy * (1 + (z > 0) + (z < nz - 1));
        const int64_t c = cx + cy - 1;
        return kind == 0 ? c : c + (z > 0) + (z < nz - 1);
    }
};

// SplitMix64 (public-domain mixing function) — the only random source used.
inline uint64_t splitmix64(uint64_t& s)
{
    uint64_t z = (s += 0x9E3779B97F4A7C15ull);
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return z ^ (z >> 31);
}
inline double u01(uint64_t& s) { return (static_cast<double>(splitmix64(s) >> 11) + 0.5) * (1.0 / 9007199254740992.0); }

// Row length of the power-law matrix: 1 diagonal + min(Lmax, ceil(lmin * u^(-1/(a-1)))) - 1 off-diagonals
inline int64_t powerlaw_len(uint64_t seed, int64_t row, double lmin, double a, int64_t lmax, int64_t n)
{
    uint64_t s = seed ^ (0xD1B54A32D192ED03ull * static_cast<uint64_t>(row + 1));
    const double u = u01(s);
    double l = std::ceil(lmin * std::pow(u, -1.0 / (a - 1.0)));
    if (!(l < static_cast<double>(lmax))) l = static_cast<double>(lmax);
    int64_t len = static_cast<int64_t>(l);
    if (len < 1) len = 1;
    if (len > n) len = n;
    return len;
}

template <typename I>
int64_t stencil_row_ptrs(const Stencil& st, int64_t r0, int64_t r1, I* row_ptrs)
{
    const int64_t m = r1 - r0;
    // parallel exclusive scan in two passes
    const int nt = omp_get_max_threads();
    std::vector<int64_t> part(nt + 1, 0);
#pragma omp parallel
    {
        const int t = omp_get_thread_num();
        const int64_t b = m * t / nt, e = m * (t + 1) / nt;
        int64_t s = 0;
        for (int64_t i = b; i < e; ++i) s += st.row_len(r0 + i);
        part[t + 1] = s;
#pragma omp barrier
#pragma omp single
        for (int i = 0; i < nt; ++i) part[i + 1] += part[i];
        s = part[t];
        for (int64_t i = b; i < e; ++i) {
            if (row_ptrs) row_ptrs[i] = static_cast<I>(s);
            s += st.row_len(r0 + i);
        }
    }
    if (row_ptrs) row_ptrs[m] = static_cast<I>(part[nt]);
    return part[nt];
}

template <typename V, typename P, typename C>
int stencil_fill(const Stencil& st, int64_t r0, int64_t r1, const P* row_ptrs, C* cols, V* vals)
{
#pragma omp parallel for schedule(static)
    for (int64_t i = 0; i < r1 - r0; ++i) {
        int64_t k = static_cast<int64_t>(row_ptrs[i]);
        st.for_each(r0 + i, [&](int64_t c, double v) {
            cols[k] = static_cast<C>(c);
            vals[k] = static_cast<V>(v);
            ++k;
        });
    }
    return 0;
}

}  // namespace
}  // namespace gkob200

using namespace gkob200;

extern "C" {

int64_t gkob200_gen_stencil_nnz(int kind, int64_t nx, int64_t ny, int64_t nz, int64_t row_begin, int64_t row_end)
{
    if (kind < 0 || kind > 2 || nx <= 0 || ny <= 0 || (kind != 0 && nz <= 0)) return GKOB200_EINVAL;
    Stencil st{kind, nx, ny, kind == 0 ? 1 : nz};
    if (row_begin < 0 || row_end < row_begin || row_end > st.n()) return GKOB200_EINVAL;
    return stencil_row_ptrs<int64_t>(st, row_begin, row_end, nullptr);
}

#define GKOB200_DEF_GEN(V, VT, P, PT, C, CT)                                                               \
    int gkob200_gen_stencil_csr_##V##_##P##_##C(int kind, int64_t nx, int64_t ny, int64_t nz,              \
                                                int64_t row_begin, int64_t row_end, PT* row_ptrs_host,     \
                                                CT* col_idxs_host, VT* values_host)                        \
    {                                                                                                      \
        if (kind < 0 || kind > 2 || nx <= 0 || ny <= 0 || (kind != 0 && nz <= 0)) return GKOB200_EINVAL;   \
        Stencil st{kind, nx, ny, kind == 0 ? 1 : nz};                                                      \
        if (row_begin < 0 || row_end < row_begin || row_end > st.n()) return GKOB200_EINVAL;               \
        if (!row_ptrs_host || !col_idxs_host || !values_host) return GKOB200_EINVAL;                       \
        stencil_row_ptrs<PT>(st, row_begin, row_end, row_ptrs_host);                                       \
        return stencil_fill<VT>(st, row_begin, row_end, row_ptrs_host, col_idxs_host, values_host);        \
    }
/* local-index CSR (row_ptrs/cols int32), and global-column CSR (row_ptrs int64, cols int64) */
GKOB200_DEF_GEN(f64, double, i32, int32_t, i32, int32_t)
GKOB200_DEF_GEN(f32, float, i32, int32_t, i32, int32_t)
GKOB200_DEF_GEN(f64, double, i64, int64_t, i64, int64_t)
GKOB200_DEF_GEN(f32, float, i64, int64_t, i64, int64_t)

/* Power-law matrix (config 3): n rows, row i has len_i entries: the diagonal plus
 * len_i-1 distinct off-diagonal columns drawn with SplitMix64(seed ^ f(i)), sorted;
 * off-diagonal values U(-1,0), diagonal = 1 + sum|off| (strictly diagonally dominant).
 * len_i = min(lmax, ceil(lmin * u^(-1/(alpha-1)))), u ~ U(0,1). */
int64_t gkob200_gen_powerlaw_row_ptrs_i64(int64_t n, uint64_t seed, double lmin, double alpha, int64_t lmax,
                                          int64_t* row_ptrs_host)
{
    if (n < 0 || !row_ptrs_host || lmin <= 0 || alpha <= 1.0 || lmax < 1) return GKOB200_EINVAL;
    row_ptrs_host[0] = 0;
#pragma omp parallel for schedule(static)
    for (int64_t i = 0; i < n; ++i) row_ptrs_host[i + 1] = powerlaw_len(seed, i, lmin, alpha, lmax, n);
    for (int64_t i = 0; i < n; ++i) row_ptrs_host[i + 1] += row_ptrs_host[i];
    return row_ptrs_host[n];
}

int gkob200_gen_powerlaw_fill_f64_i32(int64_t n, uint64_t seed, const int64_t* row_ptrs_host,
                                      int32_t* row_ptrs32_host, int32_t* col_idxs_host, double* values_host)
{
    if (n < 0 || !row_ptrs_host || !col_idxs_host || !values_host) return GKOB200_EINVAL;
    if (row_ptrs_host[n] > 2147483647ll) return GKOB200_EUNSUPPORTED;
    if (row_ptrs32_host)
        for (int64_t i = 0; i <= n; ++i) row_ptrs32_host[i] = static_cast<int32_t>(row_ptrs_host[i]);
#pragma omp parallel
    {
        std::vector<int32_t> tmp;
#pragma omp for schedule(dynamic, 1024)
        for (int64_t i = 0; i < n; ++i) {
            const int64_t b = row_ptrs_host[i], len = row_ptrs_host[i + 1] - b;
            uint64_t s = seed ^ (0xA0761D6478BD642Full * static_cast<uint64_t>(i + 1));
            tmp.clear();
            tmp.push_back(static_cast<int32_t>(i));
            // distinct columns: draw, sort, unique, top up until len distinct
            while (static_cast<int64_t>(tmp.size()) < len) {
                const int64_t need = len - static_cast<int64_t>(tmp.size());
                for (int64_t t = 0; t < need; ++t)
                    tmp.push_back(static_cast<int32_t>(splitmix64(s) % static_cast<uint64_t>(n)));
                std::sort(tmp.begin(), tmp.end());
                tmp.erase(std::unique(tmp.begin(), tmp.end()), tmp.end());
            }
            double sum = 0.0;
            int64_t dpos = -1;
            for (int64_t t = 0; t < len; ++t) {
                col_idxs_host[b + t] = tmp[t];
                if (tmp[t] == i) {
                    dpos = t;
                } else {
                    const double v = -u01(s);
                    values_host[b + t] = v;
                    sum += -v;
                }
            }
            values_host[b + dpos] = 1.0 + sum;
        }
    }
    return 0;
}

}  // extern "C"
