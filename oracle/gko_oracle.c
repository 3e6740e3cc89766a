/* gko_oracle.c — TEST INFRASTRUCTURE: CPU parity oracle for libgko_b200.so.
 *
 * A plain-C restatement of the Ginkgo 1.5.0 reference-executor kernels on the
 * sparse-solve hot path.  Pinned against (i) the known-answer
 * tests of the reference's own test-suite, restated in tests/test_oracle_kat.py,
 * and (ii) the unmodified reference compiled from /root/reference into
 * oracle/_ref (oracle/Makefile.ref), compared in tests/test_oracle_vs_ref.py.
 *
 * NOT product code: only tests/, __graft_entry__.smoke() and bench.py's CPU
 * baseline may load liboracle.so.  The product path never calls it.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

/* ---- integer kernels ------------------------------------------------------ */
#define ORACLE_INDEX_UNSIGNED
#define IT uint64_t
#define ISFX(name) name##_u64
#include "oracle_index.inc"
#undef IT
#undef ISFX
#undef ORACLE_INDEX_UNSIGNED
#define IT int32_t
#define ISFX(name) name##_i32
#define ORACLE_INDEX_I32
#include "oracle_index.inc"
#undef ORACLE_INDEX_I32
#undef IT
#undef ISFX
#define IT int64_t
#define ISFX(name) name##_i64
#include "oracle_index.inc"
#undef IT
#undef ISFX

static int cmp_u64(const void* a, const void* b)
{
    const uint64_t x = *(const uint64_t*)a, y = *(const uint64_t*)b;
    return x < y ? -1 : x > y;
}
/* Hybrid strategies  [include/ginkgo/core/matrix/hybrid.hpp:178-380]: ELL width from the
 * SORTED row lengths.  kind 0 column_limit(param_i), 1 imbalance_limit(percent),
 * 2 imbalance_bounded_limit(percent, ratio), 3 minimal_storage_limit (percent given by
 * the caller = sizeof(I)/(sizeof(V)+2 sizeof(I))), 4 automatic = bounded(1/3, 0.001).
 * Also applies Csr::convert_to(Hybrid)'s clamp ell_lim <= num_cols (core/matrix/csr.cpp:300-303). */
uint64_t oracle_hybrid_ell_width(const uint64_t* row_nnz, int64_t n, int kind, uint64_t param_i, double percent,
                                 double ratio, int64_t num_cols)
{
    uint64_t res = 0;
    if (kind == 0) {
        res = param_i;
    } else {
        if (kind == 4) {
            percent = 1.0 / 3.0;
            ratio = 0.001;
        }
        if (percent > 1.0) percent = 1.0;
        if (percent < 0.0) percent = 0.0;
        if (n == 0) {
            res = 0;
        } else {
            uint64_t* tmp = (uint64_t*)malloc(sizeof(uint64_t) * (size_t)n);
            memcpy(tmp, row_nnz, sizeof(uint64_t) * (size_t)n);
            qsort(tmp, (size_t)n, sizeof(uint64_t), cmp_u64);
            if (percent < 1) {
                res = tmp[(uint64_t)(n * percent)];
            } else {
                res = tmp[n - 1];
            }
            free(tmp);
        }
        if (kind == 2 || kind == 4) {
            const uint64_t bound = (uint64_t)(n * ratio);
            if (bound < res) res = bound;
        }
    }
    if (res > (uint64_t)num_cols) res = (uint64_t)num_cols;
    return res;
}
/* [common/unified/matrix/hybrid_kernels.cpp:51-64] */
void oracle_hybrid_compute_coo_row_ptrs(const uint64_t* row_nnz, int64_t n, uint64_t ell_lim, int64_t* coo_row_ptrs)
{
    for (int64_t i = 0; i < n; ++i) {
        const int64_t d = (int64_t)row_nnz[i] - (int64_t)ell_lim;
        coo_row_ptrs[i] = d > 0 ? d : 0;
    }
    oracle_prefix_sum_i64(coo_row_ptrs, n + 1);
}

#define V double
#define SFX(name) name##_f64
#define SQRT sqrt
#define FABS fabs
#include "oracle_kernels.inc"
#include "oracle_setup.inc"
#undef V
#undef SFX
#undef SQRT
#undef FABS

#define V float
#define SFX(name) name##_f32
#define SQRT sqrtf
#define FABS fabsf
#include "oracle_kernels.inc"
#include "oracle_setup.inc"
#undef V
#undef SFX
#undef SQRT
#undef FABS

#include "oracle_dist.inc"

/* Synthetic BASELINE stencils (SURVEY.md §8d "Synthetic inputs"), written independently of the
 * product's generator (csrc/generators.cu) so that the reference arm of bench.py never loads the
 * product library and the two generators check each other (tests/test_oracle_gen.py).
 * kind 0: 2D 5-pt (diag 4), 1: 3D 7-pt (diag 6), 2: 3D 27-pt (diag 26); off-diagonals -1;
 * Dirichlet truncation; rows [row_begin, row_end), GLOBAL columns ascending.
 * Pass cols == NULL to count: returns the number of entries. */
int64_t oracle_gen_stencil_csr(int kind, int64_t nx, int64_t ny, int64_t nz, int64_t row_begin, int64_t row_end,
                               int32_t* row_ptrs, int32_t* cols, double* vals)
{
    if (kind == 0) nz = 1;
    int64_t k = 0;
    for (int64_t row = row_begin; row < row_end; ++row) {
        const int64_t x = row % nx, y = (row / nx) % ny, z = row / (nx * ny);
        if (row_ptrs) row_ptrs[row - row_begin] = (int32_t)k;
        for (int dz = -1; dz <= 1; ++dz)
            for (int dy = -1; dy <= 1; ++dy)
                for (int dx = -1; dx <= 1; ++dx) {
                    const int manhattan = abs(dx) + abs(dy) + abs(dz);
                    if (kind != 2 && manhattan > 1) continue;
                    if (kind == 0 && dz != 0) continue;
                    const int64_t xx = x + dx, yy = y + dy, zz = z + dz;
                    if (xx < 0 || xx >= nx || yy < 0 || yy >= ny || zz < 0 || zz >= nz) continue;
                    if (cols) {
                        cols[k] = (int32_t)((zz * ny + yy) * nx + xx);
                        vals[k] = manhattan == 0 ? (kind == 0 ? 4.0 : kind == 1 ? 6.0 : 26.0) : -1.0;
                    }
                    ++k;
                }
    }
    if (row_ptrs) row_ptrs[row_end - row_begin] = (int32_t)k;
    return k;
}

int oracle_version(void) { return 101; }
