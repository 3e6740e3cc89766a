"""Synthetic generators: structure of the BASELINE stencils and the power-law matrix."""
import numpy as np
import scipy.sparse as sp


def test_stencil_sizes_match_baseline_formulas(gko):
    # nnz of the 27-pt stencil on n^3 is (3n-2)^3, of the 7-pt 7n^3 - 6n^2, of 2D 5-pt 5n^2 - 4n
    assert gko.lib.gkob200_gen_stencil_nnz(2, 200, 200, 200, 0, 200 ** 3) == 598 ** 3
    assert gko.lib.gkob200_gen_stencil_nnz(1, 512, 512, 512, 0, 512 ** 3) == 937_951_232
    assert gko.lib.gkob200_gen_stencil_nnz(0, 1000, 1000, 1, 0, 10 ** 6) == 4_996_000
    assert gko.lib.gkob200_gen_stencil_nnz(2, 256, 256, 256, 0, 256 ** 3) == 766 ** 3


def test_stencil_is_symmetric_sorted_and_slabs_concatenate(gko):
    for kind, dims in (("5pt", (7, 5, 1)), ("7pt", (5, 4, 6)), ("27pt", (4, 5, 3))):
        rp, ci, va, n = gko.gen.stencil_csr(kind, *dims)
        A = sp.csr_matrix((va, ci, rp), shape=(n, n))
        assert (abs(A - A.T)).nnz == 0
        assert all(np.all(np.diff(ci[rp[i]:rp[i + 1]]) > 0) for i in range(n))
        # rows [a,b) generated separately equal the slice of the global matrix (global columns)
        a, b = n // 3, 2 * n // 3
        rp2, ci2, va2, _ = gko.gen.stencil_csr(kind, *dims, row_begin=a, row_end=b, index_dtype=np.int64)
        assert np.array_equal(ci2, ci[rp[a]:rp[b]]) and np.array_equal(va2, va[rp[a]:rp[b]])
        assert np.array_equal(rp2, rp[a:b + 1] - rp[a])


def test_powerlaw_is_diagonally_dominant_and_skewed(gko):
    rp, ci, va = gko.gen.powerlaw_csr(20000, seed=42)
    n = 20000
    A = sp.csr_matrix((va, ci, rp), shape=(n, n))
    d = A.diagonal()
    off = abs(A).sum(axis=1).A1 - abs(d)
    assert np.all(d > off)
    lens = np.diff(rp)
    assert lens.max() > 20 * lens.mean()
    assert all(np.all(np.diff(ci[rp[i]:rp[i + 1]]) > 0) for i in range(0, n, 97))
    rp2, ci2, va2 = gko.gen.powerlaw_csr(20000, seed=42)
    assert np.array_equal(ci, ci2) and np.array_equal(va, va2)
