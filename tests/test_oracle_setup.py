"""Matrix assembly steps: the plain-C oracle against the unmodified reference
executor (oracle/_ref).  The reference's own tests for these kernels are randomised
(test/base/device_matrix_data_kernels.cpp:63-110: 100 x 200, random entries + numerical zeros +
duplicated positions; reference/test/matrix/csr_kernels.cpp transposes / sorts small literal
matrices): the same constructions are used here, plus small literal instances."""
import numpy as np
import pytest
import scipy.sparse as sp

import oracle

needs_ref = pytest.mark.skipif(oracle.ref() is None, reason="oracle/_ref not built")


def random_coo(n_rows, n_cols, nnz, seed, dtype=np.float64, idtype=np.int32, dup=True):
    rng = np.random.default_rng(seed)
    rows = rng.integers(0, n_rows, nnz).astype(idtype)
    cols = rng.integers(0, n_cols if not dup else max(1, n_cols // 8), nnz).astype(idtype)
    vals = rng.integers(-3, 4, nnz).astype(dtype)      # small integers: sums are exact, zeros appear
    return rows, cols, vals


def test_sum_duplicates_kat():
    # entries of one position add up in input order, starting from zero
    # (reference/base/device_matrix_data_kernels.cpp:117-158); zeros are dropped afterwards (:82-106)
    rows = np.array([0, 0, 0, 1, 1, 2, 2, 2], dtype=np.int32)
    cols = np.array([0, 0, 1, 1, 1, 0, 2, 2], dtype=np.int32)
    vals = np.array([1.0, 2.0, 3.0, 4.0, -4.0, 5.0, 6.0, 0.5])
    r, c, v = oracle.coo_assemble(rows, cols, vals, oracle.SUM_DUPLICATES)
    assert r.tolist() == [0, 0, 1, 2, 2] and c.tolist() == [0, 1, 1, 0, 2]
    assert v.tolist() == [3.0, 3.0, 0.0, 5.0, 6.5]
    r, c, v = oracle.coo_assemble(r, c, v, oracle.REMOVE_ZEROS)
    assert r.tolist() == [0, 0, 2, 2] and v.tolist() == [3.0, 3.0, 5.0, 6.5]


def test_transpose_and_sort_kats():
    # the 2 x 3 fixture of reference/test/matrix/csr_kernels.cpp:98-117: [[1,3,2],[0,5,0]] -> [[1,0],[3,5],[2,0]]
    rp = np.array([0, 3, 4], dtype=np.int32)
    ci = np.array([0, 1, 2, 1], dtype=np.int32)
    va = np.array([1.0, 3.0, 2.0, 5.0])
    trp, tci, tva = oracle.csr_transpose(2, 3, rp, ci, va)
    assert trp.tolist() == [0, 1, 3, 4] and tci.tolist() == [0, 0, 1, 0] and tva.tolist() == [1.0, 3.0, 5.0, 2.0]
    # SortSortedMatrix / SortUnsortedMatrix: columns ascending per row, values follow
    ci_u = np.array([2, 0, 1, 1], dtype=np.int32)
    va_u = np.array([2.0, 1.0, 3.0, 5.0])
    c, v = oracle.csr_sort_by_column_index(rp, ci_u, va_u)
    assert c.tolist() == [0, 1, 2, 1] and v.tolist() == [1.0, 3.0, 2.0, 5.0]


@needs_ref
@pytest.mark.parametrize("idtype", [np.int32, np.int64])
@pytest.mark.parametrize("dtype", [np.float64, np.float32])
@pytest.mark.parametrize("shape,nnz", [((50, 40), 600), ((1, 1), 5), ((300, 7), 0), ((1000, 1000), 20000)])
def test_assembly_matches_reference(dtype, idtype, shape, nnz):
    if dtype == np.float32 and idtype == np.int64:
        pytest.skip("combination not instantiated in ref_wrap")
    rows, cols, vals = random_coo(shape[0], shape[1], nnz, 3, dtype, idtype)
    got = oracle.coo_assemble(rows, cols, vals)
    r, c, v, rp = oracle.ref_assemble(shape[0], shape[1], rows, cols, vals)
    assert np.array_equal(got[0], r) and np.array_equal(got[1], c) and np.array_equal(got[2], v)
    assert np.array_equal(rp, np.concatenate([[0], np.cumsum(np.bincount(r, minlength=shape[0]))]).astype(idtype))
    # each step alone: integer outputs identical (value order of duplicates is unspecified in the reference)
    s = oracle.coo_assemble(rows, cols, vals, oracle.SORT)
    rs, cs, vs, _ = oracle.ref_assemble(shape[0], shape[1], rows, cols, vals, oracle.SORT)
    assert np.array_equal(s[0], rs) and np.array_equal(s[1], cs)
    assert np.array_equal(np.sort(s[2]), np.sort(vs))


@needs_ref
@pytest.mark.parametrize("dtype", [np.float64, np.float32])
def test_transpose_and_sort_match_reference(dtype):
    rng = np.random.default_rng(8)
    A = sp.random(200, 130, 0.05, random_state=4, format="csr", dtype=np.float64)
    A.data = rng.uniform(-1, 1, A.nnz)
    A.sort_indices()
    rp, ci, va = A.indptr.astype(np.int32), A.indices.astype(np.int32), A.data.astype(dtype)
    got = oracle.csr_transpose(200, 130, rp, ci, va)
    want = oracle.ref_csr_op(0, 200, 130, rp, ci, va)
    assert all(np.array_equal(g, w) for g, w in zip(got, want))
    # shuffle the columns inside every row, sort them back
    ci_u, va_u = ci.copy(), va.copy()
    for r in range(200):
        p = rng.permutation(rp[r + 1] - rp[r]) + rp[r]
        ci_u[rp[r]:rp[r + 1]], va_u[rp[r]:rp[r + 1]] = ci[p], va[p]
    c, v = oracle.csr_sort_by_column_index(rp, ci_u, va_u)
    _, wc, wv = oracle.ref_csr_op(1, 200, 130, rp, ci_u, va_u)
    assert np.array_equal(c, wc) and np.array_equal(v, wv) and np.array_equal(c, ci) and np.array_equal(v, va)
