"""Property tests (hypothesis): the plain-C oracle against the unmodified reference executor on
randomly shaped inputs — empty rows and columns, 1 x n / n x 1, fully dense rows, duplicates —
for the kernels every GPU parity test leans on."""
import numpy as np
import pytest
from hypothesis import HealthCheck, given, settings, strategies as st

import oracle

pytestmark = pytest.mark.skipif(oracle.ref() is None, reason="oracle/_ref not built")
SETTINGS = dict(max_examples=40, deadline=None, suppress_health_check=[HealthCheck.too_slow])


@st.composite
def csr_matrices(draw, max_dim=40):
    n = draw(st.integers(1, max_dim))
    m = draw(st.integers(1, max_dim))
    density = draw(st.sampled_from([0.0, 0.05, 0.3, 1.0]))
    seed = draw(st.integers(0, 2 ** 31 - 1))
    rng = np.random.default_rng(seed)
    mask = rng.random((n, m)) < density
    if draw(st.booleans()) and n > 2:
        mask[rng.integers(0, n)] = False      # an empty row
        mask[rng.integers(0, n)] = True       # a full row
    rp = np.concatenate([[0], np.cumsum(mask.sum(axis=1))]).astype(np.int32)
    ci = np.nonzero(mask)[1].astype(np.int32)
    va = rng.uniform(-2, 2, len(ci))
    return n, m, rp, ci, va, rng


@settings(**SETTINGS)
@given(csr_matrices(), st.integers(1, 4), st.booleans())
def test_csr_spmv_bit_identical(mat, nrhs, advanced):
    n, m, rp, ci, va, rng = mat
    b = rng.standard_normal((m, nrhs))
    c = rng.standard_normal((n, nrhs))
    if advanced:
        got = oracle.csr_spmv(rp, ci, va, b, -0.75, 1.5, c)
        want, _ = oracle.ref_spmv(rp, ci, va, b, n_cols=m, alpha=-0.75, beta=1.5, c=c)
    else:
        got = oracle.csr_spmv(rp, ci, va, b)
        want, _ = oracle.ref_spmv(rp, ci, va, b, n_cols=m)
    assert np.array_equal(got, want.reshape(got.shape))


@settings(**SETTINGS)
@given(csr_matrices())
def test_transpose_bit_identical(mat):
    n, m, rp, ci, va, _ = mat
    got = oracle.csr_transpose(n, m, rp, ci, va)
    want = oracle.ref_csr_op(0, n, m, rp, ci, va)
    assert all(np.array_equal(g, w) for g, w in zip(got, want))


@settings(**SETTINGS)
@given(st.integers(1, 30), st.integers(1, 30), st.integers(0, 400), st.integers(0, 2 ** 31 - 1))
def test_assembly_with_duplicates_and_zeros(n, m, nnz, seed):
    rng = np.random.default_rng(seed)
    rows = rng.integers(0, n, nnz).astype(np.int32)
    cols = rng.integers(0, m, nnz).astype(np.int32)
    vals = rng.integers(-2, 3, nnz).astype(np.float64)      # exact sums, zeros appear
    got = oracle.coo_assemble(rows, cols, vals)
    r, c, v, rp = oracle.ref_assemble(n, m, rows, cols, vals)
    assert np.array_equal(got[0], r) and np.array_equal(got[1], c) and np.array_equal(got[2], v)
