"""Pins the plain-C oracle against the UNMODIFIED reference compiled into oracle/_ref
(oracle/Makefile.ref): outputs must be bit-identical on seeded inputs.  No GPU."""
import numpy as np
import pytest
import scipy.sparse as sp


def random_csr(n, m, density, seed, dtype=np.float64):
    rng = np.random.default_rng(seed)
    a = sp.random(n, m, density=density, random_state=rng, format="csr", dtype=np.float64)
    a.data = rng.uniform(-1, 1, a.nnz)
    a.sort_indices()
    return a.indptr.astype(np.int32), a.indices.astype(np.int32), a.data.astype(dtype)


@pytest.mark.parametrize("dtype", [np.float64, np.float32])
@pytest.mark.parametrize("nrhs", [1, 3])
def test_csr_spmv_matches_reference_executor(ora, refimpl, dtype, nrhs):
    # sizes of the reference's own device-vs-reference test: 532 x 231 (test/matrix/csr_kernels2.cpp:63-90)
    rp, ci, va = random_csr(532, 231, 0.05, 42, dtype)
    rng = np.random.default_rng(7)
    b = rng.uniform(-1, 1, (231, nrhs)).astype(dtype)
    c0 = rng.uniform(-1, 1, (532, nrhs)).astype(dtype)
    got = ora.csr_spmv(rp, ci, va, b)
    want, _ = refimpl.ref_spmv(rp, ci, va, b)
    assert np.array_equal(got, want)
    got = ora.csr_spmv(rp, ci, va, b, 0.7, -1.3, c0)
    want, _ = refimpl.ref_spmv(rp, ci, va, b, alpha=0.7, beta=-1.3, c=c0)
    assert np.array_equal(got, want)


@pytest.mark.parametrize("precond", [0, 1])
def test_cg_matches_reference_executor(ora, refimpl, gko, precond):
    rp, ci, va, n = gko.gen.stencil_csr("5pt", 40, 40)
    b = np.random.default_rng(3).standard_normal(n)
    inv = 1.0 / np.full(n, 4.0)
    x, it, hist, _ = ora.cg_solve(rp, ci, va, b, np.zeros(n), precond=precond, inv_diag=inv, factor=1e-10)
    xr, itr, histr, _ = refimpl.ref_solve(rp, ci, va, b, np.zeros(n), precond_block=precond, factor=1e-10)
    assert it == itr
    assert np.array_equal(x, xr)
    assert np.array_equal(hist, histr)


def test_omp_executor_agrees_with_reference(refimpl, gko):
    # the CPU baseline (OpenMP executor) solves the same problem in the same number of iterations
    rp, ci, va, n = gko.gen.stencil_csr("7pt", 12, 12, 12)
    b = np.ones(n)
    xr, itr, _, _ = refimpl.ref_solve(rp, ci, va, b, np.zeros(n), factor=1e-9)
    xo, ito, _, _ = refimpl.ref_solve(rp, ci, va, b, np.zeros(n), factor=1e-9, omp=True)
    assert abs(itr - ito) <= 1
    assert np.allclose(xr, xo, rtol=1e-10, atol=1e-12)
