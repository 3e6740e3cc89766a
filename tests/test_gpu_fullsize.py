"""Parity at BASELINE.json's FULL sizes (SURVEY.md §8d "Parity at scale").

One oracle SpMV at full size takes about a second, so every config's SpMV is compared with
the oracle entry by entry (bit-identical where the kernel keeps the storage order, 1e-12
otherwise).  Full-size oracle solves are bounded to the first iterations' residual norms
(the oracle's sequential dot products carry an O(n eps) ~ 1e-9 error at these sizes, the GPU's
tree reductions do not: histories are compared to 1e-7),
plus size-independent properties (linearity, SpMM columns == single SpMVs, format agreement).
"""
import numpy as np
import pytest
import torch

from test_gpu_csr import entry_bound
from test_gpu_cg import build_solver
from test_gpu_krylov import build

pytestmark = pytest.mark.gpu


def dense(gko, exec_, a):
    return gko.matrix.Dense.from_numpy(exec_, a)


def apply(gko, exec_, M, x, dtype=np.float64):
    tdt = torch.float64 if dtype == np.float64 else torch.float32
    y = gko.matrix.Dense.create(exec_, (M.size[0], x.shape[1]), tdt)
    M.apply(dense(gko, exec_, x), y)
    return y.to_numpy()


def test_c1_cg_5pt_1000x1000(gko, exec_, ora):
    rp, ci, va, n = gko.gen.stencil_csr("5pt", 1000, 1000)
    assert n == 10 ** 6 and len(ci) == 4_996_000
    A = gko.matrix.Csr.from_arrays(exec_, (n, n), rp, ci, va)
    x = np.random.default_rng(0).standard_normal((n, 1))
    assert np.array_equal(apply(gko, exec_, A, x), ora.csr_spmv(rp, ci, va, x))
    # CG without preconditioner: the first 200 residual norms and the iterate after 200 iterations
    b = np.ones((n, 1))
    x_ref, it_ref, hist_ref, _ = ora.cg_solve(rp, ci, va, b, np.zeros_like(b), max_iters=200, factor=1e-30)
    s = build_solver(gko, exec_, A, 200, 1e-30)
    dx = gko.matrix.Dense.create(exec_, (n, 1))
    s.apply(dense(gko, exec_, b), dx)
    assert s.num_iterations == it_ref == 200
    assert np.allclose(s.residual_history[:201], hist_ref[:201], rtol=1e-7)
    assert np.abs(dx.to_numpy() - x_ref).max() <= 1e-7 * np.abs(x_ref).max()


def test_c2_cg_jacobi_27pt_200(gko, exec_, ora):
    rp, ci, va, n = gko.gen.stencil_csr("27pt", 200, 200, 200)
    assert n == 8_000_000 and len(ci) == 598 ** 3
    A = gko.matrix.Csr.from_arrays(exec_, (n, n), rp, ci, va)
    assert A.kernel() == "classical"
    rng = np.random.default_rng(1)
    x, y = rng.standard_normal((n, 1)), rng.standard_normal((n, 1))
    ax = apply(gko, exec_, A, x)
    assert np.array_equal(ax, ora.csr_spmv(rp, ci, va, x))          # bit-identical at full size
    S = A.convert_to("sellp")
    assert np.array_equal(apply(gko, exec_, S, x), ax)              # CSR vs SELL-P: same bits
    # linearity within rounding
    lin = apply(gko, exec_, A, x + y) - (ax + apply(gko, exec_, A, y))
    assert (np.abs(lin) / entry_bound(rp, ci, va, np.abs(x) + np.abs(y))).max() <= 1e-14
    # CG + scalar Jacobi: first 12 iterations against the oracle at full size
    b = np.ones((n, 1))
    inv = 1.0 / np.full(n, 26.0)
    x_ref, it_ref, hist_ref, _ = ora.cg_solve(rp, ci, va, b, np.zeros_like(b), precond=1, inv_diag=inv, max_iters=12,
                                              factor=1e-30)
    J = gko.preconditioner.Jacobi.build().with_max_block_size(1).on(exec_).generate(A)
    for M in (A, S):
        s = (gko.solver.Cg.build().with_criteria(gko.stop.Iteration(12), gko.stop.ResidualNorm(1e-30))
             .with_generated_preconditioner(J).on(exec_).generate(M))
        dx = gko.matrix.Dense.create(exec_, (n, 1))
        s.apply(dense(gko, exec_, b), dx)
        assert s.num_iterations == it_ref == 12
        assert np.allclose(s.residual_history[:13], hist_ref[:13], rtol=1e-7)
        assert np.abs(dx.to_numpy() - x_ref).max() <= 1e-7 * np.abs(x_ref).max()


def test_c3_gmres_block_jacobi_powerlaw_10m(gko, exec_, ora):
    n = 10_000_000
    rp, ci, va = gko.gen.powerlaw_csr(n)
    A = gko.matrix.Csr.from_arrays(exec_, (n, n), rp, ci, va)
    assert A.kernel() == "merge_path"
    x = np.random.default_rng(2).standard_normal((n, 1))
    want = ora.csr_spmv(rp, ci, va, x)
    got = apply(gko, exec_, A, x)
    assert (np.abs(got - want) / entry_bound(rp, ci, va, x)).max() <= 1e-12
    # the classical kernel on the same matrix keeps the storage order: bit-identical
    Ac = gko.matrix.Csr.from_arrays(exec_, (n, n), rp, ci, va, strategy="classical")
    assert np.array_equal(apply(gko, exec_, Ac, x), want)
    del Ac
    # block Jacobi(32): block pointers and the preconditioned residual history of GMRES(30)
    J = ora.jacobi_block_generate(rp, ci, va, 32)
    b = np.ones((n, 1))
    _, it_ref, hist_ref, _ = ora.krylov_solve("gmres", rp, ci, va, b, np.zeros_like(b), precond=2, J=J, max_iters=6,
                                              factor=1e-30, krylov_dim=30)
    s = build(gko, exec_, "Gmres", A, 6, 1e-30, precond_block=32, krylov_dim=30)
    dx = gko.matrix.Dense.create(exec_, (n, 1))
    s.apply(dense(gko, exec_, b), dx)
    assert s.num_iterations == it_ref == 6
    assert np.allclose(s.residual_history[:7], hist_ref[:7], rtol=1e-7)


def test_c4_slab_7pt_512x512x64(gko, exec_, ora):
    rp, ci, va, n = gko.gen.stencil_csr("7pt", 512, 512, 64)
    A = gko.matrix.Csr.from_arrays(exec_, (n, n), rp, ci, va)
    x = np.random.default_rng(3).standard_normal((n, 1))
    assert np.array_equal(apply(gko, exec_, A, x), ora.csr_spmv(rp, ci, va, x))


def test_c5_fp32_hybrid_and_spmm_27pt_256(gko, exec_, ora):
    rp, ci, va, n = gko.gen.stencil_csr("27pt", 256, 256, 256, value_dtype=np.float32)
    assert n == 256 ** 3 and len(ci) == 766 ** 3
    A = gko.matrix.Csr.from_arrays(exec_, (n, n), rp, ci, va)
    rng = np.random.default_rng(4)
    x = rng.standard_normal((n, 1)).astype(np.float32)
    want = ora.csr_spmv(rp, ci, va, x)
    assert np.array_equal(apply(gko, exec_, A, x, np.float32), want)
    H = A.convert_to("hybrid")                  # automatic: pure ELL on this matrix
    assert H.ell.width == 27 and H.coo.values.numel() == 0
    assert np.array_equal(apply(gko, exec_, H, x, np.float32), want)
    H16 = A.convert_to("hybrid", strategy=gko.matrix.HybridStrategy.column_limit(16))
    got = apply(gko, exec_, H16, x, np.float32)
    assert (np.abs(got.astype(np.float64) - want) / entry_bound(rp, ci, va, x)).max() <= 1e-5
    del H16
    # 32-RHS SpMM: every column is bit-identical to the single-RHS product of that column
    X = rng.standard_normal((n, 32)).astype(np.float32)
    X[:, 5] = x[:, 0]
    for M in (A, H.ell):
        Y = apply(gko, exec_, M, X, np.float32)
        assert np.array_equal(Y[:, 5], want[:, 0])
        for j in (0, 31):
            assert np.array_equal(Y[:, j], apply(gko, exec_, A, np.ascontiguousarray(X[:, j:j + 1]), np.float32)[:, 0])
        del Y
    # BiCGSTAB (fused path) on the hybrid operator: first 5 iterations.  A sequential fp32 dot
    # product over 1.7e7 entries (the oracle's, and the reference executor's) has no correct
    # digit left (n * eps ~ 1), so the fp32 run is compared with the oracle run in fp64.
    b = np.ones((n, 1), dtype=np.float32)
    _, it_ref, hist_ref, _ = ora.krylov_solve("bicgstab", rp, ci, va.astype(np.float64), b.astype(np.float64),
                                              np.zeros((n, 1)), max_iters=5, factor=1e-30)
    s = build(gko, exec_, "Bicgstab", H, 5, 1e-30)
    dx = gko.matrix.Dense.create(exec_, (n, 1), torch.float32)
    s.apply(dense(gko, exec_, b), dx)
    assert s.num_iterations == it_ref == 5
    assert np.allclose(s.residual_history[:6], hist_ref[:6], rtol=1e-3)
