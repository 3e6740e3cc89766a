"""CG on the GPU (fused graph path and multi-RHS general path) vs the oracle / reference KATs."""
import numpy as np
import pytest

import kat

pytestmark = pytest.mark.gpu


def build_solver(gko, exec_, A, max_iters, factor, jacobi=False, nrhs=1, baseline="rhs_norm", check_every=8):
    f = gko.solver.Cg.build().with_criteria(gko.stop.Iteration(max_iters), gko.stop.ResidualNorm(factor, baseline))
    if jacobi:
        f = f.with_preconditioner(gko.preconditioner.Jacobi.build().with_max_block_size(1))
    return f.with_check_every(check_every).on(exec_).generate(A, nrhs=nrhs)


@pytest.mark.parametrize("dtype", [np.float64, np.float32])
@pytest.mark.parametrize("case", kat.CG_SOLVE_KATS, ids=lambda c: c[0])
def test_reference_solve_kats(gko, exec_, dtype, case):
    _, Am, b, expect, max_iters, tol_mult = case
    rp, ci, va, shape = kat.dense_to_csr(Am, dtype)
    A = gko.matrix.Csr.from_arrays(exec_, shape, rp, ci, va, strategy="classical")
    b = np.array(b, dtype=dtype)
    solver = build_solver(gko, exec_, A, max_iters, kat.rtol(dtype), nrhs=b.shape[1])
    db = gko.matrix.Dense.from_numpy(exec_, b)
    dx = gko.matrix.Dense.create(exec_, b.shape, db.t.dtype)
    solver.apply(db, dx)
    assert kat.rel_frobenius(dx.to_numpy(), expect) <= kat.rtol(dtype) * tol_mult
    assert solver.num_iterations < max_iters
    assert all(s & 0x80 for s in solver.stop_status)


@pytest.mark.parametrize("jacobi", [False, True])
@pytest.mark.parametrize("kind,dims", [("5pt", (60, 50, 1)), ("27pt", (16, 17, 18))])
def test_fused_cg_matches_oracle_history(gko, exec_, ora, jacobi, kind, dims):
    rp, ci, va, n = gko.gen.stencil_csr(kind, *dims)
    A = gko.matrix.Csr.from_arrays(exec_, (n, n), rp, ci, va)
    b = np.random.default_rng(4).standard_normal(n)
    diag = {"5pt": 4.0, "27pt": 26.0}[kind]
    inv = 1.0 / np.full(n, diag)
    x_ref, it_ref, hist_ref, stop_ref = ora.cg_solve(rp, ci, va, b, np.zeros(n), precond=int(jacobi), inv_diag=inv,
                                                     max_iters=2000, factor=1e-9)
    solver = build_solver(gko, exec_, A, 2000, 1e-9, jacobi)
    db, dx = gko.matrix.Dense.from_numpy(exec_, b), gko.matrix.Dense.create(exec_, (n, 1))
    solver.apply(db, dx)
    it = solver.num_iterations
    assert abs(it - it_ref) <= 2
    assert list(solver.stop_status) == list(stop_ref)   # converged by criterion 2, finalized
    hist = solver.residual_history
    m = min(len(hist), len(hist_ref))
    # residual histories: early iterations to tight tolerance, whole history loosely
    # (rounding differences of the dot products are amplified by the Krylov recurrence)
    assert np.allclose(hist[:10], hist_ref[:10], rtol=1e-12)
    assert np.allclose(hist[:m], hist_ref[:m], rtol=1e-6)
    assert np.abs(dx.to_numpy()[:, 0] - x_ref).max() <= 1e-9 * np.abs(x_ref).max()
    # second apply: same answer, identical iteration count (workspace reuse, graph reuse)
    dx.fill(0.0)
    solver.apply(db, dx)
    assert solver.num_iterations == it


def test_iteration_limit_and_initial_guess(gko, exec_, ora):
    rp, ci, va, n = gko.gen.stencil_csr("7pt", 14, 15, 16)
    A = gko.matrix.Csr.from_arrays(exec_, (n, n), rp, ci, va)
    rng = np.random.default_rng(5)
    b, x0 = rng.standard_normal(n), rng.standard_normal(n)
    for max_iters in (0, 1, 7, 20):
        x_ref, it_ref, hist_ref, stop_ref = ora.cg_solve(rp, ci, va, b, x0, max_iters=max_iters, factor=1e-30)
        solver = build_solver(gko, exec_, A, max_iters, 1e-30, check_every=3)
        dx = gko.matrix.Dense.from_numpy(exec_, x0)
        solver.apply(gko.matrix.Dense.from_numpy(exec_, b), dx)
        assert solver.num_iterations == it_ref == max_iters
        assert list(solver.stop_status) == list(stop_ref) == [0x41]  # stopped by Iteration (id 1), finalized
        assert np.allclose(dx.to_numpy()[:, 0], x_ref, rtol=1e-10, atol=1e-12)
        assert np.allclose(solver.residual_history, hist_ref, rtol=1e-10)


@pytest.mark.parametrize("baseline,code", [("rhs_norm", 0), ("initial_resnorm", 1), ("absolute", 2)])
def test_baselines(gko, exec_, ora, baseline, code):
    rp, ci, va, n = gko.gen.stencil_csr("5pt", 30, 30)
    A = gko.matrix.Csr.from_arrays(exec_, (n, n), rp, ci, va)
    rng = np.random.default_rng(6)
    b, x0 = 100 * rng.standard_normal(n), rng.standard_normal(n)
    _, it_ref, _, _ = ora.cg_solve(rp, ci, va, b, x0, max_iters=500, factor=1e-6, baseline=code)
    solver = build_solver(gko, exec_, A, 500, 1e-6, baseline=baseline)
    dx = gko.matrix.Dense.from_numpy(exec_, x0)
    solver.apply(gko.matrix.Dense.from_numpy(exec_, b), dx)
    assert abs(solver.num_iterations - it_ref) <= 2


def test_multi_rhs_general_path(gko, exec_, ora):
    rp, ci, va, n = gko.gen.stencil_csr("5pt", 25, 20)
    A = gko.matrix.Csr.from_arrays(exec_, (n, n), rp, ci, va)
    rng = np.random.default_rng(7)
    b = rng.standard_normal((n, 3))
    b[:, 1] *= 1e-3  # columns converge at different iterations
    x_ref, it_ref, _, stop_ref = ora.cg_solve(rp, ci, va, b, np.zeros((n, 3)), max_iters=400, factor=1e-8)
    solver = build_solver(gko, exec_, A, 400, 1e-8, nrhs=3, check_every=4)
    dx = gko.matrix.Dense.create(exec_, (n, 3))
    solver.apply(gko.matrix.Dense.from_numpy(exec_, b), dx)
    assert abs(solver.num_iterations - it_ref) <= 2
    assert list(solver.stop_status) == list(stop_ref)
    assert np.allclose(dx.to_numpy(), x_ref, rtol=1e-7, atol=1e-10)


def test_apply_host_round_trip(gko, exec_, ora):
    import torch
    rp, ci, va, n = gko.gen.stencil_csr("27pt", 12, 12, 12)
    A = gko.matrix.Csr.from_arrays(exec_, (n, n), rp, ci, va)
    b = np.random.default_rng(8).standard_normal(n)
    x_ref, it_ref, _, _ = ora.cg_solve(rp, ci, va, b, np.zeros(n), precond=1, inv_diag=1 / np.full(n, 26.0),
                                       max_iters=300, factor=1e-9)
    solver = build_solver(gko, exec_, A, 300, 1e-9, jacobi=True)
    hb = torch.from_numpy(b).pin_memory()
    hx = torch.zeros(n, dtype=torch.float64).pin_memory()
    solver.apply_host(hb, hx)
    assert abs(solver.num_iterations - it_ref) <= 2
    assert np.abs(hx.numpy() - x_ref).max() <= 1e-9 * np.abs(x_ref).max()
