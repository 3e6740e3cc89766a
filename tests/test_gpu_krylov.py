"""Block-Jacobi, BiCGSTAB and GMRES on the GPU vs the oracle (which is pinned bit for bit
against the compiled reference in tests/test_oracle_krylov.py)."""
import numpy as np
import pytest
import scipy.sparse as sp
import torch

import kat
from test_oracle_krylov import block_matrix, jacobi_defined_mask, DENSE3

pytestmark = pytest.mark.gpu


def npy(t):
    return t.detach().cpu().numpy()


def gpu_jacobi(gko, exec_, rp, ci, va, max_bs):
    A = gko.matrix.Csr.from_arrays(exec_, (len(rp) - 1,) * 2, rp, ci, va)
    return A, gko.preconditioner.Jacobi.build().with_max_block_size(max_bs).on(exec_).generate(A)


@pytest.mark.parametrize("case", kat.JACOBI_FIND_BLOCKS_KATS, ids=lambda c: c[0])
def test_find_blocks_reference_kats(gko, exec_, case):
    _, rp, ci, max_bs, expect = case
    rp, ci = np.array(rp, np.int32), np.array(ci, np.int32)
    _, J = gpu_jacobi(gko, exec_, rp, ci, np.ones(len(ci)), max_bs)
    assert J.num_blocks == len(expect) - 1
    assert list(npy(J.block_pointers)[: J.num_blocks + 1]) == expect


@pytest.mark.parametrize("max_bs", [2, 3, 4, 8, 13, 16, 32])
@pytest.mark.parametrize("n,seed", [(257, 42), (5000, 3), (1, 0)])
def test_block_jacobi_generate_bit_exact(gko, exec_, ora, max_bs, n, seed):
    rp, ci, va = block_matrix(n, seed)
    _, J = gpu_jacobi(gko, exec_, rp, ci, va, max_bs)
    R = ora.jacobi_block_generate(rp, ci, va, max_bs)
    assert J.num_blocks == R["num_blocks"]
    assert np.array_equal(npy(J.block_pointers)[: J.num_blocks + 1], R["block_ptrs"])   # index work: bit-exact
    assert (J.block_offset, J.group_offset, J.group_power) == (R["block_offset"], R["group_offset"], R["group_power"])
    d = jacobi_defined_mask(R)
    assert np.array_equal(npy(J.blocks)[d], R["blocks"][d])                             # same pivots, same rounding


def test_find_blocks_long_runs_and_dense_rows(gko, exec_, ora):
    # long runs of identical rows (cap applies inside a run) next to singletons
    n = 3000
    rows = []
    rng = np.random.default_rng(5)
    i = 0
    while i < n:
        run = int(rng.choice([1, 1, 2, 5, 40, 97]))
        run = min(run, n - i)
        pat = sorted(set([i] + [int(c) for c in rng.choice(n, 2)]))
        rows += [pat] * run
        i += run
    rp = np.zeros(n + 1, np.int32)
    rp[1:] = np.cumsum([len(r) for r in rows])
    ci = np.array([c for r in rows for c in r], np.int32)
    for max_bs in (1 + 1, 7, 32):
        nb, ptrs = ora.jacobi_find_blocks(rp, ci, max_bs)
        _, J = gpu_jacobi(gko, exec_, rp, ci, np.ones(len(ci)), max_bs)
        assert J.num_blocks == nb and np.array_equal(npy(J.block_pointers)[: nb + 1], ptrs)


@pytest.mark.parametrize("dtype", [np.float64, np.float32])
@pytest.mark.parametrize("max_bs", [3, 16, 32])
@pytest.mark.parametrize("nrhs", [1, 3])
def test_block_jacobi_apply_bit_exact(gko, exec_, ora, dtype, max_bs, nrhs):
    rp, ci, va = block_matrix(700, 11, dtype)
    _, J = gpu_jacobi(gko, exec_, rp, ci, va, max_bs)
    R = ora.jacobi_block_generate(rp, ci, va, max_bs)
    rng = np.random.default_rng(2)
    b = rng.standard_normal((700, nrhs)).astype(dtype)
    x0 = rng.standard_normal((700, nrhs)).astype(dtype)
    tdt = torch.float64 if dtype == np.float64 else torch.float32
    db, dx = gko.matrix.Dense.from_numpy(exec_, b), gko.matrix.Dense.from_numpy(exec_, x0)
    J.apply(db, dx)
    assert np.array_equal(dx.to_numpy(), ora.jacobi_block_apply(R, b))
    dx = gko.matrix.Dense.from_numpy(exec_, x0)
    J.apply(gko.matrix.Dense.scalar(exec_, 0.5, tdt), db, gko.matrix.Dense.scalar(exec_, -2.0, tdt), dx)
    assert np.array_equal(dx.to_numpy(), ora.jacobi_block_apply(R, b, 0.5, -2.0, x0))


def build(gko, exec_, kind, A, max_iters, factor, precond_block=0, nrhs=1, krylov_dim=30, check_every=8):
    f = getattr(gko.solver, kind).build().with_criteria(gko.stop.Iteration(max_iters), gko.stop.ResidualNorm(factor))
    if precond_block:
        f = f.with_preconditioner(gko.preconditioner.Jacobi.build().with_max_block_size(precond_block))
    return f.with_krylov_dim(krylov_dim).with_check_every(check_every).on(exec_).generate(A, nrhs=nrhs)


@pytest.mark.parametrize("kind", ["Bicgstab", "Gmres"])
def test_reference_dense_kat(gko, exec_, kind):
    # SolvesDenseSystem of reference/test/solver/{bicgstab,gmres}_kernels.cpp: x = {-4,-1,4}
    rp, ci, va, shape = kat.dense_to_csr(DENSE3)
    A = gko.matrix.Csr.from_arrays(exec_, shape, rp, ci, va, strategy="classical")
    s = build(gko, exec_, kind, A, 100, kat.rtol(np.float64))
    b = gko.matrix.Dense.from_numpy(exec_, np.array([[-1.0], [3.0], [1.0]]))
    x = gko.matrix.Dense.create(exec_, (3, 1))
    s.apply(b, x)
    assert kat.rel_frobenius(x.to_numpy(), [[-4.0], [-1.0], [4.0]]) <= kat.rtol(np.float64) * 1e1


@pytest.mark.parametrize("kind", ["Bicgstab", "Gmres", "Fcg", "Cgs"])
@pytest.mark.parametrize("precond_block", [0, 1, 8, 32])
@pytest.mark.parametrize("nrhs", [1, 2])
def test_solvers_match_oracle(gko, exec_, ora, kind, precond_block, nrhs):
    rp, ci, va = block_matrix(2000, 7)
    n = 2000
    A = gko.matrix.Csr.from_arrays(exec_, (n, n), rp, ci, va)
    rng = np.random.default_rng(1)
    b = rng.standard_normal((n, nrhs))
    if nrhs == 2:
        b[:, 1] *= 1e-2
    pk = 0 if precond_block == 0 else (1 if precond_block == 1 else 2)
    diag = sp.csr_matrix((va, ci, rp), shape=(n, n)).diagonal()
    J = ora.jacobi_block_generate(rp, ci, va, precond_block) if pk == 2 else None
    x_ref, it_ref, hist_ref, stop_ref = ora.krylov_solve(kind.lower(), rp, ci, va, b, np.zeros_like(b), precond=pk,
                                                         inv_diag=1.0 / diag, J=J, max_iters=300, factor=1e-10,
                                                         krylov_dim=10)
    s = build(gko, exec_, kind, A, 300, 1e-10, precond_block, nrhs, krylov_dim=10, check_every=3)
    db, dx = gko.matrix.Dense.from_numpy(exec_, b), gko.matrix.Dense.create(exec_, (n, nrhs))
    s.apply(db, dx)
    assert abs(s.num_iterations - it_ref) <= 2
    assert list(s.stop_status) == list(stop_ref)
    m = min(len(s.residual_history), len(hist_ref), 8)
    assert np.allclose(s.residual_history[:m], hist_ref[:m], rtol=1e-9)
    assert np.abs(dx.to_numpy() - x_ref).max() <= 1e-8 * np.abs(x_ref).max()
    Asp = sp.csr_matrix((va, ci, rp), shape=(n, n))
    r = b - Asp @ dx.to_numpy()
    # the true residual: as good as the criterion promises, or — where the recurrence residual
    # drifts from it (FCG on this non-symmetric matrix) — as good as the reference's own solution
    r_ref = np.linalg.norm(b - Asp @ x_ref, axis=0)
    assert np.all(np.linalg.norm(r, axis=0) <= np.maximum(2e-10 * np.linalg.norm(b, axis=0), 2 * r_ref))


@pytest.mark.parametrize("kind", ["Bicgstab", "Gmres"])
def test_iteration_limit_exact(gko, exec_, ora, kind):
    rp, ci, va, n = gko.gen.stencil_csr("7pt", 12, 11, 10)
    A = gko.matrix.Csr.from_arrays(exec_, (n, n), rp, ci, va)
    b = np.random.default_rng(3).standard_normal((n, 1))
    for max_iters in (0, 1, 5, 13):
        x_ref, it_ref, hist_ref, stop_ref = ora.krylov_solve(kind.lower(), rp, ci, va, b, np.zeros_like(b),
                                                             max_iters=max_iters, factor=1e-30, krylov_dim=4)
        s = build(gko, exec_, kind, A, max_iters, 1e-30, krylov_dim=4, check_every=2)
        dx = gko.matrix.Dense.create(exec_, (n, 1))
        s.apply(gko.matrix.Dense.from_numpy(exec_, b), dx)
        assert s.num_iterations == it_ref == max_iters
        assert list(s.stop_status) == list(stop_ref)
        assert np.allclose(dx.to_numpy(), x_ref, rtol=1e-9, atol=1e-12)
        assert np.allclose(s.residual_history, hist_ref, rtol=1e-9)


def test_bicgstab_fp32_hybrid(gko, exec_, ora):
    # config 5 in miniature: BiCGSTAB fp32 on a Hybrid (column_limit 16) 27-pt stencil
    rp, ci, va, n = gko.gen.stencil_csr("27pt", 14, 15, 16, value_dtype=np.float32)
    A = gko.matrix.Csr.from_arrays(exec_, (n, n), rp, ci, va)
    H = A.convert_to("hybrid", strategy=gko.matrix.HybridStrategy.column_limit(16))
    assert H.coo.values.numel() > 0
    b = np.random.default_rng(9).standard_normal((n, 1)).astype(np.float32)
    x_ref, it_ref, _, _ = ora.krylov_solve("bicgstab", rp, ci, va, b, np.zeros_like(b), max_iters=200, factor=1e-4)
    s = build(gko, exec_, "Bicgstab", H, 200, 1e-4)
    dx = gko.matrix.Dense.create(exec_, (n, 1), torch.float32)
    s.apply(gko.matrix.Dense.from_numpy(exec_, b), dx)
    assert abs(s.num_iterations - it_ref) <= 2
    assert np.abs(dx.to_numpy() - x_ref).max() <= 1e-3 * np.abs(x_ref).max()


@pytest.mark.parametrize("dtype", [np.float64, np.float32])
@pytest.mark.parametrize("kind,case", [("Fcg", c) for c in kat.FCG_SOLVE_KATS] + [("Cgs", c) for c in kat.CGS_SOLVE_KATS],
                         ids=lambda v: v if isinstance(v, str) else v[0])
def test_fcg_cgs_reference_solve_kats(gko, exec_, dtype, kind, case):
    """reference/test/solver/{fcg,cgs}_kernels.cpp: the tests' literal systems, criteria and tolerances."""
    _, Am, b, expect, max_iters, tol_mult = case
    rp, ci, va, shape = kat.dense_to_csr(Am, dtype)
    A = gko.matrix.Csr.from_arrays(exec_, shape, rp, ci, va, strategy="classical")
    b = np.array(b, dtype=dtype)
    s = build(gko, exec_, kind, A, max_iters, kat.rtol(dtype), nrhs=b.shape[1])
    db = gko.matrix.Dense.from_numpy(exec_, b)
    dx = gko.matrix.Dense.create(exec_, b.shape, db.t.dtype)
    s.apply(db, dx)
    tol = np.sqrt(kat.rtol(dtype)) if tol_mult == 0.0 else kat.rtol(dtype) * tol_mult
    assert kat.rel_frobenius(dx.to_numpy(), expect) <= tol


@pytest.mark.parametrize("dtype", [np.float64, np.float32])
def test_bicg_step_kernels_match_the_reference_arithmetic(gko, exec_, dtype):
    """bicg::{initialize, step_1, step_2} (common/unified/solver/bicg_kernels.cpp:53-170) through the
    C-ABI against the same arithmetic in numpy (rounded product, rounded sum: bit-identical),
    including a stopped column and the safe_divide(·, 0) = 0 branches."""
    import ctypes as C
    import torch
    lib, V = gko.lib, ("f64" if dtype == np.float64 else "f32")
    tdt = torch.float64 if dtype == np.float64 else torch.float32
    n, k = 1000, 3
    rng = np.random.default_rng(5)
    dev = exec_.device
    T = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(dev)  # noqa: E731
    P = lambda t: C.c_void_p(t.data_ptr())  # noqa: E731
    b = rng.standard_normal((n, k)).astype(dtype)
    vec = {nm: T(rng.standard_normal((n, k)).astype(dtype)) for nm in ("r", "z", "p", "q", "r2", "z2", "p2", "q2")}
    prev_rho, rho = torch.empty(k, dtype=tdt, device=dev), torch.empty(k, dtype=tdt, device=dev)
    stop = torch.full((k,), 7, dtype=torch.uint8, device=dev)
    db = T(b)
    assert getattr(lib, f"gkob200_bicg_initialize_{V}")(None, n, k, P(db), k, *[P(vec[nm]) for nm in
                                                      ("r", "z", "p", "q", "r2", "z2", "p2", "q2")], k,
                                                      P(prev_rho), P(rho), P(stop)) == 0
    assert np.array_equal(vec["r"].cpu().numpy(), b) and np.array_equal(vec["r2"].cpu().numpy(), b)
    for nm in ("z", "p", "q", "z2", "p2", "q2"):
        assert not vec[nm].any()
    assert rho.cpu().tolist() == [0, 0, 0] and prev_rho.cpu().tolist() == [1, 1, 1] and stop.cpu().tolist() == [0, 0, 0]
    # step_1 / step_2 on random state; column 1 stopped, column 2 divides by zero
    st = {nm: rng.standard_normal((n, k)).astype(dtype) for nm in ("x", "r", "r2", "p", "q", "q2", "z", "z2", "p2")}
    rho_h = np.array([0.7, -1.3, 2.0], dtype)
    prev_h = np.array([1.9, 0.4, 0.0], dtype)
    beta_h = np.array([-0.6, 2.2, 0.0], dtype)
    stop_h = np.array([0, 0x41, 0], np.uint8)
    d = {nm: T(a) for nm, a in st.items()}
    d_rho, d_prev, d_beta, d_stop = T(rho_h), T(prev_h), T(beta_h), T(stop_h)
    assert getattr(lib, f"gkob200_bicg_step_1_{V}")(None, n, k, P(d["p"]), P(d["z"]), P(d["p2"]), P(d["z2"]), k,
                                                  P(d_rho), P(d_prev), P(d_stop)) == 0
    tmp = np.where(prev_h == 0, dtype(0), rho_h / np.where(prev_h == 0, dtype(1), prev_h)).astype(dtype)
    live = stop_h == 0
    want_p = np.where(live, st["z"] + (tmp * st["p"]).astype(dtype), st["p"]).astype(dtype)
    want_p2 = np.where(live, st["z2"] + (tmp * st["p2"]).astype(dtype), st["p2"]).astype(dtype)
    assert np.array_equal(d["p"].cpu().numpy(), want_p) and np.array_equal(d["p2"].cpu().numpy(), want_p2)
    assert getattr(lib, f"gkob200_bicg_step_2_{V}")(None, n, k, P(d["x"]), k, P(d["r"]), P(d["r2"]), P(d["p"]), P(d["q"]),
                                                  P(d["q2"]), k, P(d_beta), P(d_rho), P(d_stop)) == 0
    t2 = np.where(beta_h == 0, dtype(0), rho_h / np.where(beta_h == 0, dtype(1), beta_h)).astype(dtype)
    assert np.array_equal(d["x"].cpu().numpy(), np.where(live, st["x"] + (t2 * want_p).astype(dtype), st["x"]).astype(dtype))
    assert np.array_equal(d["r"].cpu().numpy(), np.where(live, st["r"] - (t2 * st["q"]).astype(dtype), st["r"]).astype(dtype))
    assert np.array_equal(d["r2"].cpu().numpy(), np.where(live, st["r2"] - (t2 * st["q2"]).astype(dtype), st["r2"]).astype(dtype))
