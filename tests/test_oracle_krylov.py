"""Pins the BiCGSTAB / GMRES / block-Jacobi part of the plain-C oracle bit for bit against
the compiled reference, plus known-answer literals of the reference's unit tests."""
import numpy as np
import pytest
import scipy.sparse as sp

import kat


def block_matrix(n, seed, dtype=np.float64):
    """Diagonally dominant matrix whose rows come in groups with identical sparsity pattern
    (natural Jacobi blocks of varying size) plus some irregular rows."""
    rng = np.random.default_rng(seed)
    rows, cols, vals = [], [], []
    i = 0
    while i < n:
        bs = int(rng.integers(1, 7))
        bs = min(bs, n - i)
        extra = rng.choice(n, size=min(n, int(rng.integers(0, 4))), replace=False)
        pattern = sorted(set(range(i, i + bs)) | set(int(e) for e in extra))
        for r in range(i, i + bs):
            for c in pattern:
                rows.append(r)
                cols.append(c)
                vals.append(rng.uniform(-1, 1) if c != r else 0.0)
        i += bs
    a = sp.csr_matrix((vals, (rows, cols)), shape=(n, n))
    a.setdiag(np.abs(a).sum(axis=1).A1 + 1.0)
    a.sort_indices()
    return a.indptr.astype(np.int32), a.indices.astype(np.int32), a.data.astype(dtype)


def jacobi_defined_mask(J):
    gp, bo, go = J["group_power"], J["block_offset"], J["group_offset"]
    stride = bo << gp
    mask = np.zeros(len(J["blocks"]), dtype=bool)
    for b in range(J["num_blocks"]):
        bs = J["block_ptrs"][b + 1] - J["block_ptrs"][b]
        off = go * (b >> gp) + bo * (b & ((1 << gp) - 1))
        for c in range(bs):
            mask[off + c * stride: off + c * stride + bs] = True
    return mask


@pytest.mark.parametrize("case", kat.JACOBI_FIND_BLOCKS_KATS, ids=lambda c: c[0])
def test_find_blocks_kat(ora, case):
    _, rp, ci, max_bs, expect = case
    nb, ptrs = ora.jacobi_find_blocks(np.array(rp, np.int32), np.array(ci, np.int32), max_bs)
    assert nb == len(expect) - 1 and list(ptrs) == expect


def test_block_inverse_kat(ora):
    # InvertsDiagonalBlocks (reference/test/preconditioner/jacobi_kernels.cpp:275-299), user block pointers {0,2,5}
    m = kat.JACOBI_MTX
    rp, ci, va = np.array(m["row_ptrs"], np.int32), np.array(m["col_idxs"], np.int32), np.array(m["values"])
    bo, go, gp = ora.jacobi_storage_scheme(3)
    ptrs = np.array(kat.JACOBI_BLOCK_PTRS, np.int32)
    blocks = np.zeros(-(-2 // (1 << gp)) * go)
    ora.lib().oracle_jacobi_block_generate_f64(ora.P(rp), ora.P(ci), ora.P(va), ora.i64(2), ora.P(ptrs), ora.i64(bo),
                                               ora.i64(go), int(gp), ora.P(blocks))
    p = bo << gp
    for blk, want in ((0, kat.JACOBI_INV_B1), (1, kat.JACOBI_INV_B2)):
        off = go * (blk >> gp) + bo * (blk & ((1 << gp) - 1))
        got = [[blocks[off + r + c * p] for c in range(len(want))] for r in range(len(want))]
        assert np.allclose(got, want, rtol=kat.rtol(np.float64), atol=0)


@pytest.mark.parametrize("max_bs", [1 + 1, 3, 4, 8, 13, 16, 32])
def test_block_jacobi_matches_reference(ora, refimpl, max_bs):
    rp, ci, va = block_matrix(257, 42)
    J = ora.jacobi_block_generate(rp, ci, va, max_bs)
    R = refimpl.ref_jacobi_generate(rp, ci, va, max_bs)
    assert J["num_blocks"] == R["num_blocks"] and np.array_equal(J["block_ptrs"], R["block_ptrs"])
    assert (J["block_offset"], J["group_offset"], J["group_power"]) == (R["block_offset"], R["group_offset"],
                                                                        R["group_power"])
    # padding slots of the interleaved storage are never written by the reference: compare
    # the defined entries (r, c < block size), bit for bit
    d = jacobi_defined_mask(J)
    assert np.array_equal(J["blocks"][d], R["blocks"][d])


@pytest.mark.parametrize("solver", ["bicgstab", "gmres", "fcg", "cgs"])
@pytest.mark.parametrize("precond_block", [0, 1, 8])
@pytest.mark.parametrize("nrhs", [1, 2])
def test_krylov_matches_reference(ora, refimpl, solver, precond_block, nrhs):
    rp, ci, va = block_matrix(300, 7)
    n = 300
    rng = np.random.default_rng(1)
    b = rng.standard_normal((n, nrhs))
    if nrhs == 2:
        b[:, 1] *= 1e-2
    kind = 0 if precond_block == 0 else (1 if precond_block == 1 else 2)
    diag = sp.csr_matrix((va, ci, rp), shape=(n, n)).diagonal()
    J = ora.jacobi_block_generate(rp, ci, va, precond_block) if kind == 2 else None
    x, it, hist, stop = ora.krylov_solve(solver, rp, ci, va, b, np.zeros_like(b), precond=kind, inv_diag=1.0 / diag,
                                         J=J, max_iters=200, factor=1e-10, krylov_dim=7)
    xr, itr, histr, _ = refimpl.ref_solve(rp, ci, va, b, np.zeros_like(b), solver=solver, precond_block=precond_block,
                                          max_iters=200, factor=1e-10, krylov_dim=7)
    assert it == itr
    assert np.array_equal(x, xr)
    assert np.array_equal(hist, histr[: len(hist)])


# reference/test/solver/bicgstab_kernels.cpp / gmres_kernels.cpp SolvesDenseSystem:
# {{1,-3,0},{-4,1,-3},{2,-1,2}} x = {-1,3,1}  ->  x = {-4,-1,4}
DENSE3 = [[1.0, -3.0, 0.0], [-4.0, 1.0, -3.0], [2.0, -1.0, 2.0]]


@pytest.mark.parametrize("solver", ["bicgstab", "gmres", "cgs"])
def test_solves_dense_system_kat(ora, solver):
    rp, ci, va, _ = kat.dense_to_csr(DENSE3)
    b = np.array([[-1.0], [3.0], [1.0]])
    x, it, _, stop = ora.krylov_solve(solver, rp, ci, va, b, np.zeros_like(b), max_iters=100, factor=kat.rtol(np.float64))
    assert kat.rel_frobenius(x, [[-4.0], [-1.0], [4.0]]) <= kat.rtol(np.float64) * 1e1


@pytest.mark.parametrize("dtype", [np.float64, np.float32])
@pytest.mark.parametrize("solver,case", [("fcg", c) for c in kat.FCG_SOLVE_KATS] + [("cgs", c) for c in kat.CGS_SOLVE_KATS],
                         ids=lambda v: v if isinstance(v, str) else v[0])
def test_fcg_cgs_reference_solve_kats(ora, dtype, solver, case):
    """The literal systems of reference/test/solver/{fcg,cgs}_kernels.cpp with the factories'
    criteria (Iteration(max) + ResidualNorm(r<T>::value)) and the tests' own tolerances."""
    _, A, b, expect, max_iters, tol_mult = case
    rp, ci, va, _ = kat.dense_to_csr(A, dtype)
    b = np.array(b, dtype=dtype)
    x, it, _, _ = ora.krylov_solve(solver, rp, ci, va, b, np.zeros_like(b), max_iters=max_iters, factor=kat.rtol(dtype))
    tol = np.sqrt(kat.rtol(dtype)) if tol_mult == 0.0 else kat.rtol(dtype) * tol_mult
    assert kat.rel_frobenius(x, expect) <= tol
