"""The CUDA path against the committed golden vectors of the UNMODIFIED reference
(tests/golden/ref_golden.npz, made by tests/golden/make_golden.py from oracle/_ref): inputs and the
reference's own outputs come from the file, nothing else is consulted.  Through the C-ABI.
Rules: order-preserving kernels (CSR row-block, SpMM, ELL, SELL-P) and every integer
output bit-identical; solver iteration counts within +-2, early residual norms to 1e-12."""
import os

import numpy as np
import pytest
import torch

from test_gpu_cg import build_solver
from test_gpu_csr import gpu_apply
from test_gpu_formats import npy
from test_oracle_formats import sellp_defined_mask

pytestmark = pytest.mark.gpu

G = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "ref_golden.npz"))
SHAPE = (61, 43)


def spmv_inputs(tag):
    return (G[f"spmv_{tag}_rp"], G[f"spmv_{tag}_ci"], G[f"spmv_{tag}_va"], G[f"spmv_{tag}_b"], G[f"spmv_{tag}_c0"])


@pytest.mark.parametrize("tag", ["f64", "f32"])
def test_csr_apply_and_spmm_bit_identical_to_reference(gko, exec_, tag):
    rp, ci, va, b, c0 = spmv_inputs(tag)
    got, A = gpu_apply(gko, exec_, rp, ci, va, SHAPE, b, "classical")
    assert A.kernel() == "classical"
    assert np.array_equal(got, G[f"spmv_{tag}_plain"])
    got, _ = gpu_apply(gko, exec_, rp, ci, va, SHAPE, b, "classical", 0.7, -1.3, c0)
    assert np.array_equal(got, G[f"spmv_{tag}_adv"])
    # one right-hand side: the single-vector kernel, same summation order
    b1 = np.ascontiguousarray(b[:, :1])
    got, _ = gpu_apply(gko, exec_, rp, ci, va, SHAPE, b1, "classical")
    assert np.array_equal(got, G[f"spmv_{tag}_plain"][:, :1])


@pytest.mark.parametrize("tag", ["f64", "f32"])
@pytest.mark.parametrize("fmt", ["ell", "sellp"])
def test_ell_sellp_apply_bit_identical_to_reference(gko, exec_, tag, fmt):
    rp, ci, va, b, c0 = spmv_inputs(tag)
    A = gko.matrix.Csr.from_arrays(exec_, SHAPE, rp, ci, va)
    M = A.convert_to(fmt)
    db = gko.matrix.Dense.from_numpy(exec_, b)
    dc = gko.matrix.Dense.from_numpy(exec_, c0)
    M.apply(db, dc)
    assert np.array_equal(dc.to_numpy(), G[f"spmv_{tag}_{fmt}"])


def test_conversions_bit_exact_to_reference(gko, exec_):
    rp, ci, va, _, _ = spmv_inputs("f64")
    n = SHAPE[0]
    A = gko.matrix.Csr.from_arrays(exec_, SHAPE, rp, ci, va)
    E = A.convert_to("ell")
    assert (E.width, E.stride) == (int(G["conv_ell_width"]), int(G["conv_ell_stride"]))
    assert np.array_equal(npy(E.col_idxs), G["conv_ell_cols"]) and np.array_equal(npy(E.values), G["conv_ell_vals"])
    S = A.convert_to("sellp", slice_size=8, stride_factor=2)
    sets = G["conv_sellp_slice_sets"]
    assert np.array_equal(npy(S.slice_sets).astype(np.uint64), sets)
    assert np.array_equal(npy(S.slice_lengths).astype(np.uint64), G["conv_sellp_slice_lengths"])
    d = sellp_defined_mask(n, 8, sets)
    assert np.array_equal(npy(S.col_idxs)[d], G["conv_sellp_cols"][d])
    assert np.array_equal(npy(S.values)[d], G["conv_sellp_vals"][d])
    HS = gko.matrix.HybridStrategy
    for strat, kind in ((HS.column_limit(4), "column_limit"), (HS.automatic(), "automatic")):
        H = A.convert_to("hybrid", strategy=strat)
        pre = f"conv_hybrid_{kind}_"
        assert (H.ell.width, H.ell.stride) == (int(G[pre + "ell_width"]), int(G[pre + "ell_stride"])), kind
        assert np.array_equal(npy(H.ell.col_idxs), G[pre + "ell_cols"])
        assert np.array_equal(npy(H.ell.values), G[pre + "ell_vals"])
        assert np.array_equal(npy(H.coo.row_idxs), G[pre + "coo_rows"])
        assert np.array_equal(npy(H.coo.col_idxs), G[pre + "coo_cols"])
        assert np.array_equal(npy(H.coo.values), G[pre + "coo_vals"])


@pytest.mark.parametrize("jacobi", [False, True])
def test_cg_against_reference_history(gko, exec_, jacobi):
    rp, ci, va, b = G["solve_rp"], G["solve_ci"], G["solve_va"], G["solve_b"]
    n = len(rp) - 1
    it_ref = int(G[f"solve_cg_p{int(jacobi)}_it"])
    hist_ref, x_ref = G[f"solve_cg_p{int(jacobi)}_hist"], G[f"solve_cg_p{int(jacobi)}_x"]
    A = gko.matrix.Csr.from_arrays(exec_, (n, n), rp, ci, va)
    solver = build_solver(gko, exec_, A, 400, 1e-10, jacobi)
    db, dx = gko.matrix.Dense.from_numpy(exec_, b), gko.matrix.Dense.create(exec_, (n, 1))
    solver.apply(db, dx)
    assert abs(solver.num_iterations - it_ref) <= 2
    hist = solver.residual_history
    m = min(len(hist), len(hist_ref))
    assert np.allclose(hist[:10], hist_ref[:10], rtol=1e-12)
    assert np.allclose(hist[:m], hist_ref[:m], rtol=1e-6)
    # (a count that differs by one iteration moves x by about cond(A) x tolerance)
    assert np.abs(dx.to_numpy()[:, 0] - x_ref).max() <= 1e-7 * np.abs(x_ref).max()
