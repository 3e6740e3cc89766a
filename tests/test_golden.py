"""The plain-C oracle against golden vectors computed by the UNMODIFIED reference (tests/golden/ref_golden.npz,
made by tests/golden/make_golden.py from oracle/_ref).  Unlike tests/test_oracle_vs_ref.py this needs neither
the reference tree nor oracle/_ref at run time, and no particular random stream: inputs are in the file.
The -m gpu tests compare the CUDA path with this oracle bit for bit, so the chain reference -> oracle -> CUDA
stays pinned on a box that only has the repository.  No GPU."""
import os

import numpy as np
import pytest

G = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "ref_golden.npz"))


@pytest.mark.parametrize("tag", ["f64", "f32"])
def test_csr_apply_and_spmm_bit_identical_to_reference(ora, tag):
    rp, ci, va = G[f"spmv_{tag}_rp"], G[f"spmv_{tag}_ci"], G[f"spmv_{tag}_va"]
    b, c0 = G[f"spmv_{tag}_b"], G[f"spmv_{tag}_c0"]
    assert np.array_equal(ora.csr_spmv(rp, ci, va, b), G[f"spmv_{tag}_plain"])
    assert np.array_equal(ora.csr_spmv(rp, ci, va, b, 0.7, -1.3, c0), G[f"spmv_{tag}_adv"])
    # single right-hand side = column 0 of the SpMM (the reference's per-column summation order)
    assert np.array_equal(ora.csr_spmv(rp, ci, va, b[:, :1].copy()), G[f"spmv_{tag}_plain"][:, :1])


@pytest.mark.parametrize("tag", ["f64", "f32"])
def test_format_applies_bit_identical_to_reference(ora, tag):
    rp, ci, va, b = G[f"spmv_{tag}_rp"], G[f"spmv_{tag}_ci"], G[f"spmv_{tag}_va"], G[f"spmv_{tag}_b"]
    n = len(rp) - 1
    w, stride, cols, vals = ora.csr_to_ell(rp, ci, va)
    assert np.array_equal(ora.ell_spmv(n, stride, w, cols, vals, b), G[f"spmv_{tag}_ell"])
    sets, lens, scols, svals = ora.csr_to_sellp(rp, ci, va)
    assert np.array_equal(ora.sellp_spmv(n, 64, sets, lens, scols, svals, b), G[f"spmv_{tag}_sellp"])
    rows, ccols, cvals = ora.csr_to_coo(rp, ci, va)
    assert np.array_equal(ora.coo_spmv2(rows, ccols, cvals, b, np.zeros((n, b.shape[1]), dtype=b.dtype)),
                          G[f"spmv_{tag}_coo"])


def test_conversions_bit_exact():
    import oracle as ora
    rp, ci, va = G["spmv_f64_rp"], G["spmv_f64_ci"], G["spmv_f64_va"]
    w, stride, cols, vals = ora.csr_to_ell(rp, ci, va)
    assert (w, stride) == (int(G["conv_ell_width"]), int(G["conv_ell_stride"]))
    assert np.array_equal(cols, G["conv_ell_cols"]) and np.array_equal(vals, G["conv_ell_vals"])
    sets, lens, scols, svals = ora.csr_to_sellp(rp, ci, va, slice_size=8, stride_factor=2)
    assert np.array_equal(sets, G["conv_sellp_slice_sets"]) and np.array_equal(lens, G["conv_sellp_slice_lengths"])
    assert int(sets[-1]) == int(G["conv_sellp_total_cols"])
    # rows past the end of the last slice are never written by the reference kernel: compare real rows only
    n = len(rp) - 1
    gcols, gvals = G["conv_sellp_cols"], G["conv_sellp_vals"]
    for s in range(len(lens)):
        for r in range(8):
            if s * 8 + r >= n:
                continue
            idx = int(sets[s]) * 8 + r + 8 * np.arange(int(lens[s]))
            assert np.array_equal(scols[idx], gcols[idx]) and np.array_equal(svals[idx], gvals[idx])
    for kind, param in (("column_limit", 4), ("automatic", 0)):
        h = ora.csr_to_hybrid(rp, ci, va, 43, kind=kind, param=param)
        pre = f"conv_hybrid_{kind}_"
        assert h["ell_width"] == int(G[pre + "ell_width"]) and h["ell_stride"] == int(G[pre + "ell_stride"])
        for k in ("ell_cols", "ell_vals", "coo_rows", "coo_cols", "coo_vals"):
            assert np.array_equal(h[k], G[pre + k]), (kind, k)


@pytest.mark.parametrize("precond", [0, 1])
def test_cg_history_bit_identical_to_reference(ora, precond):
    rp, ci, va, b = G["solve_rp"], G["solve_ci"], G["solve_va"], G["solve_b"]
    n = len(rp) - 1
    inv = 1.0 / np.full(n, 4.0)
    x, it, hist, _ = ora.cg_solve(rp, ci, va, b, np.zeros(n), precond=precond, inv_diag=inv, factor=1e-10, max_iters=400)
    assert it == int(G[f"solve_cg_p{precond}_it"])
    assert np.array_equal(x, G[f"solve_cg_p{precond}_x"])
    assert np.array_equal(hist, G[f"solve_cg_p{precond}_hist"])


@pytest.mark.parametrize("solver", ["bicgstab", "gmres", "fcg", "cgs"])
@pytest.mark.parametrize("precond", [0, 1])
def test_krylov_histories_bit_identical_to_reference(ora, solver, precond):
    rp, ci, va, b = G["solve_rp"], G["solve_ci"], G["solve_va"], G["solve_b"]
    n = len(rp) - 1
    inv = 1.0 / np.full(n, 4.0)
    x, it, hist, _ = ora.krylov_solve(solver, rp, ci, va, b, np.zeros(n), precond=precond, inv_diag=inv, factor=1e-10,
                                      max_iters=400)
    assert it == int(G[f"solve_{solver}_p{precond}_it"])
    assert np.array_equal(x, G[f"solve_{solver}_p{precond}_x"])
    assert np.array_equal(hist, G[f"solve_{solver}_p{precond}_hist"])
