"""Pins the CPU oracle (oracle/gko_oracle.c) against the reference's own known-answer
tests (literals restated in tests/kat.py).  No GPU."""
import ctypes as C

import numpy as np
import pytest

import kat

DT = [np.float64, np.float32]


def _csr(dtype, idtype=np.int32):
    m = kat.CSR_MTX
    return (np.array(m["row_ptrs"], dtype=idtype), np.array(m["col_idxs"], dtype=idtype),
            np.array(m["values"], dtype=dtype))


@pytest.mark.parametrize("dtype", DT)
@pytest.mark.parametrize("idtype", [np.int32, np.int64])
@pytest.mark.parametrize("case", kat.CSR_APPLY_KATS, ids=lambda c: c[0])
def test_csr_apply_kat(ora, dtype, idtype, case):
    _, b, alpha, beta, c_in, expect = case
    rp, ci, va = _csr(dtype, idtype)
    b = np.array(b, dtype=dtype)
    c = None if c_in is None else np.array(c_in, dtype=dtype)
    out = ora.csr_spmv(rp, ci, va, b, alpha, beta, c)
    assert np.array_equal(out, np.array(expect, dtype=dtype))  # EXPECT_EQ in the reference


def _cg_vecs(dtype, **fill):
    return {k: np.full((2, 2), v, dtype=dtype) for k, v in fill.items()}


@pytest.mark.parametrize("dtype", DT)
@pytest.mark.parametrize("k", [kat.CG_STEP1, kat.CG_STEP1_DIV0], ids=["Step1", "Step1DivByZero"])
def test_cg_step_1_kat(ora, dtype, k):
    V = "f64" if dtype == np.float64 else "f32"
    v = _cg_vecs(dtype, p=k["p"], z=k["z"])
    rho, prev = np.array(k["rho"], dtype=dtype), np.array(k["prev_rho"], dtype=dtype)
    stop = np.array(k["stop"], dtype=np.uint8)
    getattr(ora.lib(), f"oracle_cg_step_1_{V}")(ora.i64(2), ora.i64(2), ora.P(v["p"]), ora.P(v["z"]), ora.i64(2),
                                                ora.P(rho), ora.P(prev), ora.P(stop))
    assert np.array_equal(v["p"], np.array(k["expect_p"], dtype=dtype))


@pytest.mark.parametrize("dtype", DT)
@pytest.mark.parametrize("k", [kat.CG_STEP2, kat.CG_STEP2_DIV0], ids=["Step2", "Step2DivByZero"])
def test_cg_step_2_kat(ora, dtype, k):
    V = "f64" if dtype == np.float64 else "f32"
    v = _cg_vecs(dtype, x=k["x"], p=k["p"], r=k["r"], q=k["q"])
    rho, beta = np.array(k["rho"], dtype=dtype), np.array(k["beta"], dtype=dtype)
    stop = np.array(k["stop"], dtype=np.uint8)
    getattr(ora.lib(), f"oracle_cg_step_2_{V}")(ora.i64(2), ora.i64(2), ora.P(v["x"]), ora.i64(2), ora.P(v["r"]),
                                                ora.P(v["p"]), ora.P(v["q"]), ora.i64(2), ora.P(beta), ora.P(rho),
                                                ora.P(stop))
    assert np.array_equal(v["x"], np.array(k["expect_x"], dtype=dtype))
    assert np.array_equal(v["r"], np.array(k["expect_r"], dtype=dtype))


def test_cg_initialize_kat(ora):
    # reference/test/solver/cg_kernels.cpp:153-177 KernelInitialize
    n = k = 2
    b = np.full((n, k), 2.0)
    r, z, p, q = np.zeros((n, k)), np.ones((n, k)), np.ones((n, k)), np.ones((n, k))
    prev, rho = np.zeros(k), np.ones(k)
    stop = np.full(k, 0x41, dtype=np.uint8)
    ora.lib().oracle_cg_initialize_f64(ora.i64(n), ora.i64(k), ora.P(b), ora.i64(k), ora.P(r), ora.P(z), ora.P(p),
                                       ora.P(q), ora.i64(k), ora.P(prev), ora.P(rho), ora.P(stop))
    assert np.array_equal(r, b) and not z.any() and not p.any() and not q.any()
    assert np.array_equal(rho, [0, 0]) and np.array_equal(prev, [1, 1]) and not stop.any()


@pytest.mark.parametrize("dtype", DT)
@pytest.mark.parametrize("case", kat.CG_SOLVE_KATS, ids=lambda c: c[0])
def test_cg_solve_kat(ora, dtype, case):
    _, A, b, expect, max_iters, tol_mult = case
    rp, ci, va, _ = kat.dense_to_csr(A, dtype)
    b = np.array(b, dtype=dtype)
    x, it, hist, stop = ora.cg_solve(rp, ci, va, b, np.zeros_like(b), max_iters=max_iters, factor=kat.rtol(dtype))
    assert kat.rel_frobenius(x, expect) <= kat.rtol(dtype) * tol_mult
    assert it < max_iters and all(s & 0x80 for s in stop)  # converged, not iteration-limited


def test_residual_norm_kat(ora):
    # reference/test/stop/residual_norm_kernels.cpp WaitsTillResidualGoal: status flips only
    # when tau < factor * orig_tau; one_changed / all_converged semantics
    tau, orig = np.array([0.5, 1e-9]), np.array([1.0, 1.0])
    stop = np.zeros(2, dtype=np.uint8)
    oc = C.c_int(0)
    allc = ora.lib().oracle_residual_norm_f64(ora.i64(2), ora.P(tau), ora.P(orig), C.c_double(1e-6), 2, 1,
                                              ora.P(stop), C.byref(oc))
    assert allc == 0 and oc.value == 1 and list(stop) == [0, 0xC2]
    tau[0] = 1e-9
    allc = ora.lib().oracle_residual_norm_f64(ora.i64(2), ora.P(tau), ora.P(orig), C.c_double(1e-6), 2, 1,
                                              ora.P(stop), C.byref(oc))
    assert allc == 1 and oc.value == 1 and list(stop) == [0xC2, 0xC2]
