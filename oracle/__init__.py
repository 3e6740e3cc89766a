"""TEST INFRASTRUCTURE — Python bindings of the CPU parity oracle.

`oracle.lib()`  : liboracle.so, the plain-C restatement of the reference kernels.
`oracle.ref()`  : oracle/_ref/libgko_refwrap.so, the UNMODIFIED Ginkgo 1.5.0 reference /
                  OpenMP executors compiled from /root/reference (None when not built).

Only tests/, __graft_entry__.smoke() and bench.py's CPU-baseline legs import this
package; the product (repo-8852-ginkgo_b200/) never does.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
i64, f64, f32, vp = C.c_int64, C.c_double, C.c_float, C.c_void_p
_lib = None
_ref = None


def lib():
    global _lib
    if _lib is None:
        path = os.path.join(_HERE, "liboracle.so")
        if not os.path.exists(path):
            raise ImportError(f"{path} missing: run `make -C oracle`")
        _lib = C.CDLL(path)
        for V in ("f64", "f32"):
            getattr(_lib, f"oracle_cg_solve_csr_i32_{V}").restype = i64
    return _lib


def ref():
    """The compiled reference, or None if oracle/_ref was not built."""
    global _ref
    if _ref is None:
        path = os.path.join(_HERE, "_ref", "libgko_refwrap.so")
        if not os.path.exists(path):
            return None
        _ref = C.CDLL(path)
        for name in ("ref_solve_f64_i32", "ref_solve_f32_i32", "ref_solve_f64_i64"):
            getattr(_ref, name).restype = i64
    return _ref


def P(a):
    return vp(a.ctypes.data) if a is not None else None


def _v(dtype):
    return "f64" if np.dtype(dtype) == np.float64 else "f32"


def _i(dtype):
    return "i32" if np.dtype(dtype) == np.int32 else "i64"


def _c(dtype, x):
    return f64(x) if np.dtype(dtype) == np.float64 else f32(x)


# ---- plain-C oracle ---------------------------------------------------------
def csr_spmv(rp, ci, va, b, alpha=None, beta=None, c=None):
    """c = A b, or c = alpha A b + beta c (b, c: n x k arrays)."""
    b2 = np.ascontiguousarray(b.reshape(len(b), -1))
    n, k = len(rp) - 1, b2.shape[1]
    V, I = _v(va.dtype), _i(rp.dtype)
    out = np.zeros((n, k), dtype=va.dtype) if c is None else np.ascontiguousarray(c.reshape(n, -1)).copy()
    if alpha is None:
        getattr(lib(), f"oracle_csr_spmv_{I}_{V}")(i64(n), P(rp), P(ci), P(va), P(b2), i64(k), i64(k), P(out), i64(k))
    else:
        getattr(lib(), f"oracle_csr_advanced_spmv_{I}_{V}")(i64(n), P(rp), P(ci), P(va), _c(va.dtype, alpha), P(b2),
                                                             i64(k), i64(k), _c(va.dtype, beta), P(out), i64(k))
    return out if b.ndim > 1 else out[:, 0]


def cg_solve(rp, ci, va, b, x0, precond=0, inv_diag=None, max_iters=1000, factor=1e-8, baseline=0):
    """Returns (x, iterations, residual_history, stop_status)."""
    n = len(rp) - 1
    b2 = np.ascontiguousarray(b.reshape(n, -1))
    k = b2.shape[1]
    x = np.ascontiguousarray(x0.reshape(n, -1)).copy()
    V = _v(va.dtype)
    hist = np.zeros(max_iters + 2, dtype=va.dtype)
    stop = np.zeros(k, dtype=np.uint8)
    it = getattr(lib(), f"oracle_cg_solve_csr_i32_{V}")(
        i64(n), P(rp), P(ci), P(va), int(precond), P(inv_diag), i64(max_iters), _c(va.dtype, factor), int(baseline),
        i64(k), P(b2), i64(k), P(x), i64(k), P(hist), i64(len(hist)), P(stop))
    return x.reshape(x0.shape), int(it), hist[: it + 1].astype(np.float64), stop


# ---- compiled reference -----------------------------------------------------
FORMATS = {"csr": 0, "ell": 1, "sellp": 2, "coo": 3, "hybrid": 4}
SOLVERS = {"cg": 0, "bicgstab": 1, "gmres": 2}


def ref_spmv(rp, ci, va, b, n_cols=None, alpha=None, beta=None, c=None, fmt="csr", hybrid_limit=-1, omp=False,
             reps=1):
    """Reference-executor (or OMP) apply; returns (c, (mean_s, best_s))."""
    r = ref()
    b2 = np.ascontiguousarray(b.reshape(len(b), -1))
    n, k = len(rp) - 1, b2.shape[1]
    n_cols = len(b2) if n_cols is None else n_cols
    V, I = _v(va.dtype), _i(rp.dtype)
    out = np.zeros((n, k), dtype=va.dtype) if c is None else np.ascontiguousarray(c.reshape(n, -1)).copy()
    secs = np.zeros(2)
    a_ = np.array([alpha], dtype=va.dtype) if alpha is not None else None
    b_ = np.array([beta], dtype=va.dtype) if beta is not None else None
    rc = getattr(r, f"ref_spmv_{V}_{I}")(int(omp), FORMATS[fmt], i64(hybrid_limit), i64(n), i64(n_cols), i64(len(ci)),
                                          P(rp), P(ci), P(va), P(b2), i64(k), i64(k), P(a_), P(b_), P(out), i64(k),
                                          int(reps), P(secs))
    if rc != 0:
        raise RuntimeError(f"ref_spmv rc={rc}")
    return (out if b.ndim > 1 else out[:, 0]), tuple(secs)


def ref_solve(rp, ci, va, b, x0, solver="cg", fmt="csr", hybrid_limit=-1, precond_block=0, max_iters=1000,
              factor=1e-8, baseline=0, krylov_dim=30, omp=False, want_hist=True):
    """Returns (x, iterations, residual_history, seconds)."""
    r = ref()
    n = len(rp) - 1
    b2 = np.ascontiguousarray(b.reshape(n, -1))
    k = b2.shape[1]
    x = np.ascontiguousarray(x0.reshape(n, -1)).copy()
    V, I = _v(va.dtype), _i(rp.dtype)
    hist = np.zeros(max_iters + 2) if want_hist else None
    hl = i64(0)
    secs = f64(0)
    it = getattr(r, f"ref_solve_{V}_{I}")(
        int(omp), SOLVERS[solver], FORMATS[fmt], i64(hybrid_limit), i64(n), i64(len(ci)), P(rp), P(ci), P(va),
        int(precond_block), i64(max_iters), f64(factor), int(baseline), i64(krylov_dim), i64(k), P(b2), P(x),
        P(hist), i64(len(hist) if want_hist else 0), C.byref(hl), C.byref(secs))
    if it < 0:
        raise RuntimeError(f"ref_solve rc={it}")
    return x.reshape(x0.shape), int(it), (hist[: hl.value] if want_hist else None), secs.value


def ref_threads():
    return int(ref().ref_num_threads())
