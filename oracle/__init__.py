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


# ---- formats: plain-C oracle ------------------------------------------------
def _b2(b):
    b2 = np.ascontiguousarray(b.reshape(len(b), -1))
    return b2, b2.shape[1]


def csr_to_coo(rp, ci, va):
    I = _i(rp.dtype)
    rows = np.empty_like(ci)
    getattr(lib(), f"oracle_convert_ptrs_to_idxs_{I}")(P(rp), i64(len(rp) - 1), P(rows))
    return rows, ci.copy(), va.copy()


def max_row_nnz(rp):
    fn = getattr(lib(), f"oracle_compute_max_row_nnz_{_i(rp.dtype)}")
    fn.restype = C.c_uint64
    return int(fn(P(rp), i64(len(rp) - 1)))


def csr_to_ell(rp, ci, va):
    n = len(rp) - 1
    width, stride = max_row_nnz(rp), n
    cols = np.empty(width * stride, dtype=ci.dtype)
    vals = np.empty(width * stride, dtype=va.dtype)
    getattr(lib(), f"oracle_csr_to_ell_{_i(rp.dtype)}_{_v(va.dtype)}")(i64(n), P(rp), P(ci), P(va), i64(width),
                                                                        i64(stride), P(cols), P(vals))
    return width, stride, cols, vals


def slice_sets(rp, slice_size=64, stride_factor=1):
    n = len(rp) - 1
    ns = -(-n // slice_size)
    sets = np.zeros(ns + 1, dtype=np.uint64)
    lens = np.zeros(ns, dtype=np.uint64)
    getattr(lib(), f"oracle_compute_slice_sets_{_i(rp.dtype)}")(P(rp), i64(n), i64(slice_size), i64(stride_factor),
                                                                 P(sets), P(lens))
    return sets, lens


def csr_to_sellp(rp, ci, va, slice_size=64, stride_factor=1):
    n = len(rp) - 1
    sets, lens = slice_sets(rp, slice_size, stride_factor)
    total = int(sets[-1]) * slice_size
    # rows >= n of the last (partial) slice are never written by the reference kernel
    cols = np.zeros(total, dtype=ci.dtype)
    vals = np.zeros(total, dtype=va.dtype)
    getattr(lib(), f"oracle_csr_to_sellp_{_i(rp.dtype)}_{_v(va.dtype)}")(i64(n), P(rp), P(ci), P(va), i64(slice_size),
                                                                          P(sets), P(cols), P(vals))
    return sets, lens, cols, vals


HYB_KINDS = {"column_limit": 0, "imbalance_limit": 1, "imbalance_bounded_limit": 2, "minimal_storage_limit": 3,
             "automatic": 4}


def hybrid_ell_width(rp, n_cols, kind="automatic", param=0, percent=0.8, ratio=0.0001):
    n = len(rp) - 1
    sizes = np.diff(rp).astype(np.uint64)
    fn = lib().oracle_hybrid_ell_width
    fn.restype = C.c_uint64
    return int(fn(P(sizes), i64(n), HYB_KINDS[kind], C.c_uint64(param), f64(percent), f64(ratio), i64(n_cols)))


def csr_to_hybrid(rp, ci, va, n_cols, kind="automatic", param=0, percent=0.8, ratio=0.0001):
    n = len(rp) - 1
    if kind == "minimal_storage_limit":
        percent = ci.itemsize / (va.itemsize + 2 * ci.itemsize)
        kind_ = "imbalance_limit"
    else:
        kind_ = kind
    w = hybrid_ell_width(rp, n_cols, kind_, param, percent, ratio)
    sizes = np.diff(rp).astype(np.uint64)
    coo_ptrs = np.zeros(n + 1, dtype=np.int64)
    lib().oracle_hybrid_compute_coo_row_ptrs(P(sizes), i64(n), C.c_uint64(w), P(coo_ptrs))
    coo_nnz = int(coo_ptrs[n])
    ecols, evals = np.empty(w * n, dtype=ci.dtype), np.empty(w * n, dtype=va.dtype)
    crows, ccols = np.empty(coo_nnz, dtype=ci.dtype), np.empty(coo_nnz, dtype=ci.dtype)
    cvals = np.empty(coo_nnz, dtype=va.dtype)
    getattr(lib(), f"oracle_csr_to_hybrid_{_i(rp.dtype)}_{_v(va.dtype)}")(
        i64(n), P(rp), P(ci), P(va), P(coo_ptrs), i64(n), i64(w), P(ecols), P(evals), P(crows), P(ccols), P(cvals))
    return dict(ell_width=w, ell_stride=n, ell_cols=ecols, ell_vals=evals, coo_rows=crows, coo_cols=ccols,
                coo_vals=cvals, coo_row_ptrs=coo_ptrs)


def ell_spmv(n, stride, width, cols, vals, b, alpha=None, beta=None, c=None):
    b2, k = _b2(b)
    out = np.zeros((n, k), dtype=vals.dtype) if c is None else np.ascontiguousarray(c.reshape(n, -1)).copy()
    adv = alpha is not None
    getattr(lib(), f"oracle_ell_spmv_{_i(cols.dtype)}_{_v(vals.dtype)}")(
        i64(n), i64(stride), i64(width), P(cols), P(vals), int(adv), _c(vals.dtype, alpha or 0), P(b2), i64(k), i64(k),
        _c(vals.dtype, beta or 0), P(out), i64(k))
    return out


def sellp_spmv(n, slice_size, sets, lens, cols, vals, b, alpha=None, beta=None, c=None):
    b2, k = _b2(b)
    out = np.zeros((n, k), dtype=vals.dtype) if c is None else np.ascontiguousarray(c.reshape(n, -1)).copy()
    adv = alpha is not None
    getattr(lib(), f"oracle_sellp_spmv_{_i(cols.dtype)}_{_v(vals.dtype)}")(
        i64(n), i64(slice_size), P(sets), P(lens), P(cols), P(vals), int(adv), _c(vals.dtype, alpha or 0), P(b2),
        i64(k), i64(k), _c(vals.dtype, beta or 0), P(out), i64(k))
    return out


def coo_spmv2(rows, cols, vals, b, c, alpha=None):
    """c += [alpha] A b (returns the updated copy of c)."""
    b2, k = _b2(b)
    out = np.ascontiguousarray(c.reshape(len(c), -1)).copy()
    getattr(lib(), f"oracle_coo_spmv2_{_i(cols.dtype)}_{_v(vals.dtype)}")(
        i64(len(vals)), P(rows), P(cols), P(vals), int(alpha is not None), _c(vals.dtype, alpha or 0), P(b2), i64(k),
        i64(k), P(out), i64(k))
    return out


def prefix_sum(a):
    name = {np.dtype(np.int32): "i32", np.dtype(np.int64): "i64", np.dtype(np.uint64): "u64"}[a.dtype]
    out = a.copy()
    getattr(lib(), f"oracle_prefix_sum_{name}")(P(out), i64(len(out)))
    return out


# ---- formats: compiled reference ---------------------------------------------
def ref_convert(rp, ci, va, n_cols, fmt, hyb_kind="automatic", hyb_param=0, percent=0.8, ratio=0.0001, slice_size=64,
                stride_factor=1):
    """Csr::convert_to(<fmt>) on the ReferenceExecutor; returns a dict of the result arrays."""
    r = ref()
    n, nnz = len(rp) - 1, len(ci)
    V, I = _v(va.dtype), _i(rp.dtype)
    cap = max(1, (n + 64) * (int(np.diff(rp).max()) if n else 1) + 64 * 64)
    meta = np.zeros(4, dtype=np.int64)
    ia, ib = np.empty(cap, dtype=ci.dtype), np.empty(cap, dtype=ci.dtype)
    ov = np.empty(cap, dtype=va.dtype)
    ua, ub = np.zeros(n + 2, dtype=np.uint64), np.zeros(n + 2, dtype=np.uint64)
    cr, cc, cv = np.empty(cap, dtype=ci.dtype), np.empty(cap, dtype=ci.dtype), np.empty(cap, dtype=va.dtype)
    fn = getattr(r, f"ref_convert_{V}_{I}")
    fn.restype = i64
    total = fn(FORMATS[fmt], HYB_KINDS[hyb_kind], i64(hyb_param), f64(percent), f64(ratio), i64(slice_size),
               i64(stride_factor), i64(n), i64(n_cols), i64(nnz), P(rp), P(ci), P(va), P(meta), P(ia), P(ib), P(ov),
               P(ua), P(ub), P(cr), P(cc), P(cv), i64(cap))
    if total < 0:
        raise RuntimeError(f"ref_convert rc={total}")
    if fmt == "ell":
        return dict(width=int(meta[0]), stride=int(meta[1]), cols=ia[:total].copy(), vals=ov[:total].copy())
    if fmt == "sellp":
        ns = int(meta[1])
        return dict(total_cols=int(meta[0]), slice_sets=ua[:ns + 1].copy(), slice_lengths=ub[:ns].copy(),
                    cols=ia[:total].copy(), vals=ov[:total].copy())
    if fmt == "coo":
        return dict(rows=ia[:total].copy(), cols=ib[:total].copy(), vals=ov[:total].copy())
    if fmt == "hybrid":
        cn = int(meta[2])
        return dict(ell_width=int(meta[0]), ell_stride=int(meta[1]), ell_cols=ia[:total].copy(),
                    ell_vals=ov[:total].copy(), coo_rows=cr[:cn].copy(), coo_cols=cc[:cn].copy(),
                    coo_vals=cv[:cn].copy())
    raise ValueError(fmt)


# ---- block-Jacobi, BiCGSTAB, GMRES -------------------------------------------
def jacobi_storage_scheme(max_block_size, max_block_stride=32):
    """compute_storage_scheme (include/ginkgo/core/preconditioner/jacobi.hpp:578-610)."""
    sup = 1
    while sup < max_block_size:
        sup *= 2
    gs = max_block_stride // sup
    block_offset = max_block_size
    return block_offset, max_block_size * gs * block_offset, gs.bit_length() - 1


def jacobi_find_blocks(rp, ci, max_block_size):
    n = len(rp) - 1
    ptrs = np.zeros(n + 1, dtype=np.int32)
    fn = lib().oracle_jacobi_find_blocks_i32
    fn.restype = i64
    nb = fn(i64(n), P(rp), P(ci), C.c_int32(max_block_size), P(ptrs))
    return int(nb), ptrs[: nb + 1].copy()


def jacobi_block_generate(rp, ci, va, max_block_size):
    """Returns dict(num_blocks, block_ptrs, blocks, block_offset, group_offset, group_power)."""
    nb, ptrs = jacobi_find_blocks(rp, ci, max_block_size)
    bo, go, gp = jacobi_storage_scheme(max_block_size)
    storage = -(-nb // (1 << gp)) * go
    blocks = np.zeros(storage, dtype=va.dtype)
    getattr(lib(), f"oracle_jacobi_block_generate_{_v(va.dtype)}")(P(rp), P(ci), P(va), i64(nb), P(ptrs), i64(bo), i64(go),
                                                                   int(gp), P(blocks))
    return dict(num_blocks=nb, block_ptrs=ptrs, blocks=blocks, block_offset=bo, group_offset=go, group_power=gp)


def jacobi_block_apply(J, b, alpha=None, beta=None, x=None):
    b2, k = _b2(b)
    out = np.zeros_like(b2) if x is None else np.ascontiguousarray(x.reshape(len(b2), -1)).copy()
    dt = b2.dtype
    getattr(lib(), f"oracle_jacobi_block_apply_{_v(dt)}")(
        i64(J["num_blocks"]), P(J["block_ptrs"]), P(J["blocks"]), i64(J["block_offset"]), i64(J["group_offset"]),
        int(J["group_power"]), i64(k), int(alpha is not None), _c(dt, alpha or 0), P(b2), i64(k), _c(dt, beta or 0),
        P(out), i64(k))
    return out


def _precond_struct(dtype, precond, inv_diag, J):
    VT = f64 if np.dtype(dtype) == np.float64 else f32

    class Precond(C.Structure):
        _fields_ = [("kind", C.c_int), ("inv_diag", vp), ("num_blocks", i64), ("bptrs", vp), ("blocks", vp),
                    ("block_offset", i64), ("group_offset", i64), ("group_power", C.c_int)]
    p = Precond()
    p.kind = precond
    if precond == 1:
        p.inv_diag = inv_diag.ctypes.data
    elif precond == 2:
        p.num_blocks, p.bptrs, p.blocks = J["num_blocks"], J["block_ptrs"].ctypes.data, J["blocks"].ctypes.data
        p.block_offset, p.group_offset, p.group_power = J["block_offset"], J["group_offset"], J["group_power"]
    return p


def krylov_solve(solver, rp, ci, va, b, x0, precond=0, inv_diag=None, J=None, max_iters=1000, factor=1e-8,
                 baseline=0, krylov_dim=30):
    """solver in {"bicgstab", "gmres"}; precond 0 none / 1 scalar Jacobi (inv_diag) / 2 block Jacobi (J from
    jacobi_block_generate).  Returns (x, iterations, residual_history, stop_status)."""
    n = len(rp) - 1
    b2 = np.ascontiguousarray(b.reshape(n, -1))
    k = b2.shape[1]
    x = np.ascontiguousarray(x0.reshape(n, -1)).copy()
    V = _v(va.dtype)
    hist = np.zeros(max_iters + 2, dtype=va.dtype)
    stop = np.zeros(k, dtype=np.uint8)
    p = _precond_struct(va.dtype, precond, inv_diag, J)
    if solver == "bicgstab":
        fn = getattr(lib(), f"oracle_bicgstab_solve_csr_i32_{V}")
        fn.restype = i64
        it = fn(i64(n), P(rp), P(ci), P(va), C.byref(p), i64(max_iters), _c(va.dtype, factor), int(baseline), i64(k),
                P(b2), P(x), P(hist), i64(len(hist)), P(stop))
    else:
        fn = getattr(lib(), f"oracle_gmres_solve_csr_i32_{V}")
        fn.restype = i64
        it = fn(i64(n), P(rp), P(ci), P(va), C.byref(p), i64(max_iters), _c(va.dtype, factor), int(baseline),
                i64(krylov_dim), i64(k), P(b2), P(x), P(hist), i64(len(hist)), P(stop))
    return x.reshape(x0.shape), int(it), hist[: it + 1].astype(np.float64), stop


def ref_jacobi_generate(rp, ci, va, max_block_size):
    r = ref()
    n = len(rp) - 1
    cap = (n + 64) * 32 * 2 + 4096
    meta = np.zeros(4, dtype=np.int64)
    ptrs = np.zeros(n + 2, dtype=np.int32)
    blocks = np.zeros(cap, dtype=va.dtype)
    fn = getattr(r, f"ref_jacobi_generate_{_v(va.dtype)}_i32")
    fn.restype = i64
    stored = fn(i64(n), i64(len(ci)), P(rp), P(ci), P(va), int(max_block_size), P(meta), P(ptrs), P(blocks), i64(cap))
    if stored < 0:
        raise RuntimeError(f"ref_jacobi_generate rc={stored}")
    nb = int(meta[0])
    return dict(num_blocks=nb, block_ptrs=ptrs[: nb + 1].copy(), blocks=blocks[:stored].copy(),
                block_offset=int(meta[1]), group_offset=int(meta[2]), group_power=int(meta[3]))
