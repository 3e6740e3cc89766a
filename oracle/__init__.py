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


def gen_stencil_csr(kind, nx, ny, nz=1, row_begin=0, row_end=None):
    """(row_ptrs int32, col_idxs int32, values f64, n): the BASELINE stencil matrices from the
    oracle's own generator (independent of the product's csrc/generators.cu)."""
    k = {"5pt": 0, "7pt": 1, "27pt": 2}[kind]
    n = nx * ny * (1 if k == 0 else nz)
    row_end = n if row_end is None else row_end
    fn = lib().oracle_gen_stencil_csr
    fn.restype = i64
    nnz = fn(k, i64(nx), i64(ny), i64(nz), i64(row_begin), i64(row_end), None, None, None)
    if nnz > 2 ** 31 - 1:
        raise ValueError("int32 overflow")
    rp = np.empty(row_end - row_begin + 1, np.int32)
    ci = np.empty(nnz, np.int32)
    va = np.empty(nnz, np.float64)
    fn(k, i64(nx), i64(ny), i64(nz), i64(row_begin), i64(row_end), P(rp), P(ci), P(va))
    return rp, ci, va, n


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
SOLVERS = {"cg": 0, "bicgstab": 1, "gmres": 2, "fcg": 3, "cgs": 4}


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
    """solver in {"bicgstab", "gmres", "fcg", "cgs"}; precond 0 none / 1 scalar Jacobi (inv_diag) / 2 block Jacobi (J from
    jacobi_block_generate).  Returns (x, iterations, residual_history, stop_status)."""
    n = len(rp) - 1
    b2 = np.ascontiguousarray(b.reshape(n, -1))
    k = b2.shape[1]
    x = np.ascontiguousarray(x0.reshape(n, -1)).copy()
    V = _v(va.dtype)
    hist = np.zeros(max_iters + 2, dtype=va.dtype)
    stop = np.zeros(k, dtype=np.uint8)
    p = _precond_struct(va.dtype, precond, inv_diag, J)
    if solver in ("bicgstab", "fcg", "cgs"):
        fn = getattr(lib(), f"oracle_{solver}_solve_csr_i32_{V}")
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


# ---- distributed: partitions, build_local_nonlocal, apply ---------------------
class OPartition(C.Structure):
    _fields_ = [("num_ranges", i64), ("bounds", vp), ("part_ids", vp), ("starts", vp)]


class Partition:
    """distributed::Partition<int32,int64> restated (core/distributed/partition.cpp:60-140)."""

    def __init__(self, num_parts, bounds, part_ids):
        self.num_parts = num_parts
        self.bounds = np.ascontiguousarray(bounds, dtype=np.int64)
        self.part_ids = np.ascontiguousarray(part_ids, dtype=np.int32)
        nr = len(self.part_ids)
        self.starts = np.zeros(nr, dtype=np.int32)
        self.sizes = np.zeros(num_parts, dtype=np.int32)
        fn = lib().oracle_partition_build_starting_indices
        fn.restype = C.c_int32
        self.num_empty_parts = fn(P(self.bounds), P(self.part_ids), i64(nr), C.c_int32(num_parts), P(self.starts),
                                  P(self.sizes))
        self.size = int(self.bounds[-1])

    @classmethod
    def uniform(cls, num_parts, global_size):
        ranges = np.zeros(num_parts + 1, dtype=np.int64)
        lib().oracle_partition_build_ranges_from_global_size(C.c_int32(num_parts), i64(global_size), P(ranges))
        return cls(num_parts, ranges, np.arange(num_parts, dtype=np.int32))

    @classmethod
    def from_mapping(cls, mapping, num_parts):
        mapping = np.ascontiguousarray(mapping, dtype=np.int32)
        n = len(mapping)
        bounds, ids = np.zeros(n + 1, dtype=np.int64), np.zeros(max(n, 1), dtype=np.int32)
        fn = lib().oracle_partition_build_from_mapping
        fn.restype = i64
        nr = fn(i64(n), P(mapping), P(bounds), P(ids))
        return cls(num_parts, bounds[: nr + 1], ids[:nr])

    def struct(self):
        s = OPartition()
        s.num_ranges = len(self.part_ids)
        s.bounds, s.part_ids, s.starts = self.bounds.ctypes.data, self.part_ids.ctypes.data, self.starts.ctypes.data
        return s


def dist_build_local_nonlocal(rows, cols, vals, part, local_part):
    nnz = len(rows)
    rows, cols = np.ascontiguousarray(rows, np.int64), np.ascontiguousarray(cols, np.int64)
    cap = max(nnz, 1)
    o = dict(lrow=np.zeros(cap, np.int32), lcol=np.zeros(cap, np.int32), lval=np.zeros(cap, vals.dtype),
             nrow=np.zeros(cap, np.int32), ncol=np.zeros(cap, np.int32), nval=np.zeros(cap, vals.dtype),
             gather=np.zeros(cap, np.int32), recv_sizes=np.zeros(part.num_parts, np.int32),
             nl_to_global=np.zeros(cap, np.int64))
    counts = np.zeros(3, np.int64)
    ps = part.struct()
    getattr(lib(), f"oracle_dist_build_local_nonlocal_{_v(vals.dtype)}")(
        i64(nnz), P(rows), P(cols), P(vals), C.byref(ps), C.byref(ps), C.c_int32(part.num_parts), C.c_int32(local_part),
        P(o["lrow"]), P(o["lcol"]), P(o["lval"]), P(o["nrow"]), P(o["ncol"]), P(o["nval"]), P(o["gather"]),
        P(o["recv_sizes"]), P(o["nl_to_global"]), P(counts))
    nl, nn, ng = (int(c) for c in counts)
    for k_ in ("lrow", "lcol", "lval"):
        o[k_] = o[k_][:nl].copy()
    for k_ in ("nrow", "ncol", "nval"):
        o[k_] = o[k_][:nn].copy()
    o["gather"], o["nl_to_global"] = o["gather"][:ng].copy(), o["nl_to_global"][:ng].copy()
    return o


def ref_dist_build(rows, cols, vals, local_part, num_parts, global_size, mapping=None):
    """The real reference kernels (Partition + build_local_nonlocal); returns (parts dict, partition meta)."""
    r = ref()
    nnz = len(rows)
    cap = max(nnz, 1)
    o = dict(lrow=np.zeros(cap, np.int32), lcol=np.zeros(cap, np.int32), lval=np.zeros(cap, np.float64),
             nrow=np.zeros(cap, np.int32), ncol=np.zeros(cap, np.int32), nval=np.zeros(cap, np.float64),
             gather=np.zeros(cap, np.int32), recv_sizes=np.zeros(num_parts, np.int32),
             nl_to_global=np.zeros(cap, np.int64))
    counts = np.zeros(3, np.int64)
    meta = np.zeros(4 * (global_size + 2) + num_parts + 8, np.int64)
    mode = 0 if mapping is None else 2
    mp = None if mapping is None else np.ascontiguousarray(mapping, np.int32)
    rc = r.ref_dist_build_f64(mode, num_parts, i64(global_size), None, P(mp), i64(nnz),
                              P(np.ascontiguousarray(rows, np.int64)), P(np.ascontiguousarray(cols, np.int64)),
                              P(np.ascontiguousarray(vals, np.float64)), int(local_part), P(counts), P(o["lrow"]),
                              P(o["lcol"]), P(o["lval"]), P(o["nrow"]), P(o["ncol"]), P(o["nval"]), P(o["gather"]),
                              P(o["recv_sizes"]), P(o["nl_to_global"]), P(meta))
    if rc != 0:
        raise RuntimeError(f"ref_dist_build rc={rc}")
    nl, nn, ng = (int(c) for c in counts)
    for k_ in ("lrow", "lcol", "lval"):
        o[k_] = o[k_][:nl].copy()
    for k_ in ("nrow", "ncol", "nval"):
        o[k_] = o[k_][:nn].copy()
    o["gather"], o["nl_to_global"] = o["gather"][:ng].copy(), o["nl_to_global"][:ng].copy()
    nr = int(meta[0])
    pm = dict(bounds=meta[1: nr + 2].copy(), part_ids=meta[nr + 2: 2 * nr + 2].astype(np.int32),
              starts=meta[2 * nr + 2: 3 * nr + 2].astype(np.int32),
              sizes=meta[3 * nr + 2: 3 * nr + 2 + num_parts].astype(np.int32))
    return o, pm


def coo_to_csr(n_rows, rows, cols, vals):
    rp = np.zeros(n_rows + 1, dtype=np.int32)
    lib().oracle_convert_idxs_to_ptrs_i32(P(np.ascontiguousarray(rows, np.int32)), i64(len(rows)), i64(n_rows), P(rp))
    return rp, np.ascontiguousarray(cols, np.int32), np.ascontiguousarray(vals)


def dist_plan(parts_per_rank):
    """The exchange of distributed::Matrix::read_distributed (core/distributed/matrix.cpp:197-221)
    for all ranks at once: send_sizes[p][q] = recv_sizes[q][p]; gather_idxs of rank p = the
    gather lists its receivers q computed, in rank order."""
    Pn = len(parts_per_rank)
    recv = np.array([pp["recv_sizes"] for pp in parts_per_rank], dtype=np.int64)
    send = recv.T.copy()
    gathers = []
    for p in range(Pn):
        chunks = []
        for q in range(Pn):
            off = int(recv[q, :p].sum())
            chunks.append(parts_per_rank[q]["gather"][off: off + int(recv[q, p])])
        gathers.append(np.concatenate(chunks) if chunks else np.zeros(0, np.int32))
    return send, recv, gathers


def dist_apply(parts_per_rank, part, b_global, alpha=None, beta=None, x_global=None):
    """distributed::Matrix::apply restated for all ranks in one process
    (core/distributed/matrix.cpp:263-369): pack with row_gather, exchange, local apply, then
    x += A_nl * recv.  Returns the global result vector (rows in partition order)."""
    Pn = len(parts_per_rank)
    send, recv, gathers = dist_plan(parts_per_rank)
    b_loc = [global_to_local(part, b_global, p) for p in range(Pn)]
    send_bufs = [b_loc[p][gathers[p]] for p in range(Pn)]
    out = np.zeros_like(b_global) if x_global is None else x_global.copy()
    for p in range(Pn):
        pp = parts_per_rank[p]
        n_loc = int(part.sizes[p])
        ghosts = []
        for q in range(Pn):  # what q sends to p sits at q's send offset for p
            off = int(send[q, :p].sum())
            ghosts.append(send_bufs[q][off: off + int(send[q, p])])
        ghost = np.concatenate(ghosts) if ghosts else np.zeros((0,) + b_global.shape[1:], b_global.dtype)
        lrp, lci, lva = coo_to_csr(n_loc, pp["lrow"], pp["lcol"], pp["lval"])
        nrp, nci, nva = coo_to_csr(n_loc, pp["nrow"], pp["ncol"], pp["nval"])
        x_loc = global_to_local(part, out, p)
        if alpha is None:
            x_loc = csr_spmv(lrp, lci, lva, b_loc[p])
            if len(ghost):
                x_loc = csr_spmv(nrp, nci, nva, ghost, 1.0, 1.0, x_loc)
        else:
            x_loc = csr_spmv(lrp, lci, lva, b_loc[p], alpha, beta, x_loc)
            if len(ghost):
                x_loc = csr_spmv(nrp, nci, nva, ghost, alpha, 1.0, x_loc)
        local_to_global(part, out, p, x_loc)
    return out


def _local_rows(part, p):
    idx = []
    for r in range(len(part.part_ids)):
        if part.part_ids[r] == p:
            idx.append(np.arange(part.bounds[r], part.bounds[r + 1]))
    return np.concatenate(idx) if idx else np.zeros(0, np.int64)


def global_to_local(part, v, p):
    return np.ascontiguousarray(v[_local_rows(part, p)])


def local_to_global(part, v, p, loc):
    v[_local_rows(part, p)] = loc


# ---- matrix assembly -------------------------------------------
SORT, SUM_DUPLICATES, REMOVE_ZEROS = 1, 2, 4


def coo_assemble(rows, cols, vals, steps=SORT | SUM_DUPLICATES | REMOVE_ZEROS):
    """device_matrix_data::{sort_row_major, sum_duplicates, remove_zeros} (plain-C oracle);
    returns new (rows, cols, vals)."""
    r, c, v = rows.copy(), cols.copy(), vals.copy()
    V, I = _v(v.dtype), _i(r.dtype)
    n = len(v)
    L = lib()
    if steps & SORT:
        getattr(L, f"oracle_coo_sort_row_major_{I}_{V}")(i64(n), P(r), P(c), P(v))
    if steps & SUM_DUPLICATES:
        fn = getattr(L, f"oracle_coo_sum_duplicates_{I}_{V}")
        fn.restype = i64
        n = fn(i64(n), P(r), P(c), P(v))
    if steps & REMOVE_ZEROS:
        fn = getattr(L, f"oracle_coo_remove_zeros_{I}_{V}")
        fn.restype = i64
        n = fn(i64(n), P(r), P(c), P(v))
    return r[:n].copy(), c[:n].copy(), v[:n].copy()


def csr_transpose(n_rows, n_cols, rp, ci, va):
    out_rp = np.zeros(n_cols + 1, dtype=rp.dtype)
    out_ci, out_va = np.zeros_like(ci), np.zeros_like(va)
    getattr(lib(), f"oracle_csr_transpose_{_i(rp.dtype)}_{_v(va.dtype)}")(i64(n_rows), i64(n_cols), P(rp), P(ci), P(va),
                                                                          P(out_rp), P(out_ci), P(out_va))
    return out_rp, out_ci, out_va


def csr_sort_by_column_index(rp, ci, va):
    c, v = ci.copy(), va.copy()
    getattr(lib(), f"oracle_csr_sort_by_column_index_{_i(rp.dtype)}_{_v(va.dtype)}")(i64(len(rp) - 1), P(rp), P(c), P(v))
    return c, v


def ref_assemble(n_rows, n_cols, rows, cols, vals, steps=SORT | SUM_DUPLICATES | REMOVE_ZEROS):
    """The same steps on the real reference executor; also returns the row_ptrs Csr::read builds."""
    r, c, v = rows.copy(), cols.copy(), vals.copy()
    rp = np.zeros(n_rows + 1, dtype=r.dtype)
    fn = getattr(ref(), f"ref_assemble_{_v(v.dtype)}_{_i(r.dtype)}")
    fn.restype = i64
    n = fn(int(steps), i64(n_rows), i64(n_cols), i64(len(v)), P(r), P(c), P(v), P(rp))
    return r[:n].copy(), c[:n].copy(), v[:n].copy(), rp


def ref_csr_op(op, n_rows, n_cols, rp, ci, va):
    """op 0: transpose, 1: sort_by_column_index on the real reference executor."""
    out_rp = np.zeros((n_cols if op == 0 else n_rows) + 1, dtype=rp.dtype)
    out_ci, out_va = np.zeros_like(ci), np.zeros_like(va)
    getattr(ref(), f"ref_csr_op_{_v(va.dtype)}_{_i(rp.dtype)}")(int(op), i64(n_rows), i64(n_cols), i64(len(ci)), P(rp),
                                                                P(ci), P(va), P(out_rp), P(out_ci), P(out_va))
    return out_rp, out_ci, out_va


# ---- files: the reference's own reader / writer (gko::read_generic_raw, write_raw, write_binary_raw)
def ref_mtx_read(path, cap=1 << 22):
    """Returns ((n_rows, n_cols), rows, cols, vals) or raises RuntimeError when the reference throws."""
    dims = np.zeros(2, dtype=np.int64)
    rows, cols = np.zeros(cap, dtype=np.int64), np.zeros(cap, dtype=np.int64)
    vals = np.zeros(cap)
    fn = ref().ref_mtx_read
    fn.restype = i64
    n = fn(str(path).encode(), P(dims), i64(cap), P(rows), P(cols), P(vals))
    if n < 0:
        raise RuntimeError(f"reference reader failed ({n})")
    return (int(dims[0]), int(dims[1])), rows[:n].copy(), cols[:n].copy(), vals[:n].copy()


def ref_mtx_write(path, size, rows, cols, vals, binary=False, index32=False, value32=False, array=False):
    r, c = np.ascontiguousarray(rows, dtype=np.int64), np.ascontiguousarray(cols, dtype=np.int64)
    v = np.ascontiguousarray(vals, dtype=np.float64)
    rc = ref().ref_mtx_write(str(path).encode(), 2 if array else int(binary), i64(size[0]), i64(size[1]), i64(len(v)), P(r), P(c), P(v),
                             int(index32), int(value32))
    if rc:
        raise RuntimeError(f"reference writer failed ({rc})")
