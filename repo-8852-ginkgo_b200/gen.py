"""Synthetic BASELINE matrices, generated on the host by the C generators of
csrc/generators.cu (the same arrays feed the oracle and the GPU path)."""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import lib
from .core import check

KINDS = {"5pt": 0, "7pt": 1, "27pt": 2}


def stencil_csr(kind, nx, ny, nz=1, row_begin=0, row_end=None, value_dtype=np.float64, index_dtype=np.int32):
    """CSR arrays (row_ptrs, col_idxs, values) of rows [row_begin,row_end) of the stencil
    matrix; columns are GLOBAL indices."""
    k = KINDS[kind]
    n = nx * ny * (1 if k == 0 else nz)
    row_end = n if row_end is None else row_end
    nnz = lib.gkob200_gen_stencil_nnz(k, nx, ny, nz, row_begin, row_end)
    if nnz < 0:
        raise ValueError("bad stencil arguments")
    V = "f64" if value_dtype == np.float64 else "f32"
    P = "i32" if index_dtype == np.int32 else "i64"
    if P == "i32" and (nnz > 2 ** 31 - 1 or n > 2 ** 31 - 1):
        raise ValueError("int32 overflow; use index_dtype=np.int64")
    rp = np.empty(row_end - row_begin + 1, dtype=index_dtype)
    ci = np.empty(nnz, dtype=index_dtype)
    va = np.empty(nnz, dtype=value_dtype)
    fn = getattr(lib, f"gkob200_gen_stencil_csr_{V}_{P}_{P}")
    check(fn(k, nx, ny, nz, row_begin, row_end, C.c_void_p(rp.ctypes.data), C.c_void_p(ci.ctypes.data),
             C.c_void_p(va.ctypes.data)), "gen_stencil")
    return rp, ci, va, n


def powerlaw_csr(n, seed=42, lmin=3.0, alpha=2.5, lmax=100000):
    """Config-3 matrix: skewed row lengths, strictly diagonally dominant."""
    rp64 = np.empty(n + 1, dtype=np.int64)
    nnz = lib.gkob200_gen_powerlaw_row_ptrs_i64(n, seed, lmin, alpha, lmax, C.c_void_p(rp64.ctypes.data))
    if nnz < 0:
        raise ValueError("bad power-law arguments")
    rp = np.empty(n + 1, dtype=np.int32)
    ci = np.empty(nnz, dtype=np.int32)
    va = np.empty(nnz, dtype=np.float64)
    check(lib.gkob200_gen_powerlaw_fill_f64_i32(n, seed, C.c_void_p(rp64.ctypes.data), C.c_void_p(rp.ctypes.data),
                                                C.c_void_p(ci.ctypes.data), C.c_void_p(va.ctypes.data)),
          "gen_powerlaw")
    return rp, ci, va
