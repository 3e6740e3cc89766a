"""gko::read / gko::write for matrices (reference include/ginkgo/core/base/mtx_io.hpp:
read_raw, read_binary_raw, read_generic_raw, write_raw, write_binary_raw) over the native
reader of the C-ABI.  Host arrays in, host arrays out; `read` hands them to the device assembly."""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import lib
from .core import Error

COORDINATE, BINARY, ARRAY = 0, 1, 2


def _check(rc, what):
    if rc != 0:
        msg = lib.gkob200_mtx_last_error()
        raise Error(f"{what}: {msg.decode() if msg else ''}", rc)


def read_raw(path, value_dtype=np.float64, index_dtype=np.int32):
    """read_generic_raw: returns ((n_rows, n_cols), rows, cols, vals), row-major sorted."""
    h, n, m, nnz = C.c_void_p(), C.c_int64(), C.c_int64(), C.c_int64()
    _check(lib.gkob200_mtx_read_open(str(path).encode(), C.byref(h), C.byref(n), C.byref(m), C.byref(nnz)), "mtx read")
    try:
        rows = np.empty(nnz.value, dtype=index_dtype)
        cols = np.empty(nnz.value, dtype=index_dtype)
        vals = np.empty(nnz.value, dtype=value_dtype)
        V = "f64" if np.dtype(value_dtype) == np.float64 else "f32"
        I = "i32" if np.dtype(index_dtype) == np.int32 else "i64"  # noqa: E741
        P = lambda a: C.c_void_p(a.ctypes.data)  # noqa: E731
        _check(getattr(lib, f"gkob200_mtx_read_copy_{V}_{I}")(h, P(rows), P(cols), P(vals)), "mtx read")
    finally:
        lib.gkob200_mtx_read_close(h)
    return (n.value, m.value), rows, cols, vals


def write_raw(path, size, rows, cols, vals, layout=COORDINATE, precision=0):
    """write_raw (coordinate real general) / write_binary_raw."""
    rows, cols, vals = np.ascontiguousarray(rows), np.ascontiguousarray(cols), np.ascontiguousarray(vals)
    V = "f64" if vals.dtype == np.float64 else "f32"
    I = "i32" if rows.dtype == np.int32 else "i64"  # noqa: E741
    P = lambda a: C.c_void_p(a.ctypes.data)  # noqa: E731
    _check(getattr(lib, f"gkob200_mtx_write_{V}_{I}")(str(path).encode(), int(layout), int(precision), size[0], size[1],
                                                      len(vals), P(rows), P(cols), P(vals)), "mtx write")


def read(exec_, path, value_dtype=np.float64, index_dtype=np.int32, strategy="automatical"):
    """gko::read<Csr>(stream, exec): file -> triplets -> device assembly -> Csr."""
    from .assembly import DeviceMatrixData
    size, rows, cols, vals = read_raw(path, value_dtype, index_dtype)
    return DeviceMatrixData.from_arrays(exec_, size, rows, cols, vals).to_csr(strategy)
