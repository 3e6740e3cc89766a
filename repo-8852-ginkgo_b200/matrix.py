"""Dense and sparse LinOps of the host mirror (reference include/ginkgo/core/matrix/*.hpp)."""
from __future__ import annotations

import ctypes as C

import numpy as np
import torch

from . import _abi, lib
from .core import Error, check, current_stream, iname, ptr, vname, _TORCH_V


class Dense:
    """matrix::Dense<V>: n x k row-major values with a stride
    (reference include/ginkgo/core/matrix/dense.hpp)."""

    def __init__(self, exec_, tensor):
        assert tensor.dim() == 2 and (tensor.shape[1] <= 1 or tensor.stride(1) == 1)
        self.exec = exec_
        self.t = tensor

    # -- construction ------------------------------------------------------
    @classmethod
    def create(cls, exec_, size, dtype=torch.float64, stride=None):
        n, k = size
        stride = k if stride is None else stride
        buf = torch.zeros((n, stride), dtype=dtype, device=exec_.device)
        return cls(exec_, buf[:, :k])

    @classmethod
    def from_numpy(cls, exec_, arr, stride=None):
        arr = np.asarray(arr)
        if arr.ndim == 1:
            arr = arr[:, None]
        d = cls.create(exec_, arr.shape, dtype=torch.from_numpy(arr[:0]).dtype, stride=stride)
        d.t.copy_(torch.from_numpy(np.ascontiguousarray(arr)))
        return d

    @classmethod
    def scalar(cls, exec_, value, dtype=torch.float64):
        return cls.from_numpy(exec_, np.array([[value]], dtype=np.float64 if dtype == torch.float64 else np.float32))

    # -- accessors ---------------------------------------------------------
    @property
    def size(self):
        return tuple(self.t.shape)

    @property
    def stride(self):
        return self.t.stride(0) if self.t.shape[0] > 1 or self.t.shape[1] > 0 else max(self.t.shape[1], 1)

    @property
    def V(self):
        return vname(self.t.dtype)

    def to_numpy(self):
        return self.t.detach().cpu().numpy().copy()

    def clone(self):
        d = Dense.create(self.exec, self.size, self.t.dtype)
        d.copy_from(self)
        return d

    def _args(self):
        return self.size[0], self.size[1], ptr(self.t), self.stride

    def _fn(self, name):
        return getattr(lib, f"gkob200_dense_{name}_{self.V}"), f"dense::{name}"

    # -- BLAS-1 (reference core/matrix/dense.cpp scale/add_scaled/...) --------
    def fill(self, value):
        fn, nm = self._fn("fill")
        n, k, p, s = self._args()
        check(fn(current_stream(), n, k, p, s, value), nm)

    def copy_from(self, other):
        fn, nm = self._fn("copy")
        n, k, p, s = self._args()
        check(fn(current_stream(), n, k, ptr(other.t), other.stride, p, s), nm)

    def scale(self, alpha):
        fn, nm = self._fn("scale")
        n, k, p, s = self._args()
        check(fn(current_stream(), n, k, ptr(alpha.t), alpha.size[1], p, s), nm)

    def inv_scale(self, alpha):
        fn, nm = self._fn("inv_scale")
        n, k, p, s = self._args()
        check(fn(current_stream(), n, k, ptr(alpha.t), alpha.size[1], p, s), nm)

    def add_scaled(self, alpha, x):
        fn, nm = self._fn("add_scaled")
        n, k, p, s = self._args()
        check(fn(current_stream(), n, k, ptr(alpha.t), alpha.size[1], ptr(x.t), x.stride, p, s), nm)

    def sub_scaled(self, alpha, x):
        fn, nm = self._fn("sub_scaled")
        n, k, p, s = self._args()
        check(fn(current_stream(), n, k, ptr(alpha.t), alpha.size[1], ptr(x.t), x.stride, p, s), nm)

    def compute_dot(self, other, result):
        fn, nm = self._fn("compute_dot")
        n, k, p, s = self._args()
        check(fn(current_stream(), n, k, p, s, ptr(other.t), other.stride, ptr(result.t), self.exec.ws), nm)

    compute_conj_dot = compute_dot  # real value types

    def _norm(self, kind, result):
        fn, nm = self._fn(kind)
        n, k, p, s = self._args()
        check(fn(current_stream(), n, k, p, s, ptr(result.t), self.exec.ws), nm)

    def compute_norm2(self, result):
        self._norm("compute_norm2", result)

    def compute_squared_norm2(self, result):
        self._norm("compute_squared_norm2", result)

    def compute_norm1(self, result):
        self._norm("compute_norm1", result)

    def row_gather(self, rows, out):
        """out(i,:) = self(rows[i],:)  (reference Dense::row_gather)."""
        fn = getattr(lib, f"gkob200_dense_row_gather_{self.V}_{iname(rows.dtype)}")
        check(fn(current_stream(), out.size[0], out.size[1], ptr(rows), ptr(self.t), self.stride, ptr(out.t),
                 out.stride), "dense::row_gather")


class _SparseBase:
    """Common LinOp behaviour: apply(b, x) and apply(alpha, b, beta, x)
    (reference include/ginkgo/core/base/lin_op.hpp:158-226, dimension checks :323-346)."""

    def descriptor(self):
        raise NotImplementedError

    def apply(self, *args):
        if len(args) == 2:
            b, x = args
            alpha = beta = None
        elif len(args) == 4:
            alpha, b, beta, x = args
        else:
            raise TypeError("apply(b, x) or apply(alpha, b, beta, x)")
        n, m = self.size
        if b.size[0] != m or x.size[0] != n or b.size[1] != x.size[1]:
            raise Error("LinOp::apply DimensionMismatch", -1)
        d = self.descriptor()
        check(lib.gkob200_matrix_apply(current_stream(), C.byref(d), ptr(b.t), b.stride, b.size[1],
                                       ptr(alpha.t) if alpha is not None else None,
                                       ptr(beta.t) if beta is not None else None, ptr(x.t), x.stride),
              f"{type(self).__name__}::apply")
        return x


_STRATEGIES = {"classical": _abi.CSR_CLASSICAL, "merge_path": _abi.CSR_MERGE_PATH,
               "load_balance": _abi.CSR_MERGE_PATH, "automatical": _abi.CSR_AUTO,
               "sparselib": _abi.CSR_AUTO, "cusparse": _abi.CSR_AUTO}


class Csr(_SparseBase):
    """matrix::Csr<V,I> (reference include/ginkgo/core/matrix/csr.hpp).  The strategy
    names are the reference's (csr.hpp:178-700); `automatical` picks the kernel from
    row-length statistics gathered once on the device when the matrix is created —
    the counterpart of the reference's make_srow() (csr.hpp:1263-1267)."""

    def __init__(self, exec_, size, row_ptrs, col_idxs, values, strategy="automatical"):
        self.exec = exec_
        self.size = tuple(size)
        self.row_ptrs, self.col_idxs, self.values = row_ptrs, col_idxs, values
        if strategy not in _STRATEGIES:
            raise Error(f"Csr strategy {strategy}", -1)
        self.strategy = strategy
        self._stats = None
        self._ws = None
        self._desc = None
        self.make_srow()

    @classmethod
    def from_arrays(cls, exec_, size, row_ptrs, col_idxs, values, strategy="automatical"):
        """Host numpy arrays (or tensors) -> device CSR."""
        def dev(a):
            if isinstance(a, torch.Tensor):
                return a.to(exec_.device)
            return torch.from_numpy(np.ascontiguousarray(a)).to(exec_.device)
        return cls(exec_, size, dev(row_ptrs), dev(col_idxs), dev(values), strategy)

    @classmethod
    def from_scipy(cls, exec_, m, strategy="automatical", index_dtype=np.int32):
        m = m.tocsr()
        m.sort_indices()
        return cls.from_arrays(exec_, m.shape, m.indptr.astype(index_dtype), m.indices.astype(index_dtype),
                               m.data, strategy)

    @property
    def nnz(self):
        return int(self.col_idxs.numel())

    @property
    def V(self):
        return vname(self.values.dtype)

    @property
    def I(self):  # noqa: E743
        return iname(self.row_ptrs.dtype)

    def make_srow(self):
        stats = torch.zeros(4, dtype=torch.int64, device=self.exec.device)
        fn = getattr(lib, f"gkob200_csr_row_stats_{self.I}")
        check(fn(current_stream(), self.size[0], ptr(self.row_ptrs), ptr(stats)), "csr::row_stats")
        self._stats = [int(v) for v in stats.cpu()]
        self._desc = None

    @property
    def max_row_nnz(self):
        return self._stats[0]

    @property
    def max_block_nnz(self):
        return self._stats[1]

    def kernel(self):
        """Which kernel apply() will run: 'classical' (row-block) or 'merge_path'."""
        s = _STRATEGIES[self.strategy]
        if s == _abi.CSR_AUTO:
            s = lib.gkob200_csr_pick_strategy(self.size[0], self.nnz, self.max_row_nnz, self.max_block_nnz)
        return "classical" if s == _abi.CSR_CLASSICAL else "merge_path"

    def descriptor(self):
        if self._desc is not None:
            return self._desc
        d = _abi.Matrix()
        d.format = _abi.FMT_CSR
        d.value_type = _abi.F64 if self.V == "f64" else _abi.F32
        d.index_type = _abi.I32 if self.I == "i32" else _abi.I64
        d.csr_strategy = _abi.CSR_CLASSICAL if self.kernel() == "classical" else _abi.CSR_MERGE_PATH
        d.n_rows, d.n_cols = self.size
        d.nnz = self.nnz
        d.row_ptrs, d.col_idxs, d.values = (self.row_ptrs.data_ptr(), self.col_idxs.data_ptr(),
                                            self.values.data_ptr())
        d.csr_max_block_nnz = self.max_block_nnz
        if d.csr_strategy == _abi.CSR_MERGE_PATH:
            nbytes = lib.gkob200_csr_spmv_workspace_bytes(self.size[0], self.nnz, 1, self.values.element_size())
            self._ws = torch.empty(nbytes, dtype=torch.uint8, device=self.exec.device)
            d.workspace, d.workspace_bytes = self._ws.data_ptr(), nbytes
        self._desc = d
        return d

    def extract_diagonal(self):
        n = min(self.size)
        diag = torch.zeros(n, dtype=self.values.dtype, device=self.exec.device)
        if self.I != "i32":
            raise Error("extract_diagonal index type", -2)
        fn = getattr(lib, f"gkob200_csr_extract_diagonal_{self.V}_i32")
        check(fn(current_stream(), self.size[0], self.size[1], ptr(self.row_ptrs), ptr(self.col_idxs),
                 ptr(self.values), ptr(diag)), "csr::extract_diagonal")
        return diag

    def spmv_bytes(self, nrhs=1):
        """Algorithmic bytes of one apply."""
        v, i = self.values.element_size(), self.row_ptrs.element_size()
        n, m = self.size
        return self.nnz * (v + i) + (n + 1) * i + m * nrhs * v + n * nrhs * v
