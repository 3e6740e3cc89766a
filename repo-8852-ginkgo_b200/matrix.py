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
            # plan once per matrix (the role of the srow array of Csr::make_srow())
            nbytes = lib.gkob200_csr_spmv_workspace_bytes(self.size[0], self.nnz, 1, 8)
            self._ws = torch.empty(nbytes, dtype=torch.uint8, device=self.exec.device)
            d.workspace, d.workspace_bytes = self._ws.data_ptr(), nbytes
            check(getattr(lib, f"gkob200_csr_merge_plan_{self.I}")(current_stream(), self.size[0], self.nnz,
                                                                   ptr(self.row_ptrs), ptr(self._ws), nbytes),
                  "csr::merge_plan")
            d.csr_strategy = _abi.CSR_MERGE_PATH_PLANNED
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


# --------------------------------------------------------------------------- #
# integer helpers (device kernels; results bit-exact with the reference)
# --------------------------------------------------------------------------- #
def _scan_ws(exec_, n):
    nbytes = lib.gkob200_prefix_sum_workspace_bytes(n)
    return torch.empty(max(nbytes, 8), dtype=torch.uint8, device=exec_.device), nbytes


def prefix_sum(exec_, t):
    """In-place exclusive scan of a device tensor (int32/int64; uint64 stored as int64)."""
    ws, nb = _scan_ws(exec_, t.numel())
    name = {torch.int32: "i32", torch.int64: "i64"}[t.dtype]
    check(getattr(lib, f"gkob200_prefix_sum_{name}")(current_stream(), ptr(t), t.numel(), ptr(ws), nb), "prefix_sum")
    return t


def row_nnz_kth_smallest(exec_, row_ptrs, k):
    """k-th smallest row length (0-based) — what the reference obtains by std::sort of
    row_nnz on the host (hybrid.hpp:255-270) — from device histograms, no sort."""
    n = row_ptrs.numel() - 1
    I = iname(row_ptrs.dtype)
    mx = torch.zeros(1, dtype=torch.int64, device=exec_.device)
    check(getattr(lib, f"gkob200_compute_max_row_nnz_{I}")(current_stream(), ptr(row_ptrs), n, ptr(mx)), "max_row_nnz")
    lo, hi = 0, int(mx.item()) + 1  # answer in [lo, hi)
    bins = 4096
    hist = torch.zeros(bins, dtype=torch.int64, device=exec_.device)
    below = 0  # rows with length < lo
    while True:
        width = max(1, -(-(hi - lo) // bins))
        check(getattr(lib, f"gkob200_row_len_histogram_{I}")(current_stream(), ptr(row_ptrs), n, lo, width, bins,
                                                              ptr(hist)), "row_len_histogram")
        h = hist.cpu().numpy()
        cum = below
        for b in range(bins):
            if cum + h[b] > k:
                if width == 1:
                    return lo + b
                lo, hi, below = lo + b * width, lo + (b + 1) * width, cum
                break
            cum += h[b]
        else:
            raise Error("row_nnz_kth_smallest", -1)


class Ell(_SparseBase):
    """matrix::Ell<V,I>: col_idxs/values[stride * num_stored_elements_per_row], column-major,
    padding col = -1 / val = 0 (reference include/ginkgo/core/matrix/ell.hpp)."""

    def __init__(self, exec_, size, width, stride, col_idxs, values):
        self.exec, self.size, self.width, self.stride = exec_, tuple(size), int(width), int(stride)
        self.col_idxs, self.values = col_idxs, values
        self.V, self.I = vname(values.dtype), iname(col_idxs.dtype)

    def descriptor(self):
        d = _abi.Matrix()
        d.format = _abi.FMT_ELL
        d.value_type = _abi.F64 if self.V == "f64" else _abi.F32
        d.index_type = _abi.I32 if self.I == "i32" else _abi.I64
        d.n_rows, d.n_cols = self.size
        d.nnz = self.width * self.stride
        d.ell_stride, d.ell_width = self.stride, self.width
        d.ell_col_idxs, d.ell_values = self.col_idxs.data_ptr(), self.values.data_ptr()
        return d

    def kernel(self):
        return "ell"

    def spmv_bytes(self, nrhs=1):
        v, i = self.values.element_size(), self.col_idxs.element_size()
        return self.size[0] * self.width * (v + i) + (self.size[1] + self.size[0]) * nrhs * v


class Sellp(_SparseBase):
    """matrix::Sellp<V,I> (reference include/ginkgo/core/matrix/sellp.hpp): slice_size 64,
    stride_factor 1 by default; slice_sets/slice_lengths are size_type (uint64, kept in int64
    tensors)."""

    def __init__(self, exec_, size, slice_size, stride_factor, slice_sets, slice_lengths, col_idxs, values):
        self.exec, self.size = exec_, tuple(size)
        self.slice_size, self.stride_factor = int(slice_size), int(stride_factor)
        self.slice_sets, self.slice_lengths, self.col_idxs, self.values = slice_sets, slice_lengths, col_idxs, values
        self.V, self.I = vname(values.dtype), iname(col_idxs.dtype)
        # host-side facts the bulk-async kernel sizes its shared memory with
        self.max_slice_len = int(slice_lengths.max().item()) if slice_lengths.numel() else 0
        self.total_cols = int(slice_sets[-1].item()) if slice_sets.numel() else 0

    def descriptor(self):
        d = _abi.Matrix()
        d.format = _abi.FMT_SELLP
        d.value_type = _abi.F64 if self.V == "f64" else _abi.F32
        d.index_type = _abi.I32 if self.I == "i32" else _abi.I64
        d.n_rows, d.n_cols = self.size
        d.nnz = self.values.numel()
        d.col_idxs, d.values = self.col_idxs.data_ptr(), self.values.data_ptr()
        d.slice_size, d.stride_factor = self.slice_size, self.stride_factor
        d.n_slices = self.slice_lengths.numel()
        d.slice_sets, d.slice_lengths = self.slice_sets.data_ptr(), self.slice_lengths.data_ptr()
        d.sellp_max_slice_len, d.sellp_total_cols = self.max_slice_len, self.total_cols
        return d

    def kernel(self):
        return "sellp"

    def spmv_bytes(self, nrhs=1):
        v, i = self.values.element_size(), self.col_idxs.element_size()
        return (self.values.numel() * (v + i) + (self.slice_lengths.numel() + 1) * 8
                + (self.size[1] + self.size[0]) * nrhs * v)


class Coo(_SparseBase):
    """matrix::Coo<V,I>: row-sorted triplets (reference include/ginkgo/core/matrix/coo.hpp)."""

    def __init__(self, exec_, size, row_idxs, col_idxs, values):
        self.exec, self.size = exec_, tuple(size)
        self.row_idxs, self.col_idxs, self.values = row_idxs, col_idxs, values
        self.V, self.I = vname(values.dtype), iname(col_idxs.dtype)
        nb = lib.gkob200_coo_spmv_workspace_bytes(values.numel(), values.element_size())
        self._ws = torch.empty(nb, dtype=torch.uint8, device=exec_.device)

    def descriptor(self):
        d = _abi.Matrix()
        d.format = _abi.FMT_COO
        d.value_type = _abi.F64 if self.V == "f64" else _abi.F32
        d.index_type = _abi.I32 if self.I == "i32" else _abi.I64
        d.n_rows, d.n_cols = self.size
        d.nnz = self.values.numel()
        d.row_ptrs, d.col_idxs, d.values = self.row_idxs.data_ptr(), self.col_idxs.data_ptr(), self.values.data_ptr()
        d.workspace, d.workspace_bytes = self._ws.data_ptr(), self._ws.numel()
        return d

    def apply2(self, *args):
        """x += A b  /  x += alpha A b  (reference Coo::apply2, include/ginkgo/core/matrix/coo.hpp)."""
        if len(args) == 2:
            alpha, (b, x) = None, args
        else:
            alpha, b, x = args
        fn = getattr(lib, f"gkob200_coo_spmv2_{self.V}_{self.I}")
        check(fn(current_stream(), self.size[0], self.size[1], self.values.numel(), ptr(self.row_idxs),
                 ptr(self.col_idxs), ptr(self.values), ptr(b.t), b.stride, b.size[1],
                 ptr(alpha.t) if alpha is not None else None, ptr(x.t), x.stride, ptr(self._ws), self._ws.numel()),
              "coo::spmv2")
        return x

    def kernel(self):
        return "coo"

    def spmv_bytes(self, nrhs=1):
        v, i = self.values.element_size(), self.col_idxs.element_size()
        return self.values.numel() * (v + 2 * i) + self.size[1] * nrhs * v + 2 * self.size[0] * nrhs * v


class Hybrid(_SparseBase):
    """matrix::Hybrid<V,I> = ELL part + COO part (reference include/ginkgo/core/matrix/hybrid.hpp)."""

    def __init__(self, exec_, size, ell, coo):
        self.exec, self.size, self.ell, self.coo = exec_, tuple(size), ell, coo
        self.V, self.I = ell.V, ell.I

    def descriptor(self):
        d = self.ell.descriptor()
        c = self.coo.descriptor()
        d.format = _abi.FMT_HYBRID
        d.coo_nnz = c.nnz
        d.coo_row_idxs, d.coo_col_idxs, d.coo_values = c.row_ptrs, c.col_idxs, c.values
        d.workspace, d.workspace_bytes = c.workspace, c.workspace_bytes
        return d

    def kernel(self):
        return "hybrid"

    def spmv_bytes(self, nrhs=1):
        return self.ell.spmv_bytes(nrhs) + self.coo.spmv_bytes(nrhs)


class HybridStrategy:
    """Hybrid::strategy_type family (reference hybrid.hpp:112-380): the ELL width is an order
    statistic of the row lengths; computed here from device histograms (bit-exact)."""

    def __init__(self, kind, num_columns=0, percent=0.8, ratio=0.0001):
        self.kind, self.num_columns = kind, num_columns
        self.percent, self.ratio = min(max(percent, 0.0), 1.0), ratio

    @classmethod
    def column_limit(cls, n=0):
        return cls("column_limit", num_columns=n)

    @classmethod
    def imbalance_limit(cls, percent=0.8):
        return cls("imbalance_limit", percent=percent)

    @classmethod
    def imbalance_bounded_limit(cls, percent=0.8, ratio=0.0001):
        return cls("imbalance_bounded_limit", percent=percent, ratio=ratio)

    @classmethod
    def minimal_storage_limit(cls, value_bytes=8, index_bytes=4):
        return cls("imbalance_limit", percent=index_bytes / (value_bytes + 2 * index_bytes))

    @classmethod
    def automatic(cls):
        return cls("imbalance_bounded_limit", percent=1.0 / 3.0, ratio=0.001)

    def ell_width(self, exec_, row_ptrs):
        n = row_ptrs.numel() - 1
        if self.kind == "column_limit":
            return self.num_columns
        if n == 0:
            return 0
        pos = int(n * self.percent) if self.percent < 1 else n - 1
        w = row_nnz_kth_smallest(exec_, row_ptrs, pos)
        if self.kind == "imbalance_bounded_limit":
            w = min(w, int(n * self.ratio))
        return w


def _csr_convert_to(self, fmt, strategy=None, slice_size=64, stride_factor=1):
    """Csr::convert_to(Ell|Sellp|Hybrid|Coo) (reference core/matrix/csr.cpp:256-410)."""
    exec_, (n, m) = self.exec, self.size
    dev, V, I = exec_.device, self.V, self.I
    s = current_stream()
    if fmt == "coo":
        rows = torch.empty_like(self.col_idxs)
        check(getattr(lib, f"gkob200_convert_ptrs_to_idxs_{I}")(s, ptr(self.row_ptrs), n, ptr(rows)), "ptrs_to_idxs")
        return Coo(exec_, self.size, rows, self.col_idxs, self.values)
    if fmt == "ell":
        width, stride = self.max_row_nnz, n
        cols = torch.empty(width * stride, dtype=self.col_idxs.dtype, device=dev)
        vals = torch.empty(width * stride, dtype=self.values.dtype, device=dev)
        check(getattr(lib, f"gkob200_csr_convert_to_ell_{V}_{I}")(s, n, ptr(self.row_ptrs), ptr(self.col_idxs),
                                                                  ptr(self.values), width, stride, ptr(cols),
                                                                  ptr(vals)), "csr::convert_to_ell")
        return Ell(exec_, self.size, width, stride, cols, vals)
    if fmt == "sellp":
        ns = -(-n // slice_size)
        sets = torch.zeros(ns + 1, dtype=torch.int64, device=dev)
        lens = torch.zeros(max(ns, 1), dtype=torch.int64, device=dev)[:ns]
        ws, nb = _scan_ws(exec_, ns + 1)
        check(getattr(lib, f"gkob200_sellp_compute_slice_sets_{I}")(s, ptr(self.row_ptrs), n, slice_size,
                                                                    stride_factor, ptr(sets), ptr(lens), ptr(ws), nb),
              "sellp::compute_slice_sets")
        total = int(sets[ns].item()) * slice_size
        cols = torch.empty(total, dtype=self.col_idxs.dtype, device=dev)
        vals = torch.empty(total, dtype=self.values.dtype, device=dev)
        check(getattr(lib, f"gkob200_csr_convert_to_sellp_{V}_{I}")(s, n, ptr(self.row_ptrs), ptr(self.col_idxs),
                                                                    ptr(self.values), slice_size, ptr(sets),
                                                                    ptr(cols), ptr(vals)), "csr::convert_to_sellp")
        return Sellp(exec_, self.size, slice_size, stride_factor, sets, lens, cols, vals)
    if fmt == "hybrid":
        strategy = strategy or HybridStrategy.automatic()
        ell_lim = min(strategy.ell_width(exec_, self.row_ptrs), m)
        row_nnz = torch.empty(n, dtype=torch.int64, device=dev)
        check(getattr(lib, f"gkob200_convert_ptrs_to_sizes_{I}")(s, ptr(self.row_ptrs), n, ptr(row_nnz)),
              "ptrs_to_sizes")
        coo_ptrs = torch.zeros(n + 1, dtype=torch.int64, device=dev)
        ws, nb = _scan_ws(exec_, n + 1)
        check(lib.gkob200_hybrid_compute_coo_row_ptrs(s, ptr(row_nnz), n, ell_lim, ptr(coo_ptrs), ptr(ws), nb),
              "hybrid::compute_coo_row_ptrs")
        coo_nnz = int(coo_ptrs[n].item())
        ecols = torch.empty(ell_lim * n, dtype=self.col_idxs.dtype, device=dev)
        evals = torch.empty(ell_lim * n, dtype=self.values.dtype, device=dev)
        crows = torch.empty(coo_nnz, dtype=self.col_idxs.dtype, device=dev)
        ccols = torch.empty(coo_nnz, dtype=self.col_idxs.dtype, device=dev)
        cvals = torch.empty(coo_nnz, dtype=self.values.dtype, device=dev)
        check(getattr(lib, f"gkob200_csr_convert_to_hybrid_{V}_{I}")(
            s, n, ptr(self.row_ptrs), ptr(self.col_idxs), ptr(self.values), ptr(coo_ptrs), n, ell_lim, ptr(ecols),
            ptr(evals), ptr(crows), ptr(ccols), ptr(cvals)), "csr::convert_to_hybrid")
        return Hybrid(exec_, self.size, Ell(exec_, self.size, ell_lim, n, ecols, evals),
                      Coo(exec_, self.size, crows, ccols, cvals))
    raise Error(f"convert_to({fmt})", -2)


Csr.convert_to = _csr_convert_to
