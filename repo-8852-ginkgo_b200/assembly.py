"""device_matrix_data and the CSR assembly steps on the device: the mirror of
gko::device_matrix_data<V,I> (reference include/ginkgo/core/base/device_matrix_data.hpp,
core/base/device_matrix_data.cpp) and of Csr::read / transpose / sort_by_column_index
(core/matrix/csr.cpp:452-468, 548-580) over the setup kernels of the C-ABI."""
from __future__ import annotations

import numpy as np
import torch

from . import lib
from .core import check, current_stream, iname, ptr, vname
from .matrix import Csr


def _sort_ws(exec_, nnz, values, idxs):
    nb = lib.gkob200_setup_sort_workspace_bytes(nnz, values.element_size(), idxs.element_size())
    return torch.empty(nb, dtype=torch.uint8, device=exec_.device), nb


class DeviceMatrixData:
    """COO triplets on the device: row_idxs, col_idxs, values (any order, duplicates allowed)."""

    def __init__(self, exec_, size, row_idxs, col_idxs, values):
        self.exec, self.size = exec_, tuple(size)
        self.row_idxs, self.col_idxs, self.values = row_idxs, col_idxs, values

    @classmethod
    def from_arrays(cls, exec_, size, rows, cols, vals):
        dev = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(exec_.device)  # noqa: E731
        return cls(exec_, size, dev(rows), dev(cols), dev(vals))

    @property
    def num_elems(self):
        return int(self.values.numel())

    def _sfx(self):
        return f"{vname(self.values.dtype)}_{iname(self.row_idxs.dtype)}"

    def sort_row_major(self):
        """device_matrix_data::sort_row_major (stable)."""
        n = self.num_elems
        ws, nb = _sort_ws(self.exec, n, self.values, self.row_idxs)
        fn = getattr(lib, f"gkob200_coo_sort_row_major_{self._sfx()}")
        check(fn(current_stream(), self.size[0], self.size[1], n, ptr(self.row_idxs), ptr(self.col_idxs),
                 ptr(self.values), ptr(ws), nb), "sort_row_major")
        return self

    def _compact(self, name):
        n = self.num_elems
        nb = lib.gkob200_setup_compact_workspace_bytes(n)
        ws = torch.empty(nb, dtype=torch.uint8, device=self.exec.device)
        r, c, v = torch.empty_like(self.row_idxs), torch.empty_like(self.col_idxs), torch.empty_like(self.values)
        cnt = torch.zeros(1, dtype=torch.int64, device=self.exec.device)
        fn = getattr(lib, f"gkob200_coo_{name}_{self._sfx()}")
        check(fn(current_stream(), n, ptr(self.row_idxs), ptr(self.col_idxs), ptr(self.values), ptr(r), ptr(c), ptr(v),
                 ptr(cnt), ptr(ws), nb), name)
        m = int(cnt.item())
        self.row_idxs, self.col_idxs, self.values = r[:m], c[:m], v[:m]
        return self

    def sum_duplicates(self):
        """device_matrix_data::sum_duplicates: sorts row-major first, like the reference
        (core/base/device_matrix_data.cpp), then adds the entries of equal positions in order."""
        self.sort_row_major()
        return self._compact("sum_duplicates")

    def remove_zeros(self):
        return self._compact("remove_zeros")

    def to_csr(self, strategy="automatical"):
        """Csr::read(device_matrix_data): the data must be row-major sorted (the reference
        sorts it in read(); call sort_row_major() / sum_duplicates() first)."""
        n = self.size[0]
        I = iname(self.row_idxs.dtype)  # noqa: E741
        rp = torch.empty(n + 1, dtype=self.row_idxs.dtype, device=self.exec.device)
        check(getattr(lib, f"gkob200_convert_idxs_to_ptrs_{I}")(current_stream(), ptr(self.row_idxs), self.num_elems, n,
                                                                ptr(rp)), "idxs_to_ptrs")
        return Csr(self.exec, self.size, rp, self.col_idxs, self.values, strategy)


def csr_read(exec_, size, rows, cols, vals, strategy="automatical"):
    """Csr::read(matrix_data): sort_row_major + sum_duplicates? — the reference's read() sorts
    and keeps explicit zeros and duplicates out of scope (matrix_data is expected clean);
    this helper does what benchmark drivers do: sort, sum duplicates, build the CSR."""
    return DeviceMatrixData.from_arrays(exec_, size, rows, cols, vals).sum_duplicates().to_csr(strategy)


def transpose(A: Csr):
    """Csr::transpose."""
    n, m = A.size
    ws, nb = _sort_ws(A.exec, A.nnz, A.values, A.col_idxs)
    rp = torch.empty(m + 1, dtype=A.row_ptrs.dtype, device=A.exec.device)
    ci, va = torch.empty_like(A.col_idxs), torch.empty_like(A.values)
    fn = getattr(lib, f"gkob200_csr_transpose_{A.V}_{A.I}")
    check(fn(current_stream(), n, m, A.nnz, ptr(A.row_ptrs), ptr(A.col_idxs), ptr(A.values), ptr(rp), ptr(ci), ptr(va),
             ptr(ws), nb), "csr::transpose")
    return Csr(A.exec, (m, n), rp, ci, va, A.strategy)


def sort_by_column_index(A: Csr):
    """Csr::sort_by_column_index (in place)."""
    ws, nb = _sort_ws(A.exec, A.nnz, A.values, A.col_idxs)
    fn = getattr(lib, f"gkob200_csr_sort_by_column_index_{A.V}_{A.I}")
    check(fn(current_stream(), A.size[0], A.size[1], A.nnz, ptr(A.row_ptrs), ptr(A.col_idxs), ptr(A.values), ptr(ws), nb),
          "csr::sort_by_column_index")
    return A
