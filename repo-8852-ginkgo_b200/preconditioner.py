"""preconditioner::Jacobi (reference include/ginkgo/core/preconditioner/jacobi.hpp)."""
from __future__ import annotations

import torch

from . import _abi, lib
from .core import check, current_stream, ptr, vname


class _JacobiFactory:
    def __init__(self):
        self.max_block_size = 32
        self.exec = None

    def with_max_block_size(self, n):
        self.max_block_size = int(n)
        return self

    def on(self, exec_):
        self.exec = exec_
        return self

    def generate(self, A):
        return Jacobi(self.exec, A, self.max_block_size)


class Jacobi:
    """max_block_size == 1: scalar Jacobi — extract_diagonal + invert_diagonal
    (reference core/preconditioner/jacobi.cpp:315-332)."""

    @staticmethod
    def build():
        return _JacobiFactory()

    def __init__(self, exec_, A, max_block_size):
        self.exec = exec_
        self.size = A.size
        self.max_block_size = max_block_size
        self.V = A.V
        if max_block_size == 1:
            diag = A.extract_diagonal()
            self.inv_diag = torch.empty_like(diag)
            fn = getattr(lib, f"gkob200_jacobi_invert_diagonal_{self.V}")
            check(fn(current_stream(), diag.numel(), ptr(diag), ptr(self.inv_diag)), "jacobi::invert_diagonal")
        else:
            from .jacobi_block import generate_block_jacobi
            generate_block_jacobi(self, A)

    def apply(self, *args):
        if self.max_block_size != 1:
            from .jacobi_block import apply_block_jacobi
            return apply_block_jacobi(self, *args)
        if len(args) == 2:
            b, x = args
            fn = getattr(lib, f"gkob200_jacobi_simple_scalar_apply_{self.V}")
            check(fn(current_stream(), x.size[0], x.size[1], ptr(self.inv_diag), ptr(b.t), b.stride, ptr(x.t),
                     x.stride), "jacobi::simple_scalar_apply")
        else:
            alpha, b, beta, x = args
            fn = getattr(lib, f"gkob200_jacobi_scalar_apply_{self.V}")
            check(fn(current_stream(), x.size[0], x.size[1], ptr(self.inv_diag), ptr(alpha.t), ptr(b.t), b.stride,
                     ptr(beta.t), ptr(x.t), x.stride), "jacobi::scalar_apply")
        return x

    def descriptor(self):
        d = _abi.Precond()
        d.value_type = _abi.F64 if self.V == "f64" else _abi.F32
        if self.max_block_size == 1:
            d.kind = _abi.PRECOND_JACOBI_SCALAR
            d.inv_diag = self.inv_diag.data_ptr()
        else:
            from .jacobi_block import fill_descriptor
            fill_descriptor(self, d)
        return d

    def storage_bytes(self):
        if self.max_block_size == 1:
            return self.inv_diag.numel() * self.inv_diag.element_size()
        return self.blocks.numel() * self.blocks.element_size()
