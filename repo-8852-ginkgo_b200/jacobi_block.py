"""Block-Jacobi host logic: block detection, storage scheme, generate, apply
(reference core/preconditioner/jacobi.cpp:260-332, jacobi.hpp:578-610)."""
from __future__ import annotations

import torch

from . import _abi, lib
from .core import Error, check, current_stream, ptr


def storage_scheme(max_block_size, max_block_stride=32):
    """block_interleaved_storage_scheme for `max_block_size`
    (reference include/ginkgo/core/preconditioner/jacobi.hpp:578-610, compute_storage_scheme)."""
    if max_block_size > max_block_stride or max_block_size < 1:
        raise Error("Jacobi max_block_size", -2)
    sup = 1
    while sup < max_block_size:
        sup *= 2
    group_size = max_block_stride // sup
    block_offset = max_block_size
    block_stride = group_size * block_offset
    group_offset = max_block_size * block_stride
    group_power = group_size.bit_length() - 1
    return block_offset, group_offset, group_power


def generate_block_jacobi(self, A):
    exec_, n = self.exec, A.size[0]
    if A.I != "i32":
        raise Error("block Jacobi index type", -2)
    dev = exec_.device
    s = current_stream()
    nb = torch.zeros(1, dtype=torch.int64, device=dev)
    self.block_pointers = torch.zeros(n + 1, dtype=torch.int32, device=dev)
    wsb = lib.gkob200_jacobi_find_blocks_workspace_bytes(n)
    ws = torch.empty(wsb, dtype=torch.uint8, device=dev)
    check(lib.gkob200_jacobi_find_blocks_i32(s, n, ptr(A.row_ptrs), ptr(A.col_idxs), self.max_block_size, ptr(nb),
                                             ptr(self.block_pointers), ptr(ws), wsb), "jacobi::find_blocks")
    self.num_blocks = int(nb.item())
    self.block_offset, self.group_offset, self.group_power = storage_scheme(self.max_block_size)
    gs = 1 << self.group_power
    storage = -(-self.num_blocks // gs) * self.group_offset
    self.blocks = torch.zeros(max(storage, 1), dtype=A.values.dtype, device=dev)[:storage]
    fn = getattr(lib, f"gkob200_jacobi_block_generate_{self.V}")
    check(fn(s, n, ptr(A.row_ptrs), ptr(A.col_idxs), ptr(A.values), self.num_blocks, ptr(self.block_pointers),
             self.block_offset, self.group_offset, self.group_power, ptr(self.blocks)), "jacobi::generate")


def apply_block_jacobi(self, *args):
    if len(args) == 2:
        b, x = args
        fn = getattr(lib, f"gkob200_jacobi_block_simple_apply_{self.V}")
        check(fn(current_stream(), self.num_blocks, ptr(self.block_pointers), ptr(self.blocks), self.block_offset,
                 self.group_offset, self.group_power, x.size[0], x.size[1], ptr(b.t), b.stride, ptr(x.t), x.stride),
              "jacobi::simple_apply")
    else:
        alpha, b, beta, x = args
        fn = getattr(lib, f"gkob200_jacobi_block_apply_{self.V}")
        check(fn(current_stream(), self.num_blocks, ptr(self.block_pointers), ptr(self.blocks), self.block_offset,
                 self.group_offset, self.group_power, x.size[0], x.size[1], ptr(alpha.t), ptr(b.t), b.stride,
                 ptr(beta.t), ptr(x.t), x.stride), "jacobi::apply")
    return x


def fill_descriptor(self, d):
    d.kind = _abi.PRECOND_JACOBI_BLOCK
    d.num_blocks = self.num_blocks
    d.block_pointers = self.block_pointers.data_ptr()
    d.blocks = self.blocks.data_ptr()
    d.block_offset, d.group_offset = self.block_offset, self.group_offset
    d.group_power, d.max_block_size = self.group_power, self.max_block_size
