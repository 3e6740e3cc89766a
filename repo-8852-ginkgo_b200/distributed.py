"""experimental::distributed::{Partition, Matrix, Vector} host mirror
(reference include/ginkgo/core/distributed/*.hpp, core/distributed/*.cpp): a row-partitioned
matrix over the GPUs of one box, one process per GPU.  torch.distributed only provides the
rendezvous (to hand the NCCL unique id to every rank); the halo exchange and the
reductions run inside libgko_b200.so on its own NCCL communicator."""
from __future__ import annotations

import ctypes as C

import numpy as np
import torch

from . import _abi, lib
from . import stop as _stop
from .core import Error, check, current_stream, ptr, vname
from .matrix import Csr, Dense, _SparseBase
from .solver import _Solver


class Communicator:
    """The used subset of mpi::communicator (reference include/ginkgo/core/base/mpi.hpp:436-1500)
    on NCCL.  `Communicator.from_torch(exec)` builds it from an initialised
    torch.distributed process group; `Communicator.single(exec)` is the 1-rank case."""

    def __init__(self, exec_, handle, rank, size):
        self.exec, self._h, self.rank, self.size = exec_, handle, rank, size

    @classmethod
    def single(cls, exec_):
        h = C.c_void_p()
        check(lib.gkob200_dist_comm_create(None, 0, 1, C.byref(h)), "comm_create")
        return cls(exec_, h, 0, 1)

    @classmethod
    def from_torch(cls, exec_, group=None):
        import torch.distributed as dist
        rank, size = dist.get_rank(group), dist.get_world_size(group)
        if size == 1:
            return cls.single(exec_)
        idt = torch.zeros(128, dtype=torch.uint8)
        if rank == 0:
            buf = (C.c_uint8 * 128)()
            check(lib.gkob200_nccl_unique_id(buf), "nccl_unique_id")
            idt = torch.tensor(list(buf), dtype=torch.uint8)
        dev = exec_.device if dist.get_backend(group) == "nccl" else torch.device("cpu")
        idt = idt.to(dev)
        dist.broadcast(idt, src=0, group=group)
        raw = bytes(idt.cpu().tolist())
        h = C.c_void_p()
        check(lib.gkob200_dist_comm_create(raw, rank, size, C.byref(h)), "comm_create")
        return cls(exec_, h, rank, size)

    def __del__(self):
        h, self._h = getattr(self, "_h", None), None
        if h and lib is not None:  # lib is None while the interpreter shuts down
            lib.gkob200_dist_comm_destroy(h)

    @property
    def uses_p2p(self):
        """scalar all-reduces run over peer memory (CUDA IPC) instead of NCCL"""
        return bool(lib.gkob200_dist_comm_uses_p2p(self._h))

    def all_reduce_sum(self, t):
        fn = lib.gkob200_dist_allreduce_sum_f64 if t.dtype == torch.float64 else lib.gkob200_dist_allreduce_sum_f32
        check(fn(self._h, current_stream(), ptr(t), t.numel()), "all_reduce")
        return t

    def all_to_all_i64(self, send):
        recv = torch.empty_like(send)
        check(lib.gkob200_dist_alltoall_i64(self._h, current_stream(), ptr(send), ptr(recv), send.numel() // self.size),
              "all_to_all")
        return recv

    def all_to_all_v_i32(self, send, send_sizes, recv_sizes):
        so = np.concatenate([[0], np.cumsum(send_sizes)]).astype(np.int64)
        ro = np.concatenate([[0], np.cumsum(recv_sizes)]).astype(np.int64)
        ss, rs = np.asarray(send_sizes, np.int64), np.asarray(recv_sizes, np.int64)
        recv = torch.empty(int(ro[-1]), dtype=torch.int32, device=self.exec.device)
        check(lib.gkob200_dist_alltoallv_i32(self._h, current_stream(), ptr(send), ss.ctypes.data, so.ctypes.data,
                                             ptr(recv), rs.ctypes.data, ro.ctypes.data), "all_to_all_v")
        torch.cuda.current_stream().synchronize()
        return recv


class Partition:
    """distributed::Partition<int32, int64> (reference core/distributed/partition.cpp:60-140)."""

    def __init__(self, exec_, num_parts, range_bounds, part_ids):
        self.exec, self.num_parts = exec_, num_parts
        self.range_bounds, self.part_ids = range_bounds, part_ids
        self.num_ranges = part_ids.numel()
        dev = exec_.device
        self.range_starting_indices = torch.zeros(self.num_ranges, dtype=torch.int32, device=dev)
        self.part_sizes = torch.zeros(num_parts, dtype=torch.int32, device=dev)
        ne = torch.zeros(1, dtype=torch.int32, device=dev)
        check(lib.gkob200_partition_build_starting_indices_i32_i64(
            current_stream(), ptr(range_bounds), ptr(part_ids), self.num_ranges, num_parts, ptr(ne),
            ptr(self.range_starting_indices), ptr(self.part_sizes)), "partition::build_starting_indices")
        self.num_empty_parts = int(ne.item())
        self.size = int(range_bounds[-1].item())

    @classmethod
    def build_from_global_size_uniform(cls, exec_, num_parts, global_size):
        dev = exec_.device
        ranges = torch.zeros(num_parts + 1, dtype=torch.int64, device=dev)
        check(lib.gkob200_partition_build_ranges_from_global_size_i64(current_stream(), num_parts, global_size,
                                                                      ptr(ranges)), "partition::build_ranges")
        return cls.build_from_contiguous(exec_, ranges)

    @classmethod
    def build_from_contiguous(cls, exec_, ranges):
        dev = exec_.device
        ranges = ranges.to(dev)
        num_parts = ranges.numel() - 1
        bounds = torch.zeros(num_parts + 1, dtype=torch.int64, device=dev)
        ids = torch.zeros(num_parts, dtype=torch.int32, device=dev)
        check(lib.gkob200_partition_build_from_contiguous_i64(current_stream(), num_parts, ptr(ranges), ptr(bounds),
                                                              ptr(ids)), "partition::build_from_contiguous")
        return cls(exec_, num_parts, bounds, ids)

    @classmethod
    def build_from_mapping(cls, exec_, mapping, num_parts):
        dev = exec_.device
        mapping = mapping.to(dev).to(torch.int32)
        n = mapping.numel()
        bounds = torch.zeros(n + 1, dtype=torch.int64, device=dev)
        ids = torch.zeros(max(n, 1), dtype=torch.int32, device=dev)
        nr = torch.zeros(1, dtype=torch.int64, device=dev)
        wsb = (n + 2) * 4 + lib.gkob200_prefix_sum_workspace_bytes(n + 1) + 64
        ws = torch.empty(wsb, dtype=torch.uint8, device=dev)
        check(lib.gkob200_partition_build_from_mapping_i64(current_stream(), n, ptr(mapping), ptr(bounds), ptr(ids),
                                                           ptr(nr), ptr(ws), wsb), "partition::build_from_mapping")
        k = int(nr.item())
        return cls(exec_, num_parts, bounds[: k + 1].contiguous(), ids[:k].contiguous())

    def get_part_size(self, part):
        return int(self.part_sizes[part].item())

    def view_args(self):
        return (self.num_ranges, ptr(self.range_bounds), ptr(self.part_ids), ptr(self.range_starting_indices))


def build_local_nonlocal(exec_, rows, cols, vals, row_part, col_part, local_part):
    """distributed_matrix::build_local_nonlocal on the device; returns a dict of tensors."""
    dev = exec_.device
    nnz = rows.numel()
    V = vname(vals.dtype)
    i32 = lambda n: torch.empty(max(n, 1), dtype=torch.int32, device=dev)  # noqa: E731
    out = dict(lrow=i32(nnz), lcol=i32(nnz), lval=torch.empty(max(nnz, 1), dtype=vals.dtype, device=dev),
               nrow=i32(nnz), ncol=i32(nnz), nval=torch.empty(max(nnz, 1), dtype=vals.dtype, device=dev),
               gather=i32(nnz), recv_sizes=torch.zeros(row_part.num_parts, dtype=torch.int32, device=dev),
               nl_to_global=torch.empty(max(nnz, 1), dtype=torch.int64, device=dev))
    counts = (C.c_int64 * 3)()
    fn = getattr(lib, f"gkob200_dist_build_local_nonlocal_{V}")
    check(fn(current_stream(), nnz, ptr(rows), ptr(cols), ptr(vals), *row_part.view_args(), *col_part.view_args(),
             col_part.size, row_part.num_parts, local_part, ptr(out["lrow"]), ptr(out["lcol"]), ptr(out["lval"]),
             ptr(out["nrow"]), ptr(out["ncol"]), ptr(out["nval"]), ptr(out["gather"]), ptr(out["recv_sizes"]),
             ptr(out["nl_to_global"]), counts), "distributed_matrix::build_local_nonlocal")
    nl, nn, ng = counts[0], counts[1], counts[2]
    for k_ in ("lrow", "lcol", "lval"):
        out[k_] = out[k_][:nl].clone()
    for k_ in ("nrow", "ncol", "nval"):
        out[k_] = out[k_][:nn].clone()
    out["gather"] = out["gather"][:ng].clone()
    out["nl_to_global"] = out["nl_to_global"][:ng].clone()
    return out


def _coo_to_csr(exec_, n_rows, n_cols, rows, cols, vals, strategy):
    rp = torch.zeros(n_rows + 1, dtype=torch.int32, device=exec_.device)
    check(lib.gkob200_convert_idxs_to_ptrs_i32(current_stream(), ptr(rows), rows.numel(), n_rows, ptr(rp)),
          "convert_idxs_to_ptrs")
    return Csr(exec_, (n_rows, n_cols), rp, cols, vals, strategy)


class CsrRows(_SparseBase):
    """Row-compressed CSR (GKOB200_FMT_CSR_ROWS): only the rows that have entries are stored.
    The non-local block of a slab-partitioned stencil has entries on the two slab faces only
    (2 x 40 000 of 8 000 000 rows at 200^3), so a full-height CSR pass over it would cost more
    than the halo itself.  apply() accumulates: x[row] = beta x[row] + alpha (A b)[row]."""

    def __init__(self, exec_, size, rows_coo, cols, vals):
        self.exec, self.size = exec_, tuple(size)
        dev = exec_.device
        if rows_coo.numel():
            uniq, counts = torch.unique_consecutive(rows_coo, return_counts=True)
        else:
            uniq = torch.zeros(0, dtype=torch.int32, device=dev)
            counts = torch.zeros(0, dtype=torch.int64, device=dev)
        self.row_list = uniq.to(torch.int32).contiguous()
        self.row_ptrs = torch.zeros(uniq.numel() + 1, dtype=torch.int32, device=dev)
        if uniq.numel():
            self.row_ptrs[1:] = torch.cumsum(counts, 0).to(torch.int32)
        self.col_idxs, self.values = cols, vals
        self.nnz = int(vals.numel())
        self.V = vname(vals.dtype)

    def descriptor(self):
        d = _abi.Matrix()
        d.format = _abi.FMT_CSR_ROWS
        d.value_type = _abi.F64 if self.V == "f64" else _abi.F32
        d.index_type = _abi.I32
        d.n_rows, d.n_cols = self.size
        d.nnz = self.nnz
        d.row_ptrs, d.col_idxs, d.values = self.row_ptrs.data_ptr(), self.col_idxs.data_ptr(), self.values.data_ptr()
        d.row_list, d.n_listed = self.row_list.data_ptr(), self.row_list.numel()
        return d

    def spmv_bytes(self, nrhs=1):
        v = self.values.element_size()
        return self.nnz * (v + 4) + self.row_list.numel() * (8 + 2 * nrhs * v) + self.size[1] * nrhs * v


def halo_plan(comm, recv_sizes, recv_gather_idxs):
    """The two exchanges of read_distributed (reference core/distributed/matrix.cpp:197-221):
    step 1: all_to_all of one count per peer turns recv_sizes into send_sizes; step 2: all_to_all_v
    sends the gather indices from the receivers to the senders.  `comm` is any object with
    all_to_all_i64 / all_to_all_v_i32 (the NCCL Communicator here; a gloo double in the CPU tests).
    Returns (send_sizes, recv_sizes) as host int64 arrays and the gather index tensor."""
    rs = recv_sizes.cpu().numpy().astype(np.int64)
    ss = comm.all_to_all_i64(torch.from_numpy(rs).to(recv_sizes.device)).cpu().numpy().astype(np.int64)
    gather = comm.all_to_all_v_i32(recv_gather_idxs, rs, ss)
    return ss, rs, gather


class Matrix(_SparseBase):
    """experimental::distributed::Matrix<V, int32, int64> (reference core/distributed/matrix.cpp):
    local block and non-local (ghost-column) block, both CSR as in the reference default
    (matrix.cpp:62), and the halo plan."""

    def __init__(self, exec_, comm):
        self.exec, self.comm = exec_, comm
        self._h = None

    def read_distributed(self, rows, cols, vals, row_part, col_part=None, local_strategy="automatical"):
        """rows/cols: int64 GLOBAL indices (device or host arrays), row-major sorted.
        Collective over the communicator (every rank reads its part)."""
        exec_, comm = self.exec, self.comm
        col_part = col_part or row_part
        dev = exec_.device
        to_dev = lambda a: (a if isinstance(a, torch.Tensor) else torch.from_numpy(np.ascontiguousarray(a))).to(dev)  # noqa
        rows, cols, vals = to_dev(rows), to_dev(cols), to_dev(vals)
        parts = build_local_nonlocal(exec_, rows, cols, vals, row_part, col_part, comm.rank)
        self.size = (row_part.size, col_part.size)
        n_loc_rows, n_loc_cols = row_part.get_part_size(comm.rank), col_part.get_part_size(comm.rank)
        n_ghost = parts["gather"].numel()
        self.local = _coo_to_csr(exec_, n_loc_rows, n_loc_cols, parts["lrow"], parts["lcol"], parts["lval"], local_strategy)
        # non-local block: same entries as the reference's Csr non_local_mtx_, stored row-compressed
        self.non_local = CsrRows(exec_, (n_loc_rows, n_ghost), parts["nrow"], parts["ncol"], parts["nval"])
        self.non_local_to_global = parts["nl_to_global"]
        self.send_sizes, self.recv_sizes, self.gather_idxs = halo_plan(comm, parts["recv_sizes"], parts["gather"])
        self._ld, self._nd = self.local.descriptor(), self.non_local.descriptor()
        h = C.c_void_p()
        check(lib.gkob200_dist_matrix_create(comm._h, C.byref(self._ld), C.byref(self._nd), ptr(self.gather_idxs),
                                             self.send_sizes.ctypes.data, self.recv_sizes.ctypes.data, C.byref(h)),
              "distributed::Matrix")
        self._h = h
        return self

    def __del__(self):
        h, self._h = getattr(self, "_h", None), None
        if h and lib is not None:  # lib is None while the interpreter shuts down
            lib.gkob200_dist_matrix_destroy(h)

    @property
    def uses_fused_halo(self):
        """apply() is one launch: halo push over peer memory + local + non-local rows"""
        return bool(lib.gkob200_dist_matrix_uses_fused_halo(self._h))

    def apply(self, *args):
        """b, x are the LOCAL parts (Dense) of the distributed vectors."""
        if len(args) == 2:
            (b, x), alpha, beta = args, None, None
        else:
            alpha, b, beta, x = args
        check(lib.gkob200_dist_matrix_apply(self._h, current_stream(), ptr(b.t), b.stride, b.size[1],
                                            ptr(alpha.t) if alpha is not None else None,
                                            ptr(beta.t) if beta is not None else None, ptr(x.t), x.stride),
              "distributed::Matrix::apply")
        return x

    def spmv_bytes(self, nrhs=1):
        return self.local.spmv_bytes(nrhs) + self.non_local.spmv_bytes(nrhs)


class Vector:
    """experimental::distributed::Vector reductions over the local Dense parts
    (reference core/distributed/vector.cpp:309-440): local reduce, all_reduce(sum), and for
    norm2 the sqrt AFTER the reduction of the local squared norms."""

    def __init__(self, comm, local):
        self.comm, self.local = comm, local

    def compute_dot(self, other, result):
        self.local.compute_dot(other.local, result)
        self.comm.all_reduce_sum(result.t)

    compute_conj_dot = compute_dot

    def compute_norm2(self, result):
        self.local.compute_squared_norm2(result)
        self.comm.all_reduce_sum(result.t)
        fn = getattr(lib, f"gkob200_dense_compute_sqrt_{result.V}")
        check(fn(current_stream(), result.size[1], ptr(result.t)), "compute_sqrt")

    def compute_norm1(self, result):
        self.local.compute_norm1(result)
        self.comm.all_reduce_sum(result.t)


class _DistSolver(_Solver):
    def __init__(self, exec_, kind, A, precond, criteria, check_every):
        self.exec, self.A, self.precond, self.nrhs = exec_, A, precond, 1
        n = A.local.size[0]
        self.size = (n, n)
        self._pdesc = precond.descriptor() if precond is not None else None
        self._sdesc = _stop.to_descriptor(criteria, check_every)
        h = C.c_void_p()
        check(lib.gkob200_dist_solver_create(kind, A._h, C.byref(self._pdesc) if self._pdesc is not None else None,
                                             C.byref(self._sdesc), 1, C.byref(h)), "distributed solver::generate")
        self._h = h


def cg(exec_, A, criteria, precond=None, check_every=8):
    """solver::Cg on a distributed::Matrix (b, x passed to apply() are the local parts)."""
    return _DistSolver(exec_, _abi.SOLVER_CG, A, precond, criteria, check_every)
