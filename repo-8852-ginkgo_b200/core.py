"""Executor, error convention and small helpers of the host mirror."""
from __future__ import annotations

import ctypes as C

import numpy as np
import torch

from . import _abi
from . import lib


class Error(RuntimeError):
    """Raised when a C-ABI call returns non-zero — the mirror of gko::Error /
    gko::CudaError (reference include/ginkgo/core/base/exception.hpp:86-632)."""

    def __init__(self, fn, code):
        names = {-1: "invalid argument", -2: "unsupported on this path", -3: "workspace too small"}
        what = names.get(code, f"cudaError_t {code}" if code > 0 else str(code))
        super().__init__(f"{fn}: {what}")
        self.code = code


def check(code, fn="gkob200"):
    if code != 0:
        raise Error(fn, code)


def current_stream():
    """cudaStream_t of torch's current stream, as an integer for the C-ABI."""
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


_TORCH_V = {"f64": torch.float64, "f32": torch.float32}
_V_OF = {torch.float64: "f64", torch.float32: "f32"}
_I_OF = {torch.int32: "i32", torch.int64: "i64"}
_NP_V = {"f64": np.float64, "f32": np.float32}


def vname(dtype):
    return _V_OF[dtype]


def iname(dtype):
    return _I_OF[dtype]


def ptr(t):
    """Device (or host) pointer of a tensor / None as c_void_p."""
    if t is None:
        return C.c_void_p(0)
    return C.c_void_p(t.data_ptr())


class CudaExecutor:
    """Mirror of gko::CudaExecutor (reference include/ginkgo/core/base/executor.hpp:1384):
    owns the device id and the reduction scratch every reducing kernel needs.  Unlike the
    reference (everything on stream 0) kernels go to torch's current stream."""

    def __init__(self, device_id=0):
        if not torch.cuda.is_available():
            raise Error("CudaExecutor.create", 100)  # cudaErrorNoDevice
        self.device_id = device_id
        self.device = torch.device("cuda", device_id)
        torch.cuda.set_device(self.device)
        self._ws = torch.zeros(_abi.REDUCE_WS_BYTES, dtype=torch.uint8, device=self.device)
        check(lib.gkob200_reduce_ws_init(current_stream(), ptr(self._ws)), "reduce_ws_init")

    @classmethod
    def create(cls, device_id=0):
        return cls(device_id)

    @property
    def ws(self):
        return ptr(self._ws)

    def synchronize(self):
        torch.cuda.synchronize(self.device)

    def alloc(self, shape, dtype):
        return torch.empty(shape, dtype=dtype, device=self.device)

    def copy_from_host(self, arr):
        return torch.as_tensor(np.ascontiguousarray(arr)).to(self.device)

    def sm_count(self):
        return lib.gkob200_sm_count()
