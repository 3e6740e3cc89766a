"""gko_b200 — Python host mirror of Ginkgo's LinOp / Executor API over libgko_b200.so.

This package is the TEST AND BENCH HARNESS side of the drop-in: it drives the
C-ABI of ``include/gko_b200.h`` through ctypes, with PyTorch used only for device
memory, streams and ``torch.distributed`` plumbing.  The class and method names
follow the reference (``matrix.Csr.apply``, ``solver.Cg.build().with_criteria(...)
.on(exec).generate(A)``, ``preconditioner.Jacobi``, ``stop.Iteration`` ...;
reference include/ginkgo/core/{matrix,solver,preconditioner,stop}/*.hpp) so that the
parity tests read like the reference's own tests.

There is NO CPU fallback: importing this package without the compiled CUDA
library raises, and every compute call goes to a hand-written sm_100a kernel.
"""
from __future__ import annotations

import ctypes
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libgko_b200.so")

if not os.path.exists(LIB_PATH):
    raise ImportError(
        f"{LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
        "(make -C repo-8852-ginkgo_b200/csrc). gko_b200 has no CPU fallback."
    )


def _load():
    # NCCL symbols (distributed halo exchange) resolve against the copy torch ships.
    try:
        import torch  # noqa: F401  (loads libnccl / libcudart into the process)
    except Exception:  # pragma: no cover - torch is part of the image
        pass
    return ctypes.CDLL(LIB_PATH, mode=ctypes.RTLD_GLOBAL)


lib = _load()

from . import _abi  # noqa: E402

_abi.declare(lib)

from .core import (  # noqa: E402,F401
    CudaExecutor,
    Error,
    check,
    current_stream,
)
from . import matrix, solver, preconditioner, stop, gen, distributed, assembly, io  # noqa: E402,F401

__all__ = [
    "lib",
    "LIB_PATH",
    "CudaExecutor",
    "Error",
    "matrix",
    "solver",
    "preconditioner",
    "stop",
    "gen",
    "distributed",
]
