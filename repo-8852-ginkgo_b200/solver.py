"""solver::Cg / Bicgstab / Gmres factories (reference include/ginkgo/core/solver/*.hpp):
`Cg.build().with_criteria(...).with_preconditioner(...).on(exec).generate(A)` returns a
solver LinOp whose apply(b, x) solves A x = b starting from the guess in x."""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _abi, lib
from . import stop as _stop
from .core import Error, check, current_stream, ptr


class _Factory:
    def __init__(self, kind):
        self.kind = kind
        self.criteria = []
        self.precond_factory = None
        self.generated_precond = None
        self.krylov_dim = 100  # reference default, include/ginkgo/core/solver/gmres.hpp:57
        self.check_every = 8
        self.exec = None

    def with_criteria(self, *criteria):
        self.criteria = list(criteria)
        return self

    def with_preconditioner(self, factory):
        self.precond_factory = factory
        return self

    def with_generated_preconditioner(self, precond):
        self.generated_precond = precond
        return self

    def with_krylov_dim(self, m):
        self.krylov_dim = int(m)
        return self

    def with_check_every(self, n):
        """How many iterations are enqueued between two host polls of the device-side
        stopping status (the iteration count reported is exact regardless)."""
        self.check_every = int(n)
        return self

    def on(self, exec_):
        self.exec = exec_
        return self

    def generate(self, A, nrhs=1):
        precond = self.generated_precond
        if precond is None and self.precond_factory is not None:
            precond = self.precond_factory.on(self.exec).generate(A)
        return _Solver(self, A, precond, nrhs)


class _Solver:
    def __init__(self, factory, A, precond, nrhs):
        self.exec = factory.exec
        self.A, self.precond, self.nrhs = A, precond, nrhs
        self.size = A.size
        self._adesc = A.descriptor()
        self._pdesc = precond.descriptor() if precond is not None else None
        self._sdesc = _stop.to_descriptor(factory.criteria, factory.check_every)
        h = C.c_void_p()
        check(lib.gkob200_solver_create(factory.kind, C.byref(self._adesc),
                                        C.byref(self._pdesc) if self._pdesc is not None else None,
                                        C.byref(self._sdesc), nrhs, factory.krylov_dim, C.byref(h)),
              "solver::generate")
        self._h = h

    def __del__(self):
        h, self._h = getattr(self, "_h", None), None
        if h and lib is not None:  # lib is None while the interpreter shuts down
            lib.gkob200_solver_destroy(h)

    def get_system_matrix(self):
        return self.A

    def get_preconditioner(self):
        return self.precond

    def apply(self, b, x):
        if b.size != x.size or b.size[0] != self.size[0] or b.size[1] != self.nrhs:
            raise Error("solver::apply DimensionMismatch", -1)
        check(lib.gkob200_solver_apply(self._h, current_stream(), ptr(b.t), b.stride, ptr(x.t), x.stride),
              "solver::apply")
        return x

    def apply_host(self, b_host, x_host):
        """b_host/x_host: contiguous host tensors or numpy arrays (pinned for best speed)."""
        bp = b_host.data_ptr() if hasattr(b_host, "data_ptr") else b_host.ctypes.data
        xp = x_host.data_ptr() if hasattr(x_host, "data_ptr") else x_host.ctypes.data
        check(lib.gkob200_solver_apply_host(self._h, current_stream(), C.c_void_p(bp), C.c_void_p(xp)),
              "solver::apply_host")
        return x_host

    @property
    def num_iterations(self):
        return int(lib.gkob200_solver_num_iterations(self._h))

    @property
    def launch_count(self):
        return int(lib.gkob200_solver_launch_count(self._h))

    @property
    def stop_status(self):
        out = np.zeros(self.nrhs, dtype=np.uint8)
        check(lib.gkob200_solver_stop_status(self._h, C.c_void_p(out.ctypes.data)), "solver::stop_status")
        return out

    @property
    def residual_history(self):
        cap = self.num_iterations + 1
        out = np.zeros(cap, dtype=np.float64)
        m = lib.gkob200_solver_residual_history(self._h, C.c_void_p(out.ctypes.data), cap)
        return out[:m]


class Cg:
    @staticmethod
    def build():
        return _Factory(_abi.SOLVER_CG)


class Bicgstab:
    @staticmethod
    def build():
        return _Factory(_abi.SOLVER_BICGSTAB)


class Gmres:
    @staticmethod
    def build():
        return _Factory(_abi.SOLVER_GMRES)


class Fcg:
    """solver::Fcg (reference core/solver/fcg.cpp)."""

    @staticmethod
    def build():
        return _Factory(_abi.SOLVER_FCG)


class Cgs:
    """solver::Cgs (reference core/solver/cgs.cpp)."""

    @staticmethod
    def build():
        return _Factory(_abi.SOLVER_CGS)
