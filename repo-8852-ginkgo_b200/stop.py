"""Stopping criteria factories (reference include/ginkgo/core/stop/*.hpp)."""
from . import _abi


class Iteration:
    def __init__(self, max_iters):
        self.max_iters = int(max_iters)


class ResidualNorm:
    """stop::ResidualNorm with baseline mode rhs_norm | initial_resnorm | absolute
    (reference core/stop/residual_norm.cpp:129-186)."""

    MODES = {"rhs_norm": _abi.STOP_RHS_NORM, "initial_resnorm": _abi.STOP_INITIAL_RESNORM,
             "absolute": _abi.STOP_ABSOLUTE}

    def __init__(self, reduction_factor, baseline="rhs_norm"):
        self.reduction_factor = float(reduction_factor)
        self.baseline = self.MODES[baseline]


def to_descriptor(criteria, check_every=8):
    s = _abi.Stop()
    s.max_iters = 2 ** 31 - 2
    s.reduction_factor = 0.0
    s.baseline = _abi.STOP_RHS_NORM
    s.check_every = check_every
    for c in criteria:
        if isinstance(c, Iteration):
            s.max_iters = c.max_iters
        elif isinstance(c, ResidualNorm):
            s.reduction_factor = c.reduction_factor
            s.baseline = c.baseline
        else:
            raise TypeError(c)
    return s
