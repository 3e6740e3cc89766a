"""The oracle's stencil generator (used by bench.py's reference arm, which must not load the
product library) against the product's host generator: identical CSR arrays."""
import numpy as np
import pytest

import oracle
from __graft_entry__ import load_package


@pytest.mark.parametrize("kind,dims", [("5pt", (13, 9, 1)), ("7pt", (7, 5, 6)), ("27pt", (6, 7, 5)), ("27pt", (1, 1, 1)),
                                       ("7pt", (1, 4, 1))])
def test_generators_agree(kind, dims):
    gko = load_package()
    nx, ny, nz = dims
    a = gko.gen.stencil_csr(kind, nx, ny, nz)
    b = oracle.gen_stencil_csr(kind, nx, ny, nz)
    assert a[3] == b[3]
    for x, y in zip(a[:3], b[:3]):
        assert x.dtype == y.dtype and np.array_equal(x, y)
    n = a[3]
    lo, hi = n // 3, n - n // 4
    a = gko.gen.stencil_csr(kind, nx, ny, nz, row_begin=lo, row_end=hi)
    b = oracle.gen_stencil_csr(kind, nx, ny, nz, row_begin=lo, row_end=hi)
    for x, y in zip(a[:3], b[:3]):
        assert np.array_equal(x, y)
