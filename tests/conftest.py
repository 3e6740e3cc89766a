import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


@pytest.fixture(scope="session")
def gko():
    """The product package (ctypes over libgko_b200.so)."""
    import __graft_entry__ as g
    return g.load_package()


@pytest.fixture(scope="session")
def ora():
    import oracle
    oracle.lib()
    return oracle


@pytest.fixture(scope="session")
def refimpl():
    """The compiled, unmodified reference (oracle/_ref); skip when it was not built."""
    import oracle
    if oracle.ref() is None:
        pytest.skip("oracle/_ref not built (needs /root/reference: make -f oracle/Makefile.ref)")
    return oracle


@pytest.fixture(scope="session")
def exec_(gko):
    import torch
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    return gko.CudaExecutor.create(0)
