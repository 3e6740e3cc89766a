"""The drop-in, end to end: shim/test_dropin.cpp is an ordinary Ginkgo program (public API
only) linked against the UNMODIFIED Ginkgo core built from /root/reference (oracle/_ref/lib)
and against shim/_build/libginkgo_cuda.so — the B200 shim over libgko_b200.so.  It runs
Csr/Ell/Sellp/Coo/Hybrid::apply, Csr::transpose / sort_by_column_index and
Cg/Fcg/Cgs/Bicgstab/Gmres (+ scalar / block Jacobi) on
gko::CudaExecutor and compares with gko::ReferenceExecutor in the same process."""
import os
import subprocess

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_ginkgo_program_on_cuda_executor_matches_reference_executor():
    exe = os.path.join(ROOT, "shim", "_build", "test_dropin")
    if not os.path.exists(exe):
        pytest.skip("shim not built (needs /root/reference at build time: make -C shim)")
    env = dict(os.environ)
    env["LD_LIBRARY_PATH"] = ":".join([os.path.join(ROOT, "shim", "_build"), os.path.join(ROOT, "oracle", "_ref", "lib"),
                                       os.path.join(ROOT, "repo-8852-ginkgo_b200"), env.get("LD_LIBRARY_PATH", "")])
    out = subprocess.run([exe], capture_output=True, text=True, env=env, timeout=600)
    print(out.stdout[-4000:], out.stderr[-2000:])
    assert out.returncode == 0 and "DROPIN_OK" in out.stdout, out.stdout[-3000:] + out.stderr[-2000:]
    assert "b200-native" in out.stdout   # the shim's libginkgo_cuda.so is the one that was loaded
