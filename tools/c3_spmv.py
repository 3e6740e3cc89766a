"""C3 power-law merge-path SpMV alone (for ncu)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from __graft_entry__ import load_package
gko = load_package()
exec_ = gko.CudaExecutor.create(0)
n = int(os.environ.get("C3_ROWS", "10000000"))
rp, ci, va = gko.gen.powerlaw_csr(n)
A = gko.matrix.Csr.from_arrays(exec_, (n, n), rp, ci, va)
x = gko.matrix.Dense.create(exec_, (n, 1)); y = gko.matrix.Dense.create(exec_, (n, 1))
x.t.copy_(torch.randn(n, 1, dtype=torch.float64, device=exec_.device))
for _ in range(int(os.environ.get("C3_REPS", "6"))):
    A.apply(x, y)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(10):
    A.apply(x, y)
e1.record(); torch.cuda.synchronize()
print("kernel", A.kernel(), "us", e0.elapsed_time(e1) * 100, "GB/s", A.spmv_bytes(1) / (e0.elapsed_time(e1) * 1e-4) / 1e9)
