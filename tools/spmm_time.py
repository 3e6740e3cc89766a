"""32-RHS SpMM on the 27-pt stencil (configs[4] shape by default): time per apply, CSR and ELL.
usage: spmm_time.py [grid=256] [fp32|fp64] [nrhs=32] [formats=csr,ell]"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch
from __graft_entry__ import load_package
gko = load_package()
g = int(sys.argv[1]) if len(sys.argv) > 1 else 256
fp32 = (sys.argv[2] if len(sys.argv) > 2 else "fp32") == "fp32"
nrhs = int(sys.argv[3]) if len(sys.argv) > 3 else 32
formats = (sys.argv[4] if len(sys.argv) > 4 else "csr,ell").split(",")
exec_ = gko.CudaExecutor.create(0)
dt, tdt = (np.float32, torch.float32) if fp32 else (np.float64, torch.float64)
rp, ci, va, n = gko.gen.stencil_csr("27pt", g, g, g, value_dtype=dt)
csr = gko.matrix.Csr.from_arrays(exec_, (n, n), rp, ci, va)
x = gko.matrix.Dense.create(exec_, (n, nrhs), tdt); y = gko.matrix.Dense.create(exec_, (n, nrhs), tdt)
x.t.copy_(torch.randn(n, nrhs, dtype=tdt, device=exec_.device))
for f in formats:
    A = csr if f == "csr" else csr.convert_to(f)
    for _ in range(3):
        A.apply(x, y)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10):
        A.apply(x, y)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 10
    print(f"{f} 27pt {g}^3 {'fp32' if fp32 else 'fp64'} nrhs={nrhs}: {ms:.3f} ms  {A.spmv_bytes(nrhs)/ms/1e6:.0f} GB/s", flush=True)
