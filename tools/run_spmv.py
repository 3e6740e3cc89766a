#!/usr/bin/env python
"""Runs a handful of SpMV launches on a BASELINE matrix — the target of the ncu captures
under profiles/ (keeps the profiled command short).  Usage:
   python tools/run_spmv.py [27pt|7pt|5pt|powerlaw] [grid] [format] [reps] [dtype] [nrhs]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np  # noqa: E402
import torch  # noqa: E402

from __graft_entry__ import load_package  # noqa: E402

kind = sys.argv[1] if len(sys.argv) > 1 else "27pt"
grid = int(sys.argv[2]) if len(sys.argv) > 2 else 200
fmt = sys.argv[3] if len(sys.argv) > 3 else "csr"
reps = int(sys.argv[4]) if len(sys.argv) > 4 else 5
dtype = np.float32 if len(sys.argv) > 5 and sys.argv[5] == "f32" else np.float64
nrhs = int(sys.argv[6]) if len(sys.argv) > 6 else 1
gko = load_package()
exec_ = gko.CudaExecutor.create(0)
if kind == "powerlaw":
    n = grid
    rp, ci, va = gko.gen.powerlaw_csr(n)
else:
    rp, ci, va, n = gko.gen.stencil_csr(kind, grid, grid, grid, value_dtype=dtype)
A = gko.matrix.Csr.from_arrays(exec_, (n, n), rp, ci, va.astype(dtype))
if fmt != "csr":
    A = A.convert_to(fmt)
tdt = torch.float64 if dtype == np.float64 else torch.float32
x = gko.matrix.Dense.create(exec_, (n, nrhs), tdt)
y = gko.matrix.Dense.create(exec_, (n, nrhs), tdt)
x.t.copy_(torch.randn(n, nrhs, dtype=tdt, device=exec_.device))
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
for _ in range(3):
    A.apply(x, y)
torch.cuda.synchronize()
e0.record()
for _ in range(reps):
    A.apply(x, y)
e1.record()
torch.cuda.synchronize()
us = e0.elapsed_time(e1) * 1e3 / reps
print(f"{kind} {grid} {fmt} nrhs={nrhs} kernel={getattr(A, 'kernel', lambda: fmt)()}: {us:.1f} us/launch, "
      f"{A.spmv_bytes(nrhs) / us / 1e3:.1f} GB/s algorithmic")
