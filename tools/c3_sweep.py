"""Merge-path SpMV on the power-law matrix at several sizes: ns per entry vs size of the gathered vector."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from __graft_entry__ import load_package
gko = load_package()
exec_ = gko.CudaExecutor.create(0)
for n in [int(a) for a in sys.argv[1:]] or (1_000_000, 2_000_000, 4_000_000, 7_000_000, 10_000_000):
    rp, ci, va = gko.gen.powerlaw_csr(n)
    A = gko.matrix.Csr.from_arrays(exec_, (n, n), rp, ci, va)
    x = gko.matrix.Dense.create(exec_, (n, 1)); y = gko.matrix.Dense.create(exec_, (n, 1))
    x.t.copy_(torch.randn(n, 1, dtype=torch.float64, device=exec_.device))
    for _ in range(5):
        A.apply(x, y)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(20):
        A.apply(x, y)
    e1.record(); torch.cuda.synchronize()
    us = e0.elapsed_time(e1) * 50
    print(f"n={n} nnz={len(ci)} x_MB={n*8/1e6:.0f} us={us:.1f} ps_per_entry={us*1e6/len(ci):.1f} GB/s={A.spmv_bytes(1)/us/1e3:.0f}", flush=True)
    del A, x, y
