"""Single-GPU CG on one configs[3] slab (7-pt 512x512x64): a short run for an ncu launch list."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from __graft_entry__ import load_package
gko = load_package()
exec_ = gko.CudaExecutor.create(0)
g = int(os.environ.get("AB_GRID", "512"))
iters = int(os.environ.get("AB_ITERS", "12"))
rp, ci, va, n = gko.gen.stencil_csr("7pt", g, g, g // 8)
A = gko.matrix.Csr.from_arrays(exec_, (n, n), rp, ci, va)
s = gko.solver.Cg.build().with_criteria(gko.stop.Iteration(iters)).with_check_every(iters).on(exec_).generate(A)
b = gko.matrix.Dense.create(exec_, (n, 1)); b.fill(1.0)
x = gko.matrix.Dense.create(exec_, (n, 1))
p, q = gko.matrix.Dense.create(exec_, (n, 1)), gko.matrix.Dense.create(exec_, (n, 1))
p.t.copy_(torch.randn(n, 1, dtype=torch.float64, device=exec_.device))
for _ in range(3):
    A.apply(p, q)
for _ in range(2):
    x.fill(0.0); s.apply(b, x)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(5):
    x.fill(0.0); s.apply(b, x)
e1.record(); torch.cuda.synchronize()
print("cg us/iter", e0.elapsed_time(e1) * 1e3 / (5 * iters), "iters", s.num_iterations)
