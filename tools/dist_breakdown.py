#!/usr/bin/env python
"""torchrun tool: times the pieces of one distributed CG iteration (CUDA events, max over
ranks): local SpMV, non-local SpMV, full distributed apply, small all-reduce, CG iteration."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np  # noqa: E402
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

from __graft_entry__ import load_package  # noqa: E402

rank, world, lr = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(lr)
dist.init_process_group("nccl", device_id=torch.device("cuda", lr))
gko = load_package()
D = gko.distributed
exec_ = gko.CudaExecutor.create(lr)
comm = D.Communicator.from_torch(exec_)
g = int(sys.argv[1]) if len(sys.argv) > 1 else 200
part = D.Partition.build_from_global_size_uniform(exec_, world, g * g * g * world)
lo, hi = int(part.range_bounds[rank].item()), int(part.range_bounds[rank + 1].item())
rp, ci, va, _ = gko.gen.stencil_csr("27pt", g, g, g * world, row_begin=lo, row_end=hi, index_dtype=np.int64)
rows = np.repeat(np.arange(lo, hi, dtype=np.int64), np.diff(rp))
A = D.Matrix(exec_, comm).read_distributed(rows, ci, va, part)
n = hi - lo
p, q = gko.matrix.Dense.create(exec_, (n, 1)), gko.matrix.Dense.create(exec_, (n, 1))
p.t.copy_(torch.randn(n, 1, dtype=torch.float64, device=exec_.device))
ghost = gko.matrix.Dense.create(exec_, (max(A.non_local.size[1], 1), 1))


def timed(fn, reps=30):
    for _ in range(5):
        fn()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    dist.barrier()
    torch.cuda.synchronize()
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    t = torch.tensor([e0.elapsed_time(e1) * 1e3 / reps], dtype=torch.float64, device=exec_.device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


small = torch.zeros(2, dtype=torch.float64, device=exec_.device)
one = gko.matrix.Dense.scalar(exec_, 1.0)
res = {
    "local_spmv_us": timed(lambda: A.local.apply(p, q)),
    "nonlocal_spmv_us": timed(lambda: A.non_local.apply(one, ghost, one, q)) if A.non_local.nnz else 0.0,
    "dist_apply_us": timed(lambda: A.apply(p, q)),
    "allreduce2_us": timed(lambda: comm.all_reduce_sum(small), 200),
}
iters = 100
jac = gko.preconditioner.Jacobi.build().with_max_block_size(1).on(exec_).generate(A.local)
s = D.cg(exec_, A, [gko.stop.Iteration(iters)], precond=jac, check_every=iters)
b = gko.matrix.Dense.create(exec_, (n, 1))
b.fill(1.0)
x = gko.matrix.Dense.create(exec_, (n, 1))


def solve():
    x.fill(0.0)
    s.apply(b, x)


res["cg_iteration_us"] = timed(solve, 3) / iters
res["launches_per_iteration"] = s.launch_count / iters
res["p2p_allreduce"] = comm.uses_p2p
if rank == 0:
    print("BREAKDOWN", world, res, flush=True)
dist.destroy_process_group()
