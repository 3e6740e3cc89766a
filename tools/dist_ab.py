"""A/B timing of the distributed SpMV / CG on the configs[3] slab (torchrun, one rank per GPU):
where the distance between the distributed iteration and the same local block without any
exchange goes.  GKOB200_DIST_DEBUG switches are measurement-only (wrong results)."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

from __graft_entry__ import load_package  # noqa: E402
from bench_dist import build_matrix  # noqa: E402


def main():
    rank, world, lr = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(lr)
    dist.init_process_group("nccl", device_id=torch.device("cuda", lr))
    gko = load_package()
    D = gko.distributed
    exec_ = gko.CudaExecutor.create(lr)
    comm = D.Communicator.from_torch(exec_)
    g = int(os.environ.get("AB_GRID", "512"))
    A, _, lo, hi = build_matrix(gko, exec_, comm, "7pt", g, g, (g // 8) * world, rank, world)
    n = hi - lo
    p, q = gko.matrix.Dense.create(exec_, (n, 1)), gko.matrix.Dense.create(exec_, (n, 1))
    p.t.copy_(torch.randn(n, 1, dtype=torch.float64, device=exec_.device))
    b = gko.matrix.Dense.create(exec_, (n, 1))
    b.fill(1.0)
    x = gko.matrix.Dense.create(exec_, (n, 1))
    iters = 100

    def timed(fn, reps):
        for _ in range(3):
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

    out = {"world": world, "rows_per_gpu": n, "fused": A.uses_fused_halo}
    out["spmv_local_alone_us"] = timed(lambda: A.local.apply(p, q), 50)
    for rep in range(3):
        for dbg, name in ((0, "spmv_dist_us"), (1, "spmv_dist_no_epoch_us"), (1 + 4, "spmv_dist_no_epoch_no_push_us"),
                          (1 + 8, "spmv_dist_no_epoch_no_tail_us"), (1 + 4 + 8, "spmv_dist_no_epoch_no_push_no_tail_us")):
            os.environ["GKOB200_DIST_DEBUG"] = str(dbg)
            out.setdefault(name, []).append(round(timed(lambda: A.apply(p, q), 50), 1))
        out.setdefault("spmv_local_alone_again_us", []).append(round(timed(lambda: A.local.apply(p, q), 50), 1))
    s_loc = gko.solver.Cg.build().with_criteria(gko.stop.Iteration(iters)).with_check_every(iters).on(exec_).generate(A.local)

    def loc():
        x.fill(0.0)
        s_loc.apply(b, x)
    out["cg_local_alone_us_per_iter"] = timed(loc, 5) / iters
    for dbg, name in ((0, "cg_dist_us_per_iter"), (1, "cg_dist_no_epoch_us_per_iter"), (2, "cg_dist_no_allreduce_us_per_iter"),
                      (3, "cg_dist_no_epoch_no_allreduce_us_per_iter")):
        os.environ["GKOB200_DIST_DEBUG"] = str(dbg)
        s = D.cg(exec_, A, [gko.stop.Iteration(iters)], check_every=iters)

        def run():
            x.fill(0.0)
            s.apply(b, x)
        out[name] = timed(run, 5) / iters
        del s
    os.environ["GKOB200_DIST_DEBUG"] = "0"
    if rank == 0:
        print("AB " + json.dumps(out), flush=True)
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
