"""bench.py, N > 1: distributed CG + scalar Jacobi, weak scaling in z (each GPU owns a
grid^3 slab of the 27-pt stencil on grid x grid x grid*N), row-partitioned
distributed::Matrix with the NCCL halo exchange.  Launched by torchrun, one rank per GPU."""
from __future__ import annotations

import json
import os

import numpy as np


def run_distributed(args, gko, rank, world, local_rank):
    import sys
    import torch
    import torch.distributed as dist
    from bench import METRIC, UNIT, ClockSampler, cg_model_bytes, peaks
    # NCCL prints its version banner on stdout; the contract is ONE JSON line on stdout.
    # Everything goes to stderr until rank 0 writes the result line to the saved descriptor.
    sys.stdout.flush()
    saved_stdout = os.dup(1)
    os.dup2(2, 1)
    sampler = ClockSampler(local_rank)
    sampler.start()          # nvidia-smi needs ~1 s to come up: start it before the set-up
    torch.cuda.set_device(local_rank)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    D = gko.distributed
    exec_ = gko.CudaExecutor.create(local_rank)
    comm = D.Communicator.from_torch(exec_)
    g = args.grid
    n_global = g * g * g * world
    part = D.Partition.build_from_global_size_uniform(exec_, world, n_global)
    lo, hi = int(part.range_bounds[rank].item()), int(part.range_bounds[rank + 1].item())
    rp, ci, va, _ = gko.gen.stencil_csr("27pt", g, g, g * world, row_begin=lo, row_end=hi, index_dtype=np.int64)
    rows = np.repeat(np.arange(lo, hi, dtype=np.int64), np.diff(rp))
    A = D.Matrix(exec_, comm).read_distributed(rows, ci, va, part)
    del rows, ci, va
    n = hi - lo
    nnz_local = A.local.nnz + A.non_local.nnz
    iters = args.iters_per_step
    jac = gko.preconditioner.Jacobi.build().with_max_block_size(1).on(exec_).generate(A.local)
    solver = D.cg(exec_, A, [gko.stop.Iteration(iters)], precond=jac, check_every=max(iters, 1))
    b_host = torch.ones(n, dtype=torch.float64).pin_memory()
    x_host = torch.zeros(n, dtype=torch.float64).pin_memory()
    db, dx = gko.matrix.Dense.create(exec_, (n, 1)), gko.matrix.Dense.create(exec_, (n, 1))
    db.t.copy_(b_host.view(n, 1))

    def step_device():
        dx.fill(0.0)
        solver.apply(db, dx)
        return solver.launch_count + 1

    def step_host():
        x_host.zero_()
        solver.apply_host(b_host, x_host)
        return solver.launch_count

    def timed(fn, steps):
        launches = 0
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        dist.barrier()
        torch.cuda.synchronize()
        e0.record()
        for _ in range(steps):
            launches += fn()
        e1.record()
        torch.cuda.synchronize()
        t = torch.tensor([e0.elapsed_time(e1) * 1e-3], dtype=torch.float64, device=exec_.device)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)   # max over ranks
        dist.barrier()
        return float(t.item()), launches

    for _ in range(max(args.warmup, 3)):
        step_device()
    sampler.mark()
    secs, launches = timed(step_device, args.steps)
    clocks = sampler.stop()
    assert solver.num_iterations == iters
    value = world * args.steps * iters / secs
    for _ in range(2):
        step_host()
    secs_e2e, _ = timed(step_host, args.steps)
    # dominant kernel on this rank: the distributed SpMV (local + non-local + halo)
    p, q = gko.matrix.Dense.create(exec_, (n, 1)), gko.matrix.Dense.create(exec_, (n, 1))
    p.t.copy_(torch.randn(n, 1, dtype=torch.float64, device=exec_.device))
    for _ in range(5):
        A.apply(p, q)
    spmv_s, _ = timed(lambda: (A.apply(p, q), 1)[1], args.spmv_reps)
    spmv_s /= args.spmv_reps
    peak, peak_src = peaks()
    spmv_bytes = A.spmv_bytes(1)
    achieved = spmv_bytes / spmv_s / 1e9
    it_bytes = cg_model_bytes(n, nnz_local, jac.storage_bytes())
    if rank == 0:
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": max(args.warmup, 3), "ms_per_step": 1e3 * secs / args.steps, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": f"distributed CG + scalar Jacobi, 3D 27-pt stencil {g}x{g}x{g * world} "
                                   f"row-partitioned in z over {world} GPUs (BASELINE configs[1] slab per GPU), "
                                   "fp64/int32 local CSR, NCCL halo exchange",
                       "rows": n_global, "rows_per_gpu": n, "nnz_per_gpu": nnz_local, "iters_per_step": iters,
                       "halo_values_per_gpu": int(A.recv_sizes.sum()),
                       "l2": "per-GPU operands (2.7 GB) exceed the 126 MB L2; no flush needed",
                       "value_definition": "CG iterations/s x N slabs (global rows x iterations / s / 8e6)"},
            "e2e": {"value": world * args.steps * iters / secs_e2e, "unit": UNIT, "h2d_bytes_per_step": 2 * n * 8 * world,
                    "d2h_bytes_per_step": (n * 8 + 16) * world},
            "gpu_launches": launches, "clocks": clocks,
            "roofline": {"bound": "hbm", "kernel": "distributed csr_spmv (local + non-local, per GPU)",
                         "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                         "peak_source": peak_src, "bytes_per_launch": spmv_bytes, "us_per_launch": spmv_s * 1e6,
                         "traffic": None},
            "cg_iteration": {"us": 1e6 * secs / (args.steps * iters), "model_bytes_per_gpu": it_bytes,
                             "model_gbs_per_gpu": it_bytes * args.steps * iters / secs / 1e9},
        }
        os.write(saved_stdout, (json.dumps(line) + "\n").encode())
    dist.barrier()
    dist.destroy_process_group()
