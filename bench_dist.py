"""bench.py, N > 1: BASELINE configs[3] — distributed CG (no preconditioner) on the 3D 7-pt
stencil 512^3, row-partitioned in z over the GPUs of one box, through the distributed matrix
(halo exchange inside the SpMV launch over peer memory; NCCL fallback).

  --scaling weak    every GPU owns a grid x grid x (grid/8) slab: 512 x 512 x 64N rows (N = 8 is 512^3)
  --scaling strong  the grid^3 problem is fixed and split over the N GPUs

Launched by torchrun, one rank per GPU.  Outside the timed region every run PROVES the path it
times (reference's own check: test/mpi/distributed/matrix.cpp:257-281, 406-445):
  (a) full size: A.apply on every rank == the rows of the global stencil product computed with
      numpy in the reference's summation order (local block, then non-local block) — bit for bit;
  (b) reduced size (same stencil, same partitioning, all ranks): a tolerance solve
      (ResidualNorm 1e-8) on each exchange path (fused halo / peer all-reduce + NCCL halo / NCCL
      only) against a single-GPU solve of the same global system on rank 0: iteration count
      (+-2), first 10 residual norms (1e-10), solution (1e-8);
  (c) --verify-full: the first residual norms of the full-size distributed run against a
      single-GPU solve of the full global system on rank 0.
A mismatch fails the run (exit 1, no JSON line)."""
from __future__ import annotations

import json
import os
import sys

import numpy as np

STENCIL_DIAG = {"7pt": 6.0, "27pt": 26.0}


def stencil_offsets(kind):
    """(dz, dy, dx) in ascending column order."""
    offs = []
    for dz in (-1, 0, 1):
        for dy in (-1, 0, 1):
            for dx in (-1, 0, 1):
                if kind == "7pt" and abs(dx) + abs(dy) + abs(dz) > 1:
                    continue
                offs.append((dz, dy, dx))
    return offs


def test_vector(idx):
    """deterministic pseudo-random entries in [-0.5, 0.5), exact in fp64, a function of the GLOBAL index"""
    return ((idx * 2654435761) % (1 << 20)).astype(np.float64) / float(1 << 20) - 0.5


def expected_rows(kind, nx, ny, nz, lo, hi, chunk=1 << 22):
    """Rows [lo, hi) of A x (x = test_vector) for the global stencil matrix, summed the way
    distributed::Matrix::apply does (core/distributed/matrix.cpp:312-333): the local block's
    entries in column order first, then the non-local block's entries (ghost columns sorted by
    owner, then global index) — rounded product, rounded sum, like the reference executor."""
    out = np.empty(hi - lo)
    plane = nx * ny
    for c0 in range(lo, hi, chunk):
        c1 = min(c0 + chunk, hi)
        rows = np.arange(c0, c1, dtype=np.int64)
        x, y, z = rows % nx, (rows // nx) % ny, rows // plane
        acc = np.zeros(c1 - c0)
        for local_pass in (True, False):
            for dz, dy, dx in stencil_offsets(kind):
                xx, yy, zz = x + dx, y + dy, z + dz
                ok = (xx >= 0) & (xx < nx) & (yy >= 0) & (yy < ny) & (zz >= 0) & (zz < nz)
                col = (zz * ny + yy) * nx + xx
                is_local = (col >= lo) & (col < hi)
                m = ok & (is_local if local_pass else ~is_local)
                if not m.any():
                    continue
                val = STENCIL_DIAG[kind] if (dz, dy, dx) == (0, 0, 0) else -1.0
                term = val * test_vector(np.where(m, col, 0))
                acc = np.where(m, acc + term, acc)
        out[c0 - lo:c1 - lo] = acc
    return out


def build_matrix(gko, exec_, comm, kind, nx, ny, nz, rank, world, env=None):
    """read_distributed of this rank's rows of the global stencil; `env` toggles the exchange path"""
    D = gko.distributed
    saved = {}
    for k, v in (env or {}).items():
        saved[k] = os.environ.get(k)
        os.environ[k] = v
    try:
        c = comm if env is None else D.Communicator.from_torch(exec_)
        n_global = nx * ny * nz
        part = D.Partition.build_from_global_size_uniform(exec_, world, n_global)
        lo, hi = int(part.range_bounds[rank].item()), int(part.range_bounds[rank + 1].item())
        rp, ci, va, _ = gko.gen.stencil_csr(kind, nx, ny, nz, row_begin=lo, row_end=hi, index_dtype=np.int64)
        rows = np.repeat(np.arange(lo, hi, dtype=np.int64), np.diff(rp))
        A = D.Matrix(exec_, c).read_distributed(rows, ci, va, part)
    finally:
        for k, v in saved.items():
            if v is None:
                os.environ.pop(k, None)
            else:
                os.environ[k] = v
    return A, c, lo, hi


def verify_reduced(gko, exec_, kind, g, planes, rank, world, dist, torch, log):
    """(b): tolerance solves on every exchange path vs a single-GPU solve of the same system."""
    D = gko.distributed
    nx = ny = max(g // 4, 8)
    nz = max(planes // 4, 2) * world
    n = nx * ny * nz
    crit = lambda: [gko.stop.Iteration(3000), gko.stop.ResidualNorm(1e-8)]  # noqa: E731
    # single-GPU solve of the GLOBAL reduced system on rank 0, broadcast to everybody
    head = torch.zeros(12, dtype=torch.float64, device=exec_.device)
    xg = torch.zeros(n, dtype=torch.float64, device=exec_.device)
    if rank == 0:
        rp, ci, va, _ = gko.gen.stencil_csr(kind, nx, ny, nz)
        A1 = gko.matrix.Csr.from_arrays(exec_, (n, n), rp, ci, va)
        s1 = gko.solver.Cg.build().with_criteria(*crit()).on(exec_).generate(A1)
        b1 = gko.matrix.Dense.create(exec_, (n, 1))
        b1.fill(1.0)
        x1 = gko.matrix.Dense.create(exec_, (n, 1))
        s1.apply(b1, x1)
        hist = np.asarray(s1.residual_history[:10], dtype=np.float64)
        head[0], head[1] = s1.num_iterations, len(hist)
        head[2:2 + len(hist)] = torch.from_numpy(hist).to(exec_.device)
        xg.copy_(x1.t[:, 0])
    dist.broadcast(head, src=0)
    dist.broadcast(xg, src=0)
    it_ref, nh = int(head[0].item()), int(head[1].item())
    hist_ref = head[2:2 + nh].cpu().numpy()
    result = {"system": f"{kind} {nx}x{ny}x{nz}, ResidualNorm 1e-8, b = 1", "single_gpu_iterations": it_ref, "paths": {}}
    ok = True
    paths = {"fused-halo": {"GKOB200_P2P": "1", "GKOB200_FUSED_HALO": "1"},
             "p2p-allreduce+nccl-halo": {"GKOB200_P2P": "1", "GKOB200_FUSED_HALO": "0"},
             "nccl": {"GKOB200_P2P": "0", "GKOB200_FUSED_HALO": "0"}}
    for name, env in paths.items():
        A, c, lo, hi = build_matrix(gko, exec_, None, kind, nx, ny, nz, rank, world, env=env)
        ran = "fused-halo" if A.uses_fused_halo else ("p2p-allreduce+nccl-halo" if c.uses_p2p else "nccl")
        s = D.cg(exec_, A, crit(), precond=None, check_every=8)
        b = gko.matrix.Dense.create(exec_, (hi - lo, 1))
        b.fill(1.0)
        x = gko.matrix.Dense.create(exec_, (hi - lo, 1))
        s.apply(b, x)
        first = (int(s.num_iterations), np.asarray(s.residual_history[:nh], dtype=np.float64))
        x.fill(0.0)
        s.apply(b, x)      # a second solve on the same solver object must repeat the first one
        hist = np.asarray(s.residual_history[:nh], dtype=np.float64)
        repeat_ok = first[0] == int(s.num_iterations) and np.array_equal(first[1], hist)
        hist_err = float(np.max(np.abs(hist - hist_ref[:len(hist)]) / hist_ref[:len(hist)])) if len(hist) else 0.0
        x_err = float((x.t[:, 0] - xg[lo:hi]).abs().max().item() / xg.abs().max().item())
        bad = (abs(s.num_iterations - it_ref) > 2) or len(hist) != nh or hist_err > 1e-10 or x_err > 1e-8 or not repeat_ok
        flags = torch.tensor([float(bad), hist_err, x_err], dtype=torch.float64, device=exec_.device)
        dist.all_reduce(flags, op=dist.ReduceOp.MAX)
        result["paths"][name] = {"ran": ran, "iterations": int(s.num_iterations), "max_rel_err_first_10_residual_norms":
                                 float(flags[1].item()), "max_rel_err_solution": float(flags[2].item()),
                                 "ok": flags[0].item() == 0.0}
        ok = ok and flags[0].item() == 0.0
        del s, A, c
    result["ok"] = ok
    log(f"[bench_dist] reduced-size parity: {json.dumps(result)}")
    return result


def run_distributed(args, gko, rank, world, local_rank):
    import torch
    import torch.distributed as dist
    from bench import METRIC, UNIT, ClockSampler, cg_model_bytes, peaks
    # NCCL prints its version banner on stdout; the contract is ONE JSON line on stdout.
    # Everything goes to stderr until rank 0 writes the result line to the saved descriptor.
    sys.stdout.flush()
    saved_stdout = os.dup(1)
    os.dup2(2, 1)

    def log(msg):
        if rank == 0:
            print(msg, file=sys.stderr, flush=True)

    sampler = ClockSampler(local_rank)
    sampler.start()          # nvidia-smi needs ~1 s to come up: start it before the set-up
    torch.cuda.set_device(local_rank)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    D = gko.distributed
    exec_ = gko.CudaExecutor.create(local_rank)
    comm = D.Communicator.from_torch(exec_)
    kind = args.stencil or "7pt"
    g = args.grid or 512
    planes_weak = args.slab_planes or max(g // 8, 1)
    if args.scaling == "weak":
        nz = planes_weak * world
    else:
        nz = g
        if nz % world:
            raise SystemExit(f"strong scaling needs grid ({g}) divisible by the number of GPUs ({world})")
    nx = ny = g
    n_global = nx * ny * nz
    A, _, lo, hi = build_matrix(gko, exec_, comm, kind, nx, ny, nz, rank, world)
    n = hi - lo
    nnz_local = A.local.nnz + A.non_local.nnz
    path = "fused-halo" if A.uses_fused_halo else ("p2p-allreduce+nccl-halo" if comm.uses_p2p else "nccl")
    log(f"[bench_dist] {kind} {nx}x{ny}x{nz} over {world} GPUs ({args.scaling}), {n} rows/GPU, path={path}")

    # ---- (a) full-size apply parity on every rank -------------------------------------------
    parity = {"exchange_path": path}
    if not args.no_verify:
        xv = gko.matrix.Dense.from_numpy(exec_, test_vector(np.arange(lo, hi, dtype=np.int64)))
        yv = gko.matrix.Dense.create(exec_, (n, 1))
        A.apply(xv, yv)
        A.apply(xv, yv)      # twice: both halves of the double-buffered receive window
        got = yv.to_numpy()[:, 0]
        want = expected_rows(kind, nx, ny, nz, lo, hi)
        bound = 2.0 * STENCIL_DIAG[kind] * 0.5
        flags = torch.tensor([float(not np.array_equal(got, want)), float(np.abs(got - want).max() / bound)],
                             dtype=torch.float64, device=exec_.device)
        dist.all_reduce(flags, op=dist.ReduceOp.MAX)
        parity["apply_full_size"] = {"rows_checked": n_global, "bit_identical_on_every_rank": flags[0].item() == 0.0,
                                     "max_abs_err_over_sum_abs": float(flags[1].item()),
                                     "expected": "numpy, global stencil, reference summation order"}
        log(f"[bench_dist] full-size apply parity: {json.dumps(parity['apply_full_size'])}")
        if flags[1].item() > 1e-12:
            raise SystemExit("distributed apply differs from the global stencil product")
        del xv, yv, got, want
        parity["cg_reduced_size"] = verify_reduced(gko, exec_, kind, g, n // (nx * ny), rank, world, dist, torch, log)
        if not parity["cg_reduced_size"]["ok"]:
            raise SystemExit("distributed CG differs from the single-GPU solve of the same system")

    iters = args.iters_per_step
    jac = None
    if args.dist_precond == "jacobi":
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

    step_device()
    hist_first = np.asarray(solver.residual_history[:10], dtype=np.float64)
    for _ in range(max(args.warmup, 3) - 1):
        step_device()
    sampler.mark()
    secs, launches = timed(step_device, args.steps)
    clocks = sampler.stop()
    assert solver.num_iterations == iters
    hist_head = [float(v) for v in solver.residual_history[:10]]
    # every timed solve is the same solve: its residual norms must repeat the first one's and move
    if not np.array_equal(hist_first, np.asarray(hist_head)) or (len(hist_head) > 2 and hist_head[1] == hist_head[0]):
        raise SystemExit(f"timed solves do not repeat the first solve: {hist_first} vs {hist_head}")
    slabs = world if args.scaling == "weak" else 1
    value = slabs * args.steps * iters / secs
    for _ in range(2):
        step_host()
    secs_e2e, _ = timed(step_host, args.steps)

    # ---- (c) optional: full-size residual norms vs a single-GPU solve of the global system ---
    if args.verify_full:
        head = torch.zeros(10, dtype=torch.float64, device=exec_.device)
        if rank == 0:
            del_keep = (db, dx)  # noqa: F841
            rp, ci, va, _ = gko.gen.stencil_csr(kind, nx, ny, nz)
            A1 = gko.matrix.Csr.from_arrays(exec_, (n_global, n_global), rp, ci, va)
            del rp, ci, va
            s1 = gko.solver.Cg.build().with_criteria(gko.stop.Iteration(12)).on(exec_).generate(A1)
            b1 = gko.matrix.Dense.create(exec_, (n_global, 1))
            b1.fill(1.0)
            x1 = gko.matrix.Dense.create(exec_, (n_global, 1))
            s1.apply(b1, x1)
            head.copy_(torch.tensor(s1.residual_history[:10], dtype=torch.float64))
            del s1, A1, b1, x1
        dist.broadcast(head, src=0)
        ref_head = head.cpu().numpy()
        err = float(np.max(np.abs(np.asarray(hist_head) - ref_head) / ref_head))
        parity["cg_full_size"] = {"first_10_residual_norms_vs_single_gpu_global_solve_max_rel_err": err, "ok": err < 1e-9}
        log(f"[bench_dist] full-size CG parity: {json.dumps(parity['cg_full_size'])}")
        if err >= 1e-9:
            raise SystemExit("full-size distributed residual norms differ from the single-GPU global solve")

    # ---- the same local block without any exchange (the denominator weak scaling is judged by)
    s_loc = (gko.solver.Cg.build().with_criteria(gko.stop.Iteration(iters)).with_check_every(max(iters, 1))
             .on(exec_).generate(A.local))
    if jac is not None:
        s_loc = (gko.solver.Cg.build().with_criteria(gko.stop.Iteration(iters)).with_generated_preconditioner(jac)
                 .with_check_every(max(iters, 1)).on(exec_).generate(A.local))

    def step_local():
        dx.fill(0.0)
        s_loc.apply(db, dx)
        return 0

    for _ in range(3):
        step_local()
    secs_loc, _ = timed(step_local, args.steps)
    del s_loc

    # dominant kernel on this rank: the distributed SpMV (halo push + local + non-local rows)
    p, q = gko.matrix.Dense.create(exec_, (n, 1)), gko.matrix.Dense.create(exec_, (n, 1))
    p.t.copy_(torch.randn(n, 1, dtype=torch.float64, device=exec_.device))
    for _ in range(5):
        A.apply(p, q)
    spmv_s, _ = timed(lambda: (A.apply(p, q), 1)[1], args.spmv_reps)
    spmv_s /= args.spmv_reps
    for _ in range(5):
        A.local.apply(p, q)
    spmv_loc_s, _ = timed(lambda: (A.local.apply(p, q), 1)[1], args.spmv_reps)
    spmv_loc_s /= args.spmv_reps
    peak, peak_src = peaks()
    spmv_bytes = A.spmv_bytes(1)
    achieved = spmv_bytes / spmv_s / 1e9
    it_bytes = cg_model_bytes(n, nnz_local, jac.storage_bytes() if jac is not None else 0)
    if rank == 0:
        shape = (f"{nx}x{ny}x{nz} ({n_global} rows), "
                 + (f"weak: one {nx}x{ny}x{planes_weak} slab per GPU" if args.scaling == "weak"
                    else f"strong: {nx}^3 fixed, {nz // world} planes per GPU"))
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": max(args.warmup, 3), "ms_per_step": 1e3 * secs / args.steps, "higher_is_better": True,
            "scaling": args.scaling, "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": f"distributed CG ({'scalar Jacobi' if jac is not None else 'no preconditioner'}), "
                                   f"3D {kind} stencil {shape}, row-partitioned in z over {world} GPUs "
                                   "(BASELINE configs[3]), fp64, int32 local / int64 global indices, "
                                   f"exchange path: {path}",
                       "rows": n_global, "rows_per_gpu": n, "nnz_per_gpu": nnz_local, "iters_per_step": iters,
                       "halo_values_per_gpu": int(A.recv_sizes.sum()),
                       "l2": "per-GPU operands (1.4 GB matrix + vectors) exceed the 126 MB L2; no flush needed",
                       "value_definition": ("CG iterations/s x N slabs" if args.scaling == "weak"
                                            else "CG iterations/s of the fixed global problem"),
                       "n1_basis": "the N=1 line of bench.py is BASELINE configs[1] (another matrix); the weak-scaling "
                                   "denominator for THIS workload is `no_exchange_baseline` below and "
                                   "`configs.C4_slab` of the N=1 line"},
            "e2e": {"value": slabs * args.steps * iters / secs_e2e, "unit": UNIT, "h2d_bytes_per_step": 2 * n * 8 * world,
                    "d2h_bytes_per_step": (n * 8 + 16) * world},
            "gpu_launches": launches, "clocks": clocks,
            "roofline": {"bound": "hbm", "kernel": "distributed csr_spmv (halo push + local + non-local rows, per GPU)",
                         "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                         "peak_source": peak_src, "bytes_per_launch": spmv_bytes, "us_per_launch": spmv_s * 1e6,
                         "us_per_launch_local_block_alone": spmv_loc_s * 1e6, "traffic": None},
            "cg_iteration": {"us": 1e6 * secs / (args.steps * iters), "model_bytes_per_gpu": it_bytes,
                             "model_gbs_per_gpu": it_bytes * args.steps * iters / secs / 1e9},
            "no_exchange_baseline": {"what": "single-GPU CG on the same local block (no halo, no all-reduce), all GPUs "
                                             "at once, max over ranks", "us_per_iteration": 1e6 * secs_loc / (args.steps * iters),
                                     "value": slabs * args.steps * iters / secs_loc, "unit": UNIT},
            "residual_norms_head": hist_head,
            "parity": parity,
        }
        os.write(saved_stdout, (json.dumps(line) + "\n").encode())
    dist.barrier()
    dist.destroy_process_group()
