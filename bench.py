#!/usr/bin/env python
"""bench.py — headline benchmark: CG iterations/s and CSR SpMV GB/s (fraction of the HBM
roofline) on BASELINE.json configs[1]: CG + scalar Jacobi on the 3D 27-pt stencil 200^3,
fp64 / int32 CSR, one B200; at N > 1 the same slab per GPU (200 x 200 x 200N grid,
row-partitioned in z, weak scaling) through the distributed matrix.

A "step" is one CG solve of --iters-per-step iterations (residual criterion off so every
step does exactly that many iterations; x reset to 0 each step).
  value  = CG iterations/s x N  (whole job: N slabs of 8M rows advance one iteration each),
           device-resident b, x; CUDA-event timed, max over ranks.
  e2e    = same through solver.apply_host(): pinned HOST b and x0 copied in, x copied back,
           inside the timed region.
  roofline = the dominant kernel (CSR SpMV, row-block kernel) timed alone with CUDA events
           on the launching stream: algorithmic bytes / average launch time vs the measured
           HBM copy peak (MEASURED_PEAKS.json).
  cpu_baseline = the UNMODIFIED reference (oracle/_ref, OpenMP executor, all host cores) on a
           bounded sample (same matrix, few iterations), rank 0 at N=1 only.
`--impl reference` times only that CPU reference and prints the same line shape.
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402

METRIC = "cg_iters_per_s"
UNIT = "iters/s"


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--grid", type=int, default=None,
                    help="stencil grid edge (default 200 at N=1 -> configs[1]; 512 at N>1 -> configs[3])")
    ap.add_argument("--stencil", default=None, choices=["7pt", "27pt"], help="N>1 only (default 7pt -> configs[3])")
    ap.add_argument("--scaling", default="weak", choices=["weak", "strong"],
                    help="N>1: weak = one grid x grid x (grid/8) slab per GPU, strong = grid^3 fixed")
    ap.add_argument("--slab-planes", type=int, default=None, help="weak scaling: planes per GPU (default grid/8)")
    ap.add_argument("--dist-precond", default="none", choices=["none", "jacobi"])
    ap.add_argument("--dist", action="store_true", help="run the distributed bench also at N=1 (strong-scaling base)")
    ap.add_argument("--no-verify", action="store_true", help="N>1: skip the parity checks outside the timed region")
    ap.add_argument("--verify-full", action="store_true",
                    help="N>1: also compare the full-size residual norms with a single-GPU solve of the global system")
    ap.add_argument("--iters-per-step", type=int, default=100)
    ap.add_argument("--ref-iters-per-step", type=int, default=4)
    ap.add_argument("--cpu-baseline-iters", type=int, default=8)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-config-block", action="store_true",
                    help="N=1: skip the per-config block (C1, C3, C4 slab, C5 SpMV roofline + solver iterations/s)")
    ap.add_argument("--small-config-block", action="store_true", help="N=1: tiny sizes in the per-config block (smoke)")
    ap.add_argument("--spmv-reps", type=int, default=50)
    ap.add_argument("--format", default="csr", choices=["csr", "sellp", "ell", "hybrid", "coo"])
    return ap.parse_args()


def peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        try:
            return float(json.load(open(path))["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region."""
    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index=0):
        self.index, self.rows, self.proc, self.t_mark = index, [], None, 0.0

    def mark(self):
        """Samples taken before this call (set-up, warm-up) are dropped."""
        self.t_mark = time.time()

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append((time.time(), [c.strip() for c in line.split(",")]))

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ts, r in self.rows:
            if len(r) < 6 or ts < self.t_mark:
                continue
            try:
                sm.append(float(r[0]))
                mx.append(float(r[1]))
            except ValueError:
                continue
            for nm, v in zip(names, r[2:6]):
                if v.lower().startswith("active"):
                    reasons.add(nm)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def cg_model_bytes(n, nnz, precond_bytes, v=8, i=4):
    """The reference's own traffic model: 18 n V + matrix + preconditioner (core/solver/cg.cpp:148-156)."""
    return 18 * n * v + nnz * (v + i) + (n + 1) * i + precond_bytes


def host_threads():
    try:
        return len(os.sched_getaffinity(0))
    except AttributeError:
        return os.cpu_count() or 1


def run_reference(args, rank, world):
    """The reference's own CPU implementation (Ginkgo OpenMP executor from oracle/_ref, all host
    threads) on this arm's workload.  N=1: configs[1] (CG + scalar Jacobi, 27-pt 200^3).  N>1:
    configs[3] — the reference's distributed classes need MPI (not in this image), so the CPU
    arm solves the same GLOBAL system (all N slabs) with its non-distributed CG; the arithmetic
    is the same, and `value` uses the same definition as the GPU arm (iterations/s x N slabs for
    weak scaling).  The matrix comes from the oracle's own generator: the product library is
    never loaded in this process.  Rank 0 alone runs; torchrun's OMP_NUM_THREADS=1 is overridden."""
    if rank != 0:
        return
    import oracle
    cores = host_threads()
    dist_cfg = args.gpus > 1 or args.dist
    if dist_cfg:
        kind, g = args.stencil or "7pt", args.grid or 512
        planes = args.slab_planes or max(g // 8, 1)
        nz = planes * args.gpus if args.scaling == "weak" else g
        nx = ny = g
        precond_block = 1 if args.dist_precond == "jacobi" else 0
        slabs = args.gpus if args.scaling == "weak" else 1
        workload = (f"CG ({'scalar Jacobi' if precond_block else 'no preconditioner'}), 3D {kind} stencil "
                    f"{nx}x{ny}x{nz} (the global system of BASELINE configs[3] at {args.gpus} GPUs, {args.scaling} "
                    "scaling), fp64/int32 CSR, one host")
    else:
        kind, g = "27pt", args.grid or 200
        nx = ny = nz = g
        precond_block, slabs = 1, 1
        workload = f"CG + scalar Jacobi, 3D 27-pt stencil {g}^3, fp64/int32 CSR (BASELINE configs[1])"
    rp, ci, va, n = oracle.gen_stencil_csr(kind, nx, ny, nz)
    b = np.ones(n)
    it_per = args.ref_iters_per_step
    if oracle.ref() is not None:
        oracle.ref().ref_set_num_threads(cores)
        kind_, cores = "reference", oracle.ref_threads()

        def step():
            _, it, _, secs = oracle.ref_solve(rp, ci, va, b, np.zeros(n), precond_block=precond_block, max_iters=it_per,
                                              factor=0.0, omp=True, want_hist=False)
            return it, secs
    else:
        kind_, cores = "port", 1
        inv = 1.0 / np.full(n, STENCIL_DIAG_REF[kind])

        def step():
            t0 = time.perf_counter()
            _, it, _, _ = oracle.cg_solve(rp, ci, va, b, np.zeros(n), precond=precond_block, inv_diag=inv,
                                          max_iters=it_per, factor=0.0)
            return it, time.perf_counter() - t0
    for _ in range(args.warmup):
        step()
    total_s, total_it = 0.0, 0
    for _ in range(args.steps):
        it, s = step()
        total_it += it
        total_s += s
    value = slabs * total_it / total_s
    sample = f"{workload}; {it_per} iterations per step on the host"
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * total_s / args.steps,
        "higher_is_better": True, "scaling": args.scaling if dist_cfg else "weak", "vs_baseline": None, "dtype": "f64",
        "data": "synthetic",
        "config": {"workload": workload, "rows": n, "nnz": int(len(ci)), "iters_per_step": it_per,
                   "value_definition": "CG iterations/s x N slabs" if slabs > 1 else "CG iterations/s"},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": kind_, "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


STENCIL_DIAG_REF = {"7pt": 6.0, "27pt": 26.0}


def ncu_traffic(fmt, grid):
    """DRAM read + write bytes per launch of the dominant kernel.  ncu cannot run inside a timed
    benchmark, so this is NOT measured in this run: it is read from the committed `ncu --set full`
    capture of the same kernel and matrix (profiles/r02_traffic.json, written by
    tools/ncu_summary.py from profiles/r02_csr_spmv_rowblock_tma_ncu_full.txt)."""
    try:
        with open(os.path.join(ROOT, "profiles", "r02_traffic.json")) as f:
            return json.load(f).get(f"{fmt}_27pt_{grid}")
    except (OSError, ValueError):
        return None


def main():
    args = parse()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference(args, rank, world)
        return

    import torch
    import torch.distributed as dist
    from __graft_entry__ import load_package
    gko = load_package()
    if world > 1 or args.dist:
        if world == 1:   # not under torchrun: a one-rank process group
            for k, v in (("MASTER_ADDR", "127.0.0.1"), ("MASTER_PORT", "29555"), ("RANK", "0"), ("WORLD_SIZE", "1")):
                os.environ.setdefault(k, v)
        from bench_dist import run_distributed
        run_distributed(args, gko, rank, world, local_rank)
        return

    exec_ = gko.CudaExecutor.create(local_rank)
    g = args.grid or 200
    rp, ci, va, n = gko.gen.stencil_csr("27pt", g, g, g)
    nnz = int(len(ci))
    A = gko.matrix.Csr.from_arrays(exec_, (n, n), rp, ci, va)
    Aop = A
    if args.format != "csr":
        Aop = A.convert_to(args.format)
    iters = args.iters_per_step
    # the preconditioner is generated from the CSR matrix (Jacobi::generate converts its
    # system matrix to CSR in the reference too, core/preconditioner/jacobi.cpp)
    jacobi = gko.preconditioner.Jacobi.build().with_max_block_size(1).on(exec_).generate(A)
    solver = (gko.solver.Cg.build()
              .with_criteria(gko.stop.Iteration(iters))
              .with_generated_preconditioner(jacobi)
              .with_check_every(max(iters, 1))
              .on(exec_).generate(Aop))
    b_host = torch.ones(n, dtype=torch.float64).pin_memory()
    x_host = torch.zeros(n, dtype=torch.float64).pin_memory()
    db = gko.matrix.Dense.create(exec_, (n, 1))
    dx = gko.matrix.Dense.create(exec_, (n, 1))
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
        torch.cuda.synchronize()
        e0.record()
        for _ in range(steps):
            launches += fn()
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) * 1e-3, launches

    sampler = ClockSampler(local_rank)
    sampler.start()
    for _ in range(max(args.warmup, 3)):
        step_device()
    sampler.mark()
    secs, launches = timed(step_device, args.steps)
    clocks = sampler.stop()
    assert solver.num_iterations == iters, solver.num_iterations
    value = args.steps * iters / secs

    for _ in range(2):
        step_host()
    secs_e2e, _ = timed(step_host, args.steps)
    e2e_value = args.steps * iters / secs_e2e

    # dominant kernel alone: SpMV q = A p, CUDA events on the launching stream; operands
    # (2.7 GB) exceed the 126 MB L2 so consecutive launches cannot hit in cache
    p = gko.matrix.Dense.create(exec_, (n, 1))
    q = gko.matrix.Dense.create(exec_, (n, 1))
    p.t.copy_(torch.randn(n, 1, dtype=torch.float64, device=exec_.device))
    for _ in range(5):
        Aop.apply(p, q)
    spmv_s, _ = timed(lambda: (Aop.apply(p, q), 1)[1], args.spmv_reps)
    spmv_s /= args.spmv_reps
    peak, peak_src = peaks()
    spmv_bytes = Aop.spmv_bytes(1)
    achieved = spmv_bytes / spmv_s / 1e9
    precond_bytes = solver.get_preconditioner().storage_bytes()
    it_bytes = cg_model_bytes(n, nnz, precond_bytes)

    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": 1, "steps": args.steps, "warmup": max(args.warmup, 3),
        "ms_per_step": 1e3 * secs / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f64", "data": "synthetic",
        "config": {"workload": f"CG + scalar Jacobi, 3D 27-pt stencil {g}^3, fp64/int32 {args.format.upper()} "
                               "(BASELINE configs[1])",
                   "rows": n, "nnz": nnz, "iters_per_step": iters, "spmv_kernel": getattr(Aop, "kernel", lambda: args.format)(),
                   "l2": "operands (2.7 GB matrix + vectors) exceed the 126 MB L2; no flush needed",
                   "value_definition": "CG iterations/s (x N slabs at N GPUs)"},
        "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": 2 * n * 8, "d2h_bytes_per_step": n * 8 + 16},
        "gpu_launches": launches,
        "clocks": clocks,
        "roofline": {"bound": "hbm", "kernel": f"{args.format}_spmv", "achieved": achieved, "peak": peak, "unit": "GB/s",
                     "frac": achieved / peak, "peak_source": peak_src, "frac_of_nominal_8000": achieved / 8000.0,
                     "bytes_per_launch": spmv_bytes, "us_per_launch": spmv_s * 1e6,
                     "traffic": ncu_traffic(args.format, g),
                     "traffic_source": "committed ncu --set full capture of this kernel on this matrix "
                                       "(profiles/r02_csr_spmv_rowblock_tma_ncu_full.txt), not measured in this run"},
        "cg_iteration": {"us": 1e6 * secs / (args.steps * iters), "model_bytes": it_bytes,
                         "model_gbs": it_bytes * args.steps * iters / secs / 1e9,
                         "frac_of_peak": it_bytes * args.steps * iters / secs / 1e9 / peak},
    }
    if not args.no_cpu_baseline:
        line["cpu_baseline"] = cpu_baseline(args, rp, ci, va, n)
    if not args.no_config_block:
        del A, Aop, solver, jacobi, db, dx, p, q, rp, ci, va
        sys.path.insert(0, os.path.join(ROOT, "tools"))
        from config_block import run_block
        line["configs"] = run_block(gko, exec_, peak, small=args.small_config_block)
    print(json.dumps(line), flush=True)


def cpu_baseline(args, rp, ci, va, n):
    import oracle
    g, its = args.grid or 200, args.cpu_baseline_iters
    b = np.ones(n)
    sample = f"same matrix (27-pt {g}^3), CG + scalar Jacobi, {its} iterations, 1 warm-up solve of 1 iteration"
    if oracle.ref() is not None:
        oracle.ref().ref_set_num_threads(host_threads())
        oracle.ref_solve(rp, ci, va, b, np.zeros(n), precond_block=1, max_iters=1, factor=0.0, omp=True, want_hist=False)
        _, it, _, secs = oracle.ref_solve(rp, ci, va, b, np.zeros(n), precond_block=1, max_iters=its, factor=0.0,
                                          omp=True, want_hist=False)
        _, spmv_t = oracle.ref_spmv(rp, ci, va, b, omp=True, reps=3)
        return {"value": it / secs, "unit": UNIT, "cores": oracle.ref_threads(), "kind": "reference",
                "sample": sample + "; Ginkgo 1.5.0 OmpExecutor built from /root/reference (oracle/_ref)",
                "spmv_gbs": (len(ci) * 12 + (n + 1) * 4 + 2 * n * 8) / spmv_t[1] / 1e9}
    inv = 1.0 / np.full(n, 26.0)
    t0 = time.perf_counter()
    _, it, _, _ = oracle.cg_solve(rp, ci, va, b, np.zeros(n), precond=1, inv_diag=inv, max_iters=its, factor=0.0)
    secs = time.perf_counter() - t0
    return {"value": it / secs, "unit": UNIT, "cores": 1, "kind": "port", "sample": sample + "; plain-C oracle"}


if __name__ == "__main__":
    main()
