#!/usr/bin/env python
"""All five BASELINE.json configs on one B200 (config 4 single-GPU slab here; its
multi-GPU numbers come from bench.py --gpus N), with the reference's OpenMP executor timed
on the host beside each of them (bounded samples).  Prints one JSON object per config and a
markdown table; results are committed under profiles/.

  python tools/bench_configs.py [--quick] [--no-cpu] [--only c1,c3]
"""
import argparse
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np  # noqa: E402
import torch  # noqa: E402

from __graft_entry__ import load_package  # noqa: E402

PEAK = 6547.2
ap = argparse.ArgumentParser()
ap.add_argument("--quick", action="store_true")
ap.add_argument("--no-cpu", action="store_true")
ap.add_argument("--only", default="")
args = ap.parse_args()
gko = load_package()
exec_ = gko.CudaExecutor.create(0)
try:
    p = json.load(open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json")))
    PEAK = float(p["hbm_gbs"])
except Exception:
    pass
import oracle  # noqa: E402  (CPU baseline leg only)
HAVE_REF = oracle.ref() is not None and not args.no_cpu


def ev_time(fn, reps, warm=3):
    for _ in range(warm):
        fn()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) * 1e-3 / reps


def spmv_row(name, A, dtype, nrhs=1, reps=20):
    tdt = torch.float64 if dtype == np.float64 else torch.float32
    n, m = A.size
    x = gko.matrix.Dense.create(exec_, (m, nrhs), tdt)
    y = gko.matrix.Dense.create(exec_, (n, nrhs), tdt)
    x.t.copy_(torch.randn(m, nrhs, dtype=tdt, device=exec_.device))
    t = ev_time(lambda: A.apply(x, y), reps)
    gbs = A.spmv_bytes(nrhs) / t / 1e9
    return {"op": name, "us": t * 1e6, "gbs": gbs, "frac_of_measured_peak": gbs / PEAK, "frac_of_8TBs": gbs / 8000}


def solve_row(name, kind, A, precond, dtype, iters, krylov_dim=30, nrhs=1):
    tdt = torch.float64 if dtype == np.float64 else torch.float32
    n = A.size[0]
    f = getattr(gko.solver, kind).build().with_criteria(gko.stop.Iteration(iters)).with_krylov_dim(krylov_dim)
    if precond is not None:
        f = f.with_generated_preconditioner(precond)
    s = f.with_check_every(iters).on(exec_).generate(A, nrhs=nrhs)
    b = gko.matrix.Dense.create(exec_, (n, nrhs), tdt)
    b.fill(1.0)
    x = gko.matrix.Dense.create(exec_, (n, nrhs), tdt)

    def run():
        x.fill(0.0)
        s.apply(b, x)
    t = ev_time(run, 2, warm=1)
    assert s.num_iterations == iters, (s.num_iterations, iters)
    return {"op": name, "iters_per_s": iters / t, "us_per_iter": t / iters * 1e6, "launches_per_iter": s.launch_count / iters}


def cpu_spmv(rp, ci, va, fmt="csr", hybrid_limit=-1, nrhs=1):
    if not HAVE_REF:
        return None
    b = np.ones((len(rp) - 1, nrhs), dtype=va.dtype)
    _, (mean_s, best_s) = oracle.ref_spmv(rp, ci, va, b, fmt=fmt, hybrid_limit=hybrid_limit, omp=True, reps=3)
    return best_s


def cpu_solve(rp, ci, va, solver, precond_block, iters, krylov_dim=30, fmt="csr", hybrid_limit=-1):
    if not HAVE_REF:
        return None
    n = len(rp) - 1
    b = np.ones(n, dtype=va.dtype)
    _, it, _, secs = oracle.ref_solve(rp, ci, va, b, np.zeros(n, dtype=va.dtype), solver=solver, fmt=fmt,
                                      hybrid_limit=hybrid_limit, precond_block=precond_block, max_iters=iters,
                                      factor=0.0, krylov_dim=krylov_dim, omp=True, want_hist=False)
    return it / secs


out = []
only = set(args.only.split(",")) if args.only else None
q = args.quick


def want(c):
    return only is None or c in only


cores = oracle.ref_threads() if HAVE_REF else 0

if want("c1"):
    g = 1000
    rp, ci, va, n = gko.gen.stencil_csr("5pt", g, g)
    A = gko.matrix.Csr.from_arrays(exec_, (n, n), rp, ci, va)
    r = {"config": "C1: CG (no precond) on 2D 5-pt 1000x1000, fp64 CSR", "rows": n, "nnz": len(ci),
         "gpu": [spmv_row("CSR SpMV", A, np.float64, reps=200), solve_row("CG", "Cg", A, None, np.float64, 1000)]}
    t = cpu_spmv(rp, ci, va)
    r["cpu_omp"] = {"cores": cores, "spmv_gbs": A.spmv_bytes() / t / 1e9 if t else None,
                    "cg_iters_per_s": cpu_solve(rp, ci, va, "cg", 0, 50)}
    out.append(r)

if want("c2"):
    g = 200 if not q else 100
    rp, ci, va, n = gko.gen.stencil_csr("27pt", g, g, g)
    A = gko.matrix.Csr.from_arrays(exec_, (n, n), rp, ci, va)
    S = A.convert_to("sellp")
    J = gko.preconditioner.Jacobi.build().with_max_block_size(1).on(exec_).generate(A)
    r = {"config": f"C2: CG + scalar Jacobi on 3D 27-pt {g}^3, fp64, CSR vs SELL-P", "rows": n, "nnz": len(ci),
         "gpu": [spmv_row("CSR SpMV", A, np.float64), spmv_row("SELL-P SpMV", S, np.float64),
                 solve_row("CG+Jacobi (CSR)", "Cg", A, J, np.float64, 100),
                 solve_row("CG+Jacobi (SELL-P)", "Cg", S, J, np.float64, 100)]}
    t = cpu_spmv(rp, ci, va)
    ts = cpu_spmv(rp, ci, va, fmt="sellp")
    r["cpu_omp"] = {"cores": cores, "csr_spmv_gbs": A.spmv_bytes() / t / 1e9 if t else None,
                    "sellp_spmv_gbs": S.spmv_bytes() / ts / 1e9 if ts else None,
                    "cg_iters_per_s": cpu_solve(rp, ci, va, "cg", 1, 6)}
    out.append(r)
    del A, S, J

if want("c3"):
    n = 10_000_000 if not q else 1_000_000
    rp, ci, va = gko.gen.powerlaw_csr(n)
    A = gko.matrix.Csr.from_arrays(exec_, (n, n), rp, ci, va)
    t0 = time.time()
    J = gko.preconditioner.Jacobi.build().with_max_block_size(32).on(exec_).generate(A)
    exec_.synchronize()
    gen_s = time.time() - t0
    tdt = torch.float64
    bb, xx = gko.matrix.Dense.create(exec_, (n, 1), tdt), gko.matrix.Dense.create(exec_, (n, 1), tdt)
    bb.fill(1.0)
    tj = ev_time(lambda: J.apply(bb, xx), 10)
    jbytes = J.storage_bytes() + 2 * n * 8
    r = {"config": f"C3: GMRES(30) + block-Jacobi(32) on power-law matrix, {n} rows, fp64, merge-path CSR",
         "rows": n, "nnz": len(ci), "max_row_nnz": A.max_row_nnz, "spmv_kernel": A.kernel(),
         "jacobi_blocks": J.num_blocks, "jacobi_generate_s": gen_s,
         "gpu": [spmv_row("CSR merge-path SpMV", A, np.float64),
                 {"op": "block-Jacobi(32) apply", "us": tj * 1e6, "gbs": jbytes / tj / 1e9,
                  "frac_of_measured_peak": jbytes / tj / 1e9 / PEAK},
                 solve_row("GMRES(30)+block-Jacobi", "Gmres", A, J, np.float64, 60)]}
    t = cpu_spmv(rp, ci, va)
    r["cpu_omp"] = {"cores": cores, "spmv_gbs": A.spmv_bytes() / t / 1e9 if t else None,
                    "gmres_iters_per_s": cpu_solve(rp, ci, va, "gmres", 32, 4) if not q else None}
    out.append(r)
    del A, J

if want("c4"):
    g = 512 if not q else 128
    nz = g // 8   # the per-GPU slab of the 8-GPU run, on one GPU (multi-GPU numbers: bench.py --gpus N)
    rp, ci, va, n = gko.gen.stencil_csr("7pt", g, g, nz)
    A = gko.matrix.Csr.from_arrays(exec_, (n, n), rp, ci, va)
    r = {"config": f"C4 (one slab of 8): CG on 3D 7-pt {g}x{g}x{nz}, fp64 CSR", "rows": n, "nnz": len(ci),
         "gpu": [spmv_row("CSR SpMV", A, np.float64), solve_row("CG", "Cg", A, None, np.float64, 100)]}
    t = cpu_spmv(rp, ci, va)
    r["cpu_omp"] = {"cores": cores, "spmv_gbs": A.spmv_bytes() / t / 1e9 if t else None,
                    "cg_iters_per_s": cpu_solve(rp, ci, va, "cg", 0, 6)}
    out.append(r)
    del A

if want("c5"):
    g = 256 if not q else 96
    rp, ci, va, n = gko.gen.stencil_csr("27pt", g, g, g, value_dtype=np.float32)
    A = gko.matrix.Csr.from_arrays(exec_, (n, n), rp, ci, va)
    H16 = A.convert_to("hybrid", strategy=gko.matrix.HybridStrategy.column_limit(16))
    Hauto = A.convert_to("hybrid")
    E = Hauto.ell
    r = {"config": f"C5: BiCGSTAB fp32 on Hybrid ELL+COO + 32-RHS SpMM, 3D 27-pt {g}^3", "rows": n, "nnz": len(ci),
         "hybrid_column_limit16": {"ell_width": H16.ell.width, "coo_nnz": int(H16.coo.values.numel())},
         "hybrid_automatic": {"ell_width": Hauto.ell.width, "coo_nnz": int(Hauto.coo.values.numel())},
         "gpu": [spmv_row("Hybrid(column_limit 16) SpMV", H16, np.float32),
                 spmv_row("Hybrid(automatic = pure ELL) SpMV", Hauto, np.float32),
                 spmv_row("ELL SpMM 32 RHS", E, np.float32, nrhs=32, reps=5),
                 spmv_row("CSR SpMM 32 RHS", A, np.float32, nrhs=32, reps=5),
                 solve_row("BiCGSTAB (Hybrid column_limit 16)", "Bicgstab", H16, None, np.float32, 50),
                 solve_row("BiCGSTAB (Hybrid automatic)", "Bicgstab", Hauto, None, np.float32, 50)]}
    t = cpu_spmv(rp, ci, va, fmt="hybrid", hybrid_limit=16)
    r["cpu_omp"] = {"cores": cores, "hybrid16_spmv_gbs": H16.spmv_bytes() / t / 1e9 if t else None,
                    "bicgstab_iters_per_s": cpu_solve(rp, ci, va, "bicgstab", 0, 4, fmt="hybrid", hybrid_limit=16)}
    out.append(r)

for r in out:
    print(json.dumps(r))
print("\n| config | op | GPU | of measured HBM peak | OMP reference (%d cores) |" % cores)
print("|---|---|---|---|---|")
for r in out:
    cpu = r.get("cpu_omp", {})
    for gline in r["gpu"]:
        if "gbs" in gline:
            val = f"{gline['us']:.0f} us, {gline['gbs']:.0f} GB/s"
            frac = f"{100 * gline['frac_of_measured_peak']:.0f} %"
        else:
            val = f"{gline['iters_per_s']:.0f} iters/s ({gline['us_per_iter']:.0f} us/iter)"
            frac = ""
        cpus = ", ".join(f"{k}={v:.1f}" for k, v in cpu.items() if isinstance(v, float))
        print(f"| {r['config'][:60]} | {gline['op']} | {val} | {frac} | {cpus} |")
