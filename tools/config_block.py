"""Compact per-config block for the N=1 line of bench.py: every BASELINE config's dominant SpMV
against the HBM roofline (algorithmic bytes of SURVEY.md §8d / launch time, CUDA events) and its
solver's iterations/s, measured in the same process as the headline number.  configs[1] is the
headline itself and is not repeated here; configs[3] appears as its single-GPU slab (the
weak-scaling denominator of the N>1 runs)."""
import time

import numpy as np
import torch


def _ev_time(fn, reps, warm=3):
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


def _spmv(gko, exec_, peak, A, dtype, nrhs=1, reps=20):
    tdt = torch.float64 if dtype == np.float64 else torch.float32
    n, m = A.size
    x = gko.matrix.Dense.create(exec_, (m, nrhs), tdt)
    y = gko.matrix.Dense.create(exec_, (n, nrhs), tdt)
    x.t.copy_(torch.randn(m, nrhs, dtype=tdt, device=exec_.device))
    t = _ev_time(lambda: A.apply(x, y), reps)
    gbs = A.spmv_bytes(nrhs) / t / 1e9
    return {"us": round(t * 1e6, 1), "gbs": round(gbs, 1), "frac": round(gbs / peak, 4), "bytes": A.spmv_bytes(nrhs)}


def _solve(gko, exec_, kind, A, precond, dtype, iters, krylov_dim=30):
    tdt = torch.float64 if dtype == np.float64 else torch.float32
    n = A.size[0]
    f = getattr(gko.solver, kind).build().with_criteria(gko.stop.Iteration(iters)).with_krylov_dim(krylov_dim)
    if precond is not None:
        f = f.with_generated_preconditioner(precond)
    s = f.with_check_every(iters).on(exec_).generate(A)
    b = gko.matrix.Dense.create(exec_, (n, 1), tdt)
    b.fill(1.0)
    x = gko.matrix.Dense.create(exec_, (n, 1), tdt)

    def run():
        x.fill(0.0)
        s.apply(b, x)
    t = _ev_time(run, 2, warm=1)
    assert s.num_iterations == iters, (s.num_iterations, iters)
    hist = s.residual_history
    assert len(hist) > 2 and hist[1] != hist[0], "the solve does not move"
    return {"iters_per_s": round(iters / t, 1), "us_per_iter": round(t / iters * 1e6, 1)}


def run_block(gko, exec_, peak, small=False):
    out = {"unit_spmv": "GB/s of algorithmic bytes; frac = of the measured HBM copy peak"}
    t0 = time.time()
    # C1
    g = 1000 if not small else 200
    rp, ci, va, n = gko.gen.stencil_csr("5pt", g, g)
    A = gko.matrix.Csr.from_arrays(exec_, (n, n), rp, ci, va)
    out["C1"] = {"workload": f"CG (no preconditioner), 2D 5-pt {g}x{g}, fp64 CSR", "spmv_kernel": A.kernel(),
                 "spmv": _spmv(gko, exec_, peak, A, np.float64, reps=200),
                 "cg": _solve(gko, exec_, "Cg", A, None, np.float64, 1000 if not small else 100)}
    del A
    # C3
    n = 10_000_000 if not small else 300_000
    rp, ci, va = gko.gen.powerlaw_csr(n)
    A = gko.matrix.Csr.from_arrays(exec_, (n, n), rp, ci, va)
    J = gko.preconditioner.Jacobi.build().with_max_block_size(32).on(exec_).generate(A)
    bb, xx = gko.matrix.Dense.create(exec_, (n, 1)), gko.matrix.Dense.create(exec_, (n, 1))
    bb.fill(1.0)
    tj = _ev_time(lambda: J.apply(bb, xx), 10)
    jbytes = J.storage_bytes() + 2 * n * 8
    out["C3"] = {"workload": f"GMRES(30) + block-Jacobi(32), power-law matrix {n} rows / {len(ci)} nnz, fp64 CSR",
                 "spmv_kernel": A.kernel(), "max_row_nnz": int(A.max_row_nnz),
                 "spmv": _spmv(gko, exec_, peak, A, np.float64),
                 "block_jacobi_apply": {"us": round(tj * 1e6, 1), "gbs": round(jbytes / tj / 1e9, 1),
                                        "frac": round(jbytes / tj / 1e9 / peak, 4)},
                 "gmres": _solve(gko, exec_, "Gmres", A, J, np.float64, 60)}
    del A, J, bb, xx
    # C4: the single-GPU slab of the multi-GPU runs
    g = 512 if not small else 64
    rp, ci, va, n = gko.gen.stencil_csr("7pt", g, g, max(g // 8, 1))
    A = gko.matrix.Csr.from_arrays(exec_, (n, n), rp, ci, va)
    out["C4_slab"] = {"workload": f"CG (no preconditioner), 3D 7-pt {g}x{g}x{max(g // 8, 1)} (one slab of configs[3]), fp64 CSR",
                      "spmv_kernel": A.kernel(), "spmv": _spmv(gko, exec_, peak, A, np.float64),
                      "cg": _solve(gko, exec_, "Cg", A, None, np.float64, 100)}
    del A
    # C5
    g = 256 if not small else 48
    rp, ci, va, n = gko.gen.stencil_csr("27pt", g, g, g, value_dtype=np.float32)
    A = gko.matrix.Csr.from_arrays(exec_, (n, n), rp, ci, va)
    H16 = A.convert_to("hybrid", strategy=gko.matrix.HybridStrategy.column_limit(16))
    Hauto = A.convert_to("hybrid")
    out["C5"] = {"workload": f"BiCGSTAB fp32 on Hybrid ELL+COO + 32-RHS SpMM, 3D 27-pt {g}^3",
                 "hybrid_column_limit16": {"ell_width": int(H16.ell.width), "coo_nnz": int(H16.coo.values.numel()),
                                           "spmv": _spmv(gko, exec_, peak, H16, np.float32),
                                           "bicgstab": _solve(gko, exec_, "Bicgstab", H16, None, np.float32, 50)},
                 "hybrid_automatic": {"ell_width": int(Hauto.ell.width), "coo_nnz": int(Hauto.coo.values.numel()),
                                      "spmv": _spmv(gko, exec_, peak, Hauto, np.float32),
                                      "bicgstab": _solve(gko, exec_, "Bicgstab", Hauto, None, np.float32, 50)},
                 "csr_spmm_32rhs": _spmv(gko, exec_, peak, A, np.float32, nrhs=32, reps=5),
                 "ell_spmm_32rhs": _spmv(gko, exec_, peak, Hauto.ell, np.float32, nrhs=32, reps=5)}
    out["seconds"] = round(time.time() - t0, 1)
    return out
