"""Generates tests/golden/ref_golden.npz from the UNMODIFIED reference (Ginkgo's ReferenceExecutor compiled
into oracle/_ref by oracle/Makefile.ref).  Run in the container that has /root/reference:

    python tests/golden/make_golden.py

The file holds inputs AND the reference's outputs, so tests/test_golden.py needs neither the reference nor a
particular numpy / scipy random stream: it pins the plain-C oracle (and through it the CUDA path, which the
-m gpu tests compare with the oracle bit for bit) to what the reference itself computed.
Cases follow the reference's own tests: csr apply sizes of test/matrix/csr_kernels2.cpp:63-90 (scaled down),
solver systems like reference/test/solver/{cg,bicgstab,gmres}_kernels.cpp, conversions of
reference/test/matrix/csr_kernels.cpp (convert_to ell / sellp / hybrid)."""
import os
import sys

import numpy as np
import scipy.sparse as sp

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import oracle as ora  # noqa: E402


def random_csr(n, m, density, seed, dtype=np.float64, skew=False):
    rng = np.random.default_rng(seed)
    a = sp.random(n, m, density=density, random_state=rng, format="lil", dtype=np.float64)
    if skew:
        a[3, :] = 1.0          # one full row, a few empty ones
        a[7, :] = 0.0
        a[8, :] = 0.0
    a = a.tocsr()
    a.data = rng.uniform(-1, 1, a.nnz)
    a.sort_indices()
    return a.indptr.astype(np.int32), a.indices.astype(np.int32), a.data.astype(dtype)


def main():
    out = {}
    # ---- SpMV / advanced SpMV / SpMM -----------------------------------------------------------
    for tag, dtype in (("f64", np.float64), ("f32", np.float32)):
        rp, ci, va = random_csr(61, 43, 0.12, 42, dtype, skew=True)
        rng = np.random.default_rng(7)
        b = rng.uniform(-1, 1, (43, 3)).astype(dtype)
        c0 = rng.uniform(-1, 1, (61, 3)).astype(dtype)
        out[f"spmv_{tag}_rp"], out[f"spmv_{tag}_ci"], out[f"spmv_{tag}_va"] = rp, ci, va
        out[f"spmv_{tag}_b"], out[f"spmv_{tag}_c0"] = b, c0
        out[f"spmv_{tag}_plain"], _ = ora.ref_spmv(rp, ci, va, b)
        out[f"spmv_{tag}_adv"], _ = ora.ref_spmv(rp, ci, va, b, alpha=0.7, beta=-1.3, c=c0)
        for fmt in ("ell", "sellp", "coo", "hybrid"):
            out[f"spmv_{tag}_{fmt}"], _ = ora.ref_spmv(rp, ci, va, b, fmt=fmt)
    # ---- conversions (integer outputs must be bit-exact) ---------------------------------------
    rp, ci, va = out["spmv_f64_rp"], out["spmv_f64_ci"], out["spmv_f64_va"]
    for fmt, kw in (("ell", {}), ("sellp", {"slice_size": 8, "stride_factor": 2}),
                    ("hybrid", {"hyb_kind": "column_limit", "hyb_param": 4}),
                    ("hybrid", {"hyb_kind": "automatic"})):
        r = ora.ref_convert(rp, ci, va, 43, fmt, **kw)
        name = fmt + ("_" + kw["hyb_kind"] if "hyb_kind" in kw else "")
        for k, v in r.items():
            out[f"conv_{name}_{k}"] = np.asarray(v)
    # ---- solvers: 5-pt Laplacian 12 x 12, b = sin, tolerance 1e-10 --------------------------------
    rp, ci, va = ora.gen_stencil_csr("5pt", 12, 12)[:3]
    n = len(rp) - 1
    b = np.sin(0.1 * np.arange(n))
    out["solve_rp"], out["solve_ci"], out["solve_va"], out["solve_b"] = rp, ci, va, b
    for solver in ("cg", "bicgstab", "gmres", "fcg", "cgs"):
        for pb in (0, 1):
            x, it, hist, _ = ora.ref_solve(rp, ci, va, b, np.zeros(n), solver=solver, precond_block=pb, factor=1e-10,
                                           max_iters=400)
            out[f"solve_{solver}_p{pb}_x"], out[f"solve_{solver}_p{pb}_hist"] = x, np.asarray(hist)
            out[f"solve_{solver}_p{pb}_it"] = np.int64(it)
    np.savez_compressed(os.path.join(ROOT, "tests", "golden", "ref_golden.npz"), **out)
    print("wrote", len(out), "arrays")


if __name__ == "__main__":
    main()
