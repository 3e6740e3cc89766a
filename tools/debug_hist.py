"""Debug: residual histories of repeated applies, Iteration-only vs ResidualNorm, single-GPU and
one-rank distributed CG, against the oracle."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import oracle
from __graft_entry__ import load_package
gko = load_package()
D = gko.distributed
exec_ = gko.CudaExecutor.create(0)
for dims in ((32, 32, 16), (128, 128, 32)):
    nx, ny, nz = dims
    rp, ci, va, n = gko.gen.stencil_csr("7pt", nx, ny, nz)
    b = np.ones(n)
    _, it, hist, _ = oracle.cg_solve(rp, ci, va, b, np.zeros(n), max_iters=12, factor=0.0)
    print(dims, "oracle", hist[:6])
    A1 = gko.matrix.Csr.from_arrays(exec_, (n, n), rp, ci, va)
    rp64, ci64, va64, _ = gko.gen.stencil_csr("7pt", nx, ny, nz, index_dtype=np.int64)
    rows = np.repeat(np.arange(n, dtype=np.int64), np.diff(rp64))
    comm = D.Communicator.single(exec_)
    part = D.Partition.build_from_global_size_uniform(exec_, 1, n)
    Ad = D.Matrix(exec_, comm).read_distributed(rows, ci64, va64, part)
    for name, mk in (("single", lambda crit, ce: gko.solver.Cg.build().with_criteria(*crit).with_check_every(ce).on(exec_).generate(A1)),
                     ("dist1", lambda crit, ce: D.cg(exec_, Ad, crit, check_every=ce))):
        for crit, ce in (([gko.stop.Iteration(12)], 12), ([gko.stop.Iteration(12)], 4), ([gko.stop.Iteration(300), gko.stop.ResidualNorm(1e-8)], 8)):
            s = mk(crit, ce)
            db = gko.matrix.Dense.from_numpy(exec_, b)
            for rep in range(3):
                dx = gko.matrix.Dense.create(exec_, (n, 1))
                s.apply(db, dx)
                print(dims, name, len(crit), ce, rep, s.num_iterations, np.asarray(s.residual_history[:5]))
