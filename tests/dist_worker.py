"""torchrun worker for tests/test_gpu_dist.py::test_multi_gpu_torchrun: one rank per GPU.
Checks, against the single-process oracle on the SAME global matrix:
  * halo plan (send/recv sizes, gather idxs) bit-exact with oracle.dist_plan,
  * distributed apply and advanced apply == global CSR apply (1e-12 relative to sum|a||b|),
  * distributed CG: iteration count within 2 of the global oracle CG, solution to 1e-9,
  * distributed::Vector dot / norm2 across ranks."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

import oracle  # noqa: E402
from __graft_entry__ import load_package  # noqa: E402


def main():
    rank, world, lr = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(lr)
    dist.init_process_group("nccl", device_id=torch.device("cuda", lr))
    gko = load_package()
    D = gko.distributed
    exec_ = gko.CudaExecutor.create(lr)
    comm = D.Communicator.from_torch(exec_)
    nx, ny, nz = 20, 18, 6 * world
    rp, ci, va, n = gko.gen.stencil_csr("27pt", nx, ny, nz, index_dtype=np.int64)
    rows = np.repeat(np.arange(n, dtype=np.int64), np.diff(rp))
    part = D.Partition.build_from_global_size_uniform(exec_, world, n)
    opart = oracle.Partition.uniform(world, n)
    lo, hi = int(opart.bounds[rank]), int(opart.bounds[rank + 1])
    mine = slice(int(rp[lo]), int(rp[hi]))          # each rank only hands in its own rows
    A = D.Matrix(exec_, comm).read_distributed(rows[mine], ci[mine], va[mine], part)
    # halo plan vs oracle
    allp = [oracle.dist_build_local_nonlocal(rows, ci, va, opart, p) for p in range(world)]
    send, recv, gathers = oracle.dist_plan(allp)
    assert np.array_equal(A.send_sizes, send[rank]) and np.array_equal(A.recv_sizes, recv[rank])
    assert np.array_equal(A.gather_idxs.cpu().numpy(), gathers[rank])
    # apply vs global
    rng = np.random.default_rng(0)
    xg, yg = rng.standard_normal((n, 1)), rng.standard_normal((n, 1))
    rp32, ci32 = rp.astype(np.int32), ci.astype(np.int32)
    want = oracle.csr_spmv(rp32, ci32, va, xg)
    want_adv = oracle.csr_spmv(rp32, ci32, va, xg, 0.5, -2.0, yg)
    bound = 27.0 * 26.0 * np.abs(xg).max() + 2 * np.abs(yg).max()
    dx = gko.matrix.Dense.from_numpy(exec_, xg[lo:hi])
    dy = gko.matrix.Dense.create(exec_, (hi - lo, 1))
    A.apply(dx, dy)
    assert np.abs(dy.to_numpy() - want[lo:hi]).max() <= 1e-12 * bound
    dy2 = gko.matrix.Dense.from_numpy(exec_, yg[lo:hi])
    A.apply(gko.matrix.Dense.scalar(exec_, 0.5), dx, gko.matrix.Dense.scalar(exec_, -2.0), dy2)
    assert np.abs(dy2.to_numpy() - want_adv[lo:hi]).max() <= 1e-12 * bound
    # bit-exact against the CPU restatement of the distributed apply
    assert np.array_equal(dy.to_numpy(), oracle.dist_apply(allp, opart, xg)[lo:hi])
    # Vector reductions
    v, w = D.Vector(comm, dx), D.Vector(comm, gko.matrix.Dense.from_numpy(exec_, yg[lo:hi]))
    res = gko.matrix.Dense.create(exec_, (1, 1))
    v.compute_dot(w, res)
    assert np.isclose(res.to_numpy()[0, 0], float((xg * yg).sum()), rtol=1e-12)
    v.compute_norm2(res)
    assert np.isclose(res.to_numpy()[0, 0], np.linalg.norm(xg), rtol=1e-13)
    # distributed CG (+ scalar Jacobi) vs the global oracle CG
    b = rng.standard_normal(n)
    for jac in (False, True):
        x_ref, it_ref, hist_ref, _ = oracle.cg_solve(rp32, ci32, va, b, np.zeros(n), precond=int(jac),
                                                     inv_diag=1 / np.full(n, 26.0), max_iters=500, factor=1e-9)
        M = gko.preconditioner.Jacobi.build().with_max_block_size(1).on(exec_).generate(A.local) if jac else None
        s = D.cg(exec_, A, [gko.stop.Iteration(500), gko.stop.ResidualNorm(1e-9)], precond=M, check_every=5)
        dxs = gko.matrix.Dense.create(exec_, (hi - lo, 1))
        s.apply(gko.matrix.Dense.from_numpy(exec_, b[lo:hi]), dxs)
        assert abs(s.num_iterations - it_ref) <= 2, (s.num_iterations, it_ref)
        assert np.allclose(s.residual_history[:10], hist_ref[:10], rtol=1e-10)
        assert np.abs(dxs.to_numpy()[:, 0] - x_ref[lo:hi]).max() <= 1e-9 * np.abs(x_ref).max()
    dist.barrier()
    if rank == 0:
        print(f"DIST_OK world={world}", flush=True)
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
