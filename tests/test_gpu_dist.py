"""Distributed path on the GPU.  On one GPU: the index kernels for every part in a loop
(the reference tests multi-rank logic the same way, reference/test/distributed/matrix_kernels.cpp:137)
and the 1-rank distributed matrix / CG.  With >= 2 GPUs: a torchrun job (tests/dist_worker.py)
checks the NCCL halo exchange and distributed CG against the global oracle."""
import os
import subprocess
import sys

import numpy as np
import pytest
import torch

from test_oracle_dist import random_global

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def npy(t):
    return t.detach().cpu().numpy()


def test_partition_kernels_bit_exact(gko, exec_, ora):
    D = gko.distributed
    for parts, size in ((5, 13), (8, 134_217_728), (3, 2), (1, 0)):
        p = D.Partition.build_from_global_size_uniform(exec_, parts, size)
        o = ora.Partition.uniform(parts, size)
        assert np.array_equal(npy(p.range_bounds), o.bounds) and np.array_equal(npy(p.part_ids), o.part_ids)
        assert np.array_equal(npy(p.range_starting_indices), o.starts) and np.array_equal(npy(p.part_sizes), o.sizes)
        assert p.num_empty_parts == o.num_empty_parts
    mapping = np.array([2, 2, 0, 1, 1, 2, 0, 0, 1, 0, 1, 1, 1, 2, 2, 0], np.int32)
    p = D.Partition.build_from_mapping(exec_, torch.from_numpy(mapping), 3)
    o = ora.Partition.from_mapping(mapping, 3)
    assert np.array_equal(npy(p.range_bounds), o.bounds) and np.array_equal(npy(p.part_ids), o.part_ids)
    assert np.array_equal(npy(p.range_starting_indices), o.starts) and np.array_equal(npy(p.part_sizes), o.sizes)


@pytest.mark.parametrize("num_parts,use_mapping,n,density", [(1, False, 57, 0.08), (3, False, 57, 0.08), (4, True, 300, 0.03),
                                                             (8, False, 5000, 0.002)])
def test_build_local_nonlocal_bit_exact(gko, exec_, ora, num_parts, use_mapping, n, density):
    D = gko.distributed
    A, rows, cols, vals = random_global(n, density, 5)
    if use_mapping:
        mapping = np.random.default_rng(2).integers(0, num_parts, n).astype(np.int32)
        part, opart = D.Partition.build_from_mapping(exec_, torch.from_numpy(mapping), num_parts), ora.Partition.from_mapping(mapping, num_parts)
    else:
        part, opart = D.Partition.build_from_global_size_uniform(exec_, num_parts, n), ora.Partition.uniform(num_parts, n)
    dr, dc, dv = (torch.from_numpy(a).to(exec_.device) for a in (rows, cols, vals))
    for lp in range(num_parts):
        got = D.build_local_nonlocal(exec_, dr, dc, dv, part, part, lp)
        want = ora.dist_build_local_nonlocal(rows, cols, vals, opart, lp)
        for k in ("lrow", "lcol", "lval", "nrow", "ncol", "nval", "gather", "recv_sizes", "nl_to_global"):
            assert np.array_equal(npy(got[k]), want[k]), (lp, k)


def test_single_rank_distributed_matrix_and_cg(gko, exec_, ora):
    D = gko.distributed
    rp, ci, va, n = gko.gen.stencil_csr("7pt", 12, 13, 14, index_dtype=np.int64)
    rows = np.repeat(np.arange(n, dtype=np.int64), np.diff(rp))
    comm = D.Communicator.single(exec_)
    part = D.Partition.build_from_global_size_uniform(exec_, 1, n)
    A = D.Matrix(exec_, comm).read_distributed(rows, ci, va, part)
    x = np.random.default_rng(0).standard_normal((n, 1))
    dx, dy = gko.matrix.Dense.from_numpy(exec_, x), gko.matrix.Dense.create(exec_, (n, 1))
    A.apply(dx, dy)
    rp32, ci32 = rp.astype(np.int32), ci.astype(np.int32)
    assert np.array_equal(dy.to_numpy(), ora.csr_spmv(rp32, ci32, va, x))
    b = np.random.default_rng(1).standard_normal(n)
    x_ref, it_ref, hist_ref, _ = ora.cg_solve(rp32, ci32, va, b, np.zeros(n), max_iters=300, factor=1e-9)
    s = D.cg(exec_, A, [gko.stop.Iteration(300), gko.stop.ResidualNorm(1e-9)])
    dxs = gko.matrix.Dense.create(exec_, (n, 1))
    s.apply(gko.matrix.Dense.from_numpy(exec_, b), dxs)
    assert abs(s.num_iterations - it_ref) <= 2
    assert np.allclose(s.residual_history[:10], hist_ref[:10], rtol=1e-12)
    assert np.abs(dxs.to_numpy()[:, 0] - x_ref).max() <= 1e-9 * np.abs(x_ref).max()
    # distributed::Vector reductions (1 rank): norm2 = sqrt(all_reduce(local squared norm))
    v = D.Vector(comm, gko.matrix.Dense.from_numpy(exec_, b))
    res = gko.matrix.Dense.create(exec_, (1, 1))
    v.compute_norm2(res)
    assert np.isclose(res.to_numpy()[0, 0], np.linalg.norm(b), rtol=1e-14)


# exchange paths: everything over peer memory inside the kernels (default), peer-memory scalar
# all-reduce + NCCL halo, everything on NCCL
PATHS = {"fused-halo": {}, "p2p-allreduce+nccl-halo": {"GKOB200_FUSED_HALO": "0"}, "nccl": {"GKOB200_P2P": "0"}}


@pytest.mark.parametrize("path", list(PATHS))
@pytest.mark.parametrize("world", [2, 4, 8])
def test_multi_gpu_torchrun(world, path):
    if torch.cuda.device_count() < world:
        pytest.skip(f"needs {world} GPUs (run with gpurun --gpus {world})")
    port = 29600 + world + 16 * list(PATHS).index(path)
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}",
           "--master-addr", "127.0.0.1", "--master-port", str(port), os.path.join(ROOT, "tests", "dist_worker.py")]
    env = dict(os.environ, GKOB200_P2P_TIMEOUT_MS="20000", **PATHS[path])
    out = subprocess.run(cmd, cwd=ROOT, capture_output=True, text=True, timeout=900, env=env)
    assert out.returncode == 0, out.stdout[-3000:] + out.stderr[-3000:]
    assert f"DIST_OK world={world} path={path}" in out.stdout, out.stdout[-2000:]
