"""world_size-2 gloo test on the CPU of the N>1 HOST logic: the halo-plan exchange of
distributed::Matrix::read_distributed (product function gko_b200.distributed.halo_plan) run
over a real process group, fed with the oracle's build_local_nonlocal output, and checked
against the single-process restatement oracle.dist_plan.  No GPU, no compute kernels."""
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


class GlooComm:
    """Test double with the two collectives halo_plan needs, on torch.distributed/gloo."""

    def __init__(self):
        self.rank, self.size = dist.get_rank(), dist.get_world_size()

    def all_to_all_i64(self, send):
        outs = [torch.zeros(1, dtype=torch.int64) for _ in range(self.size)]
        ins = [send[p: p + 1].clone() for p in range(self.size)]
        # gloo has no all_to_all: all_gather the whole vector and pick our column
        gathered = [torch.zeros_like(send) for _ in range(self.size)]
        dist.all_gather(gathered, send)
        return torch.stack([g[self.rank] for g in gathered])

    def all_to_all_v_i32(self, send, send_sizes, recv_sizes):
        mx = torch.tensor([send.numel()], dtype=torch.int64)
        dist.all_reduce(mx, op=dist.ReduceOp.MAX)
        pad = torch.zeros(int(mx.item()), dtype=torch.int32)
        pad[: send.numel()] = send
        all_send = [torch.zeros_like(pad) for _ in range(self.size)]
        dist.all_gather(all_send, pad)
        all_sizes = [torch.zeros(self.size, dtype=torch.int64) for _ in range(self.size)]
        dist.all_gather(all_sizes, torch.from_numpy(np.asarray(send_sizes, np.int64)))
        chunks = []
        for q in range(self.size):
            off = int(all_sizes[q][: self.rank].sum())
            chunks.append(all_send[q][off: off + int(all_sizes[q][self.rank])])
        return torch.cat(chunks) if chunks else torch.zeros(0, dtype=torch.int32)


def _worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        import scipy.sparse as sp
        import oracle
        from __graft_entry__ import load_package
        gko = load_package()
        n = 64
        rng = np.random.default_rng(3)
        a = (sp.random(n, n, density=0.1, random_state=rng) + sp.eye(n)).tocoo()
        order = np.lexsort((a.col, a.row))
        rows, cols, vals = a.row[order].astype(np.int64), a.col[order].astype(np.int64), a.data[order]
        part = oracle.Partition.uniform(world, n)
        mine = oracle.dist_build_local_nonlocal(rows, cols, vals, part, rank)
        ss, rs, gather = gko.distributed.halo_plan(GlooComm(), torch.from_numpy(mine["recv_sizes"]),
                                                   torch.from_numpy(mine["gather"]))
        allp = [oracle.dist_build_local_nonlocal(rows, cols, vals, part, p) for p in range(world)]
        send, recv, gathers = oracle.dist_plan(allp)
        ok = (np.array_equal(ss, send[rank]) and np.array_equal(rs, recv[rank])
              and np.array_equal(gather.numpy(), gathers[rank]))
        q.put((rank, bool(ok)))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 3])
def test_halo_plan_over_gloo(world):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + os.getpid() % 2000 + world
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(timeout=180)
    results = sorted(q.get(timeout=5) for _ in range(world))
    assert results == [(r, True) for r in range(world)]
