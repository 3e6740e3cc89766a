"""Distributed index maps and the distributed apply restated on the CPU: pinned against
literal expectations of the reference's own tests, against the compiled reference kernels,
and (the reference's MPI-test approach, test/mpi/distributed/matrix.cpp:257-281) against a
replicated global CSR apply."""
import numpy as np
import pytest
import scipy.sparse as sp


def test_partition_kats(ora):
    # reference/test/distributed/partition_kernels.cpp BuildsFromGlobalSize: 13 rows over 5 parts
    p = ora.Partition.uniform(5, 13)
    assert list(p.bounds) == [0, 3, 6, 9, 11, 13] and list(p.sizes) == [3, 3, 3, 2, 2]
    # BuildsFromMapping {2,2,0,1,1,2,0,0,1,0,1,1,1,2,2,0}, 3 parts (partition_kernels.cpp:88-117)
    mapping = [2, 2, 0, 1, 1, 2, 0, 0, 1, 0, 1, 1, 1, 2, 2, 0]
    p = ora.Partition.from_mapping(mapping, 3)
    assert list(p.bounds) == [0, 2, 3, 5, 6, 8, 9, 10, 13, 15, 16]
    assert list(p.part_ids) == [2, 0, 1, 2, 0, 1, 0, 1, 2, 0]
    assert list(p.starts) == [0, 0, 0, 2, 1, 2, 3, 3, 3, 4]
    assert list(p.sizes) == [5, 6, 5] and p.num_empty_parts == 0


def random_global(n, density, seed):
    rng = np.random.default_rng(seed)
    a = sp.random(n, n, density=density, random_state=rng, format="csr")
    a.data = rng.uniform(-1, 1, a.nnz)
    a = (a + sp.eye(n)).tocsr()
    a.sort_indices()
    coo = a.tocoo()
    return a, coo.row.astype(np.int64), coo.col.astype(np.int64), coo.data.astype(np.float64)


@pytest.mark.parametrize("num_parts,use_mapping", [(1, False), (3, False), (4, True)])
def test_build_local_nonlocal_matches_reference_kernels(ora, refimpl, num_parts, use_mapping):
    n = 57
    A, rows, cols, vals = random_global(n, 0.08, 5)
    mapping = np.random.default_rng(2).integers(0, num_parts, n).astype(np.int32) if use_mapping else None
    part = ora.Partition.from_mapping(mapping, num_parts) if use_mapping else ora.Partition.uniform(num_parts, n)
    for lp in range(num_parts):   # loop over local_part in one process, like the reference's tests
        got = ora.dist_build_local_nonlocal(rows, cols, vals, part, lp)
        want, pm = refimpl.ref_dist_build(rows, cols, vals, lp, num_parts, n, mapping)
        assert np.array_equal(part.bounds, pm["bounds"]) and np.array_equal(part.part_ids, pm["part_ids"])
        assert np.array_equal(part.starts, pm["starts"]) and np.array_equal(part.sizes, pm["sizes"])
        for k in ("lrow", "lcol", "lval", "nrow", "ncol", "nval", "gather", "recv_sizes", "nl_to_global"):
            assert np.array_equal(got[k], want[k]), (lp, k)


@pytest.mark.parametrize("num_parts", [2, 3, 8])
@pytest.mark.parametrize("nrhs", [1, 2])
def test_distributed_apply_equals_global_apply(ora, num_parts, nrhs):
    n = 101
    A, rows, cols, vals = random_global(n, 0.06, 9)
    part = ora.Partition.uniform(num_parts, n)
    parts = [ora.dist_build_local_nonlocal(rows, cols, vals, part, p) for p in range(num_parts)]
    rng = np.random.default_rng(1)
    b = rng.standard_normal((n, nrhs))
    x0 = rng.standard_normal((n, nrhs))
    rp, ci, va = A.indptr.astype(np.int32), A.indices.astype(np.int32), A.data
    assert np.allclose(ora.dist_apply(parts, part, b), ora.csr_spmv(rp, ci, va, b), rtol=1e-13, atol=1e-14)
    assert np.allclose(ora.dist_apply(parts, part, b, 0.5, -2.0, x0), ora.csr_spmv(rp, ci, va, b, 0.5, -2.0, x0),
                       rtol=1e-13, atol=1e-14)
    # halo plan invariants: what p expects from q is what q sends to p; gather idxs are local rows of the sender
    send, recv, gathers = ora.dist_plan(parts)
    assert np.array_equal(send, recv.T)
    for p in range(num_parts):
        assert len(gathers[p]) == send[p].sum() and (len(gathers[p]) == 0 or gathers[p].max() < part.sizes[p])
