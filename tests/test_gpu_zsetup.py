"""Matrix assembly on the device against the oracle, through the C-ABI: every
integer output bit-exact, values bit-exact (duplicates are added in the oracle's order)."""
import numpy as np
import pytest
import scipy.sparse as sp
import torch

from test_oracle_setup import random_coo

pytestmark = pytest.mark.gpu


def npy(t):
    return t.detach().cpu().numpy()


@pytest.mark.parametrize("idtype", [np.int32, np.int64])
@pytest.mark.parametrize("dtype", [np.float64, np.float32])
@pytest.mark.parametrize("shape,nnz", [((100, 200), 3000), ((1, 1), 7), ((300, 7), 0), ((5, 5), 1),
                                       ((4000, 3000), 1_000_003)])
def test_assembly_steps_bit_exact(gko, exec_, ora, dtype, idtype, shape, nnz):
    rows, cols, vals = random_coo(shape[0], shape[1], nnz, 17, dtype, idtype)
    D = gko.assembly.DeviceMatrixData.from_arrays(exec_, shape, rows, cols, vals)
    # sort_row_major: stable, like the oracle's merge sort
    D.sort_row_major()
    want = ora.coo_assemble(rows, cols, vals, ora.SORT)
    assert np.array_equal(npy(D.row_idxs), want[0]) and np.array_equal(npy(D.col_idxs), want[1])
    assert np.array_equal(npy(D.values), want[2])
    # sum_duplicates (sorts first, like the reference), then remove_zeros
    D.sum_duplicates()
    want = ora.coo_assemble(rows, cols, vals, ora.SORT | ora.SUM_DUPLICATES)
    assert D.num_elems == len(want[2])
    assert np.array_equal(npy(D.row_idxs), want[0]) and np.array_equal(npy(D.col_idxs), want[1])
    assert np.array_equal(npy(D.values), want[2])
    D.remove_zeros()
    want = ora.coo_assemble(rows, cols, vals)
    assert D.num_elems == len(want[2])
    assert np.array_equal(npy(D.row_idxs), want[0]) and np.array_equal(npy(D.col_idxs), want[1])
    assert np.array_equal(npy(D.values), want[2])
    # Csr::read
    A = D.to_csr()
    rp = np.concatenate([[0], np.cumsum(np.bincount(want[0], minlength=shape[0]))]).astype(idtype)
    assert np.array_equal(npy(A.row_ptrs), rp)


def test_sum_duplicates_rounding_order(gko, exec_, ora):
    """Non-integer values: the duplicates of a position are added in input order from zero."""
    rng = np.random.default_rng(5)
    n = 200_000
    rows = rng.integers(0, 300, n).astype(np.int32)
    cols = rng.integers(0, 40, n).astype(np.int32)
    vals = rng.standard_normal(n) * 10.0 ** rng.integers(-8, 8, n)
    D = gko.assembly.DeviceMatrixData.from_arrays(exec_, (300, 40), rows, cols, vals).sum_duplicates()
    want = ora.coo_assemble(rows, cols, vals, ora.SORT | ora.SUM_DUPLICATES)
    assert np.array_equal(npy(D.values), want[2])


@pytest.mark.parametrize("idtype", [np.int32, np.int64])
@pytest.mark.parametrize("dtype", [np.float64, np.float32])
def test_transpose_and_sort_by_column_index(gko, exec_, ora, dtype, idtype):
    rng = np.random.default_rng(8)
    M = sp.random(3000, 1700, 0.01, random_state=4, format="csr", dtype=np.float64)
    M.data = rng.uniform(-1, 1, M.nnz)
    M.sort_indices()
    rp, ci, va = M.indptr.astype(idtype), M.indices.astype(idtype), M.data.astype(dtype)
    A = gko.matrix.Csr.from_arrays(exec_, M.shape, rp, ci, va, strategy="classical")
    T = gko.assembly.transpose(A)
    want = ora.csr_transpose(3000, 1700, rp, ci, va)
    assert T.size == (1700, 3000)
    assert np.array_equal(npy(T.row_ptrs), want[0]) and np.array_equal(npy(T.col_idxs), want[1])
    assert np.array_equal(npy(T.values), want[2])
    TT = gko.assembly.transpose(T)
    assert np.array_equal(npy(TT.row_ptrs), rp) and np.array_equal(npy(TT.col_idxs), ci) and np.array_equal(npy(TT.values), va)
    # shuffle the columns inside every row, sort them back
    ci_u, va_u = ci.copy(), va.copy()
    for r in range(3000):
        p = rng.permutation(rp[r + 1] - rp[r]) + rp[r]
        ci_u[rp[r]:rp[r + 1]], va_u[rp[r]:rp[r + 1]] = ci[p], va[p]
    B = gko.matrix.Csr.from_arrays(exec_, M.shape, rp, ci_u, va_u, strategy="classical")
    gko.assembly.sort_by_column_index(B)
    assert np.array_equal(npy(B.col_idxs), ci) and np.array_equal(npy(B.values), va)
    # empty matrix
    E = gko.matrix.Csr.from_arrays(exec_, (4, 6), np.zeros(5, idtype), np.zeros(0, idtype), np.zeros(0, dtype),
                                   strategy="classical")
    ET = gko.assembly.transpose(E)
    assert npy(ET.row_ptrs).tolist() == [0] * 7


def test_csr_read_of_shuffled_stencil_equals_generator(gko, exec_, ora):
    """C1's matrix assembled on the device from shuffled triplets == the generator's CSR; the
    assembled operator then gives the oracle's SpMV bit for bit."""
    rp, ci, va, n = gko.gen.stencil_csr("5pt", 1000, 1000)
    rows = np.repeat(np.arange(n, dtype=np.int32), np.diff(rp))
    p = np.random.default_rng(0).permutation(len(ci))
    A = gko.assembly.csr_read(exec_, (n, n), rows[p], ci[p], va[p])
    assert np.array_equal(npy(A.row_ptrs), rp) and np.array_equal(npy(A.col_idxs), ci) and np.array_equal(npy(A.values), va)
    x = np.random.default_rng(1).standard_normal((n, 1))
    y = gko.matrix.Dense.create(exec_, (n, 1))
    A.apply(gko.matrix.Dense.from_numpy(exec_, x), y)
    assert np.array_equal(y.to_numpy(), ora.csr_spmv(rp, ci, va, x))


@pytest.mark.parametrize("layout", ["coordinate", "binary"])
def test_file_to_device_csr(gko, exec_, ora, tmp_path, layout):
    """gko::read<Csr>(file, exec): written with write_raw / write_binary_raw in shuffled order, read
    back through the native reader and the device assembly == the generator's CSR."""
    rp, ci, va, n = gko.gen.stencil_csr("7pt", 30, 20, 10)
    rows = np.repeat(np.arange(n, dtype=np.int32), np.diff(rp))
    p = np.random.default_rng(2).permutation(len(ci))
    path = tmp_path / ("m.mtx" if layout == "coordinate" else "m.bin")
    gko.io.write_raw(path, (n, n), rows[p], ci[p], va[p], layout=gko.io.COORDINATE if layout == "coordinate" else gko.io.BINARY,
                     precision=17)
    A = gko.io.read(exec_, path)
    assert A.size == (n, n)
    assert np.array_equal(npy(A.row_ptrs), rp) and np.array_equal(npy(A.col_idxs), ci) and np.array_equal(npy(A.values), va)
    x = np.random.default_rng(1).standard_normal((n, 1))
    y = gko.matrix.Dense.create(exec_, (n, 1))
    A.apply(gko.matrix.Dense.from_numpy(exec_, x), y)
    assert np.array_equal(y.to_numpy(), ora.csr_spmv(rp, ci, va, x))
