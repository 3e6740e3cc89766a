"""CUDA CSR SpMV / SpMM vs the oracle — through the C-ABI.  Parity rules:
the row-block ("classical") kernel keeps the oracle's summation order and rounding, so it
must be BIT-IDENTICAL; the merge-path kernel re-associates partial rows across threads,
so it is held to 1e-12 (fp64) / 1e-5 (fp32) relative to sum_j |a_ij||b_j| per entry."""
import numpy as np
import pytest
import scipy.sparse as sp

import kat

pytestmark = pytest.mark.gpu

TOL = {np.float64: 1e-12, np.float32: 1e-5}


def random_csr(n, m, density, seed, dtype=np.float64, idtype=np.int32, skew=False):
    rng = np.random.default_rng(seed)
    a = sp.random(n, m, density=density, random_state=rng, format="lil", dtype=np.float64)
    if skew:  # a few very long rows and some empty ones
        for r in rng.choice(n, 5, replace=False):
            cols = rng.choice(m, size=min(m, 3000), replace=False)
            a[r, cols] = 1.0
        for r in rng.choice(n, 20, replace=False):
            a[r, :] = 0
    a = a.tocsr()
    a.data = rng.uniform(-1, 1, a.nnz)
    a.eliminate_zeros()
    a.sort_indices()
    return a.indptr.astype(idtype), a.indices.astype(idtype), a.data.astype(dtype)


def entry_bound(rp, ci, va, b, alpha=1.0, beta=0.0, c=None):
    """sum_j |a_ij||b_j| (+ |beta c|): the componentwise scale rounding errors are relative to."""
    A = sp.csr_matrix((np.abs(va.astype(np.float64)), ci, rp), shape=(len(rp) - 1, len(b)))
    s = abs(alpha) * (A @ np.abs(b.astype(np.float64).reshape(len(b), -1)))
    if c is not None:
        s = s + abs(beta) * np.abs(c.reshape(s.shape))
    return s + np.finfo(np.float64).tiny


def gpu_apply(gko, exec_, rp, ci, va, shape, b, strategy, alpha=None, beta=None, c=None):
    A = gko.matrix.Csr.from_arrays(exec_, shape, rp, ci, va, strategy=strategy)
    db = gko.matrix.Dense.from_numpy(exec_, b)
    k = db.size[1]
    dc = gko.matrix.Dense.create(exec_, (shape[0], k), db.t.dtype) if c is None else gko.matrix.Dense.from_numpy(exec_, c)
    if alpha is None:
        A.apply(db, dc)
    else:
        dt = db.t.dtype
        A.apply(gko.matrix.Dense.scalar(exec_, alpha, dt), db, gko.matrix.Dense.scalar(exec_, beta, dt), dc)
    return dc.to_numpy(), A


@pytest.mark.parametrize("dtype", [np.float64, np.float32])
@pytest.mark.parametrize("idtype", [np.int32, np.int64])
@pytest.mark.parametrize("strategy", ["classical", "merge_path"])
@pytest.mark.parametrize("case", kat.CSR_APPLY_KATS, ids=lambda c: c[0])
def test_reference_kats(gko, exec_, dtype, idtype, strategy, case):
    _, b, alpha, beta, c_in, expect = case
    m = kat.CSR_MTX
    rp, ci = np.array(m["row_ptrs"], dtype=idtype), np.array(m["col_idxs"], dtype=idtype)
    va = np.array(m["values"], dtype=dtype)
    b = np.array(b, dtype=dtype)
    if strategy == "merge_path" and b.shape[1] > 1:
        pytest.skip("SpMM always takes the row-block kernel")
    c = None if c_in is None else np.array(c_in, dtype=dtype)
    out, _ = gpu_apply(gko, exec_, rp, ci, va, m["shape"], b, strategy, alpha, beta, c)
    assert np.array_equal(out, np.array(expect, dtype=dtype))


@pytest.mark.parametrize("dtype", [np.float64, np.float32])
@pytest.mark.parametrize("idtype", [np.int32, np.int64])
@pytest.mark.parametrize("advanced", [False, True])
def test_classical_is_bit_identical_to_oracle(gko, exec_, ora, dtype, idtype, advanced):
    # 532 x 231, seed 42: the shape of the reference's device-vs-reference test
    # (test/matrix/csr_kernels2.cpp:63-90)
    rp, ci, va = random_csr(532, 231, 0.06, 42, dtype, idtype)
    rng = np.random.default_rng(1)
    b = rng.uniform(-1, 1, (231, 1)).astype(dtype)
    c0 = rng.uniform(-1, 1, (532, 1)).astype(dtype)
    if advanced:
        want = ora.csr_spmv(rp, ci, va, b, 0.37, -1.5, c0)
        got, A = gpu_apply(gko, exec_, rp, ci, va, (532, 231), b, "classical", 0.37, -1.5, c0)
    else:
        want = ora.csr_spmv(rp, ci, va, b)
        got, A = gpu_apply(gko, exec_, rp, ci, va, (532, 231), b, "classical")
    assert A.kernel() == "classical"
    assert np.array_equal(got, want)


@pytest.mark.parametrize("dtype", [np.float64, np.float32])
@pytest.mark.parametrize("nrhs", [2, 7, 32, 40])
def test_spmm_is_bit_identical_to_oracle(gko, exec_, ora, dtype, nrhs):
    rp, ci, va = random_csr(300, 257, 0.05, 5, dtype)
    rng = np.random.default_rng(2)
    b = rng.uniform(-1, 1, (257, nrhs)).astype(dtype)
    c0 = rng.uniform(-1, 1, (300, nrhs)).astype(dtype)
    got, _ = gpu_apply(gko, exec_, rp, ci, va, (300, 257), b, "classical")
    assert np.array_equal(got, ora.csr_spmv(rp, ci, va, b))
    got, _ = gpu_apply(gko, exec_, rp, ci, va, (300, 257), b, "classical", -2.0, 0.5, c0)
    assert np.array_equal(got, ora.csr_spmv(rp, ci, va, b, -2.0, 0.5, c0))


@pytest.mark.parametrize("dtype", [np.float64, np.float32])
@pytest.mark.parametrize("advanced", [False, True])
@pytest.mark.parametrize("shape,density,skew", [((2000, 1500), 0.01, True), ((1, 50), 0.5, False),
                                                ((5000, 5000), 0.0008, False), ((700, 9000), 0.3, False)])
def test_merge_path_within_tolerance(gko, exec_, ora, dtype, advanced, shape, density, skew):
    rp, ci, va = random_csr(shape[0], shape[1], density, 11, dtype, skew=skew)
    rng = np.random.default_rng(3)
    b = rng.uniform(-1, 1, (shape[1], 1)).astype(dtype)
    c0 = rng.uniform(-1, 1, (shape[0], 1)).astype(dtype)
    if advanced:
        want = ora.csr_spmv(rp, ci, va, b, 1.7, 0.25, c0)
        got, A = gpu_apply(gko, exec_, rp, ci, va, shape, b, "merge_path", 1.7, 0.25, c0)
        bound = entry_bound(rp, ci, va, b, 1.7, 0.25, c0)
    else:
        want = ora.csr_spmv(rp, ci, va, b)
        got, A = gpu_apply(gko, exec_, rp, ci, va, shape, b, "merge_path")
        bound = entry_bound(rp, ci, va, b)
    assert A.kernel() == "merge_path"
    err = np.abs(got.astype(np.float64) - want.astype(np.float64)) / bound
    assert err.max() <= TOL[dtype], err.max()


def test_automatical_picks_from_row_statistics(gko, exec_):
    rp, ci, va, n = gko.gen.stencil_csr("27pt", 20, 20, 20)
    A = gko.matrix.Csr.from_arrays(exec_, (n, n), rp, ci, va)
    assert A.kernel() == "classical" and A.max_row_nnz == 27
    assert A.max_block_nnz == max(rp[min(i + 128, n)] - rp[i] for i in range(0, n, 128))
    rp, ci, va = gko.gen.powerlaw_csr(30000, seed=1)
    A = gko.matrix.Csr.from_arrays(exec_, (30000, 30000), rp, ci, va)
    assert A.kernel() == "merge_path"
    assert A.max_row_nnz == int(np.diff(rp).max())


@pytest.mark.parametrize("strategy", ["classical", "merge_path"])
def test_empty_and_degenerate(gko, exec_, ora, strategy):
    # empty matrix: successful no-op (cuda/matrix/csr_kernels.cu:435-436)
    z = np.zeros(1, dtype=np.int32)
    A = gko.matrix.Csr.from_arrays(exec_, (0, 0), z, np.zeros(0, np.int32), np.zeros(0), strategy=strategy)
    A.apply(gko.matrix.Dense.create(exec_, (0, 1)), gko.matrix.Dense.create(exec_, (0, 1)))
    # all-empty rows must still overwrite c with zeros; rows with one entry; a dense row
    rp = np.array([0, 0, 0, 1, 1, 6, 6], dtype=np.int32)
    ci = np.array([2, 0, 1, 2, 3, 4], dtype=np.int32)
    va = np.arange(1.0, 7.0)
    b = np.arange(1.0, 6.0)[:, None]
    c = np.full((6, 1), 99.0)
    got, _ = gpu_apply(gko, exec_, rp, ci, va, (6, 5), b, strategy, c=c.copy())
    assert np.array_equal(got, ora.csr_spmv(rp, ci, va, b))
    got, _ = gpu_apply(gko, exec_, rp, ci, va, (6, 5), b, strategy, 2.0, 3.0, c.copy())
    assert np.array_equal(got, ora.csr_spmv(rp, ci, va, b, 2.0, 3.0, c))
    with pytest.raises(gko.Error):  # DimensionMismatch (lin_op.hpp:323-346)
        A2 = gko.matrix.Csr.from_arrays(exec_, (6, 5), rp, ci, va, strategy=strategy)
        A2.apply(gko.matrix.Dense.create(exec_, (6, 1)), gko.matrix.Dense.create(exec_, (6, 1)))


@pytest.mark.parametrize("kind,dims", [("5pt", (300, 200, 1)), ("7pt", (40, 50, 30)), ("27pt", (33, 41, 29))])
def test_stencils_bit_identical_and_linear(gko, exec_, ora, kind, dims):
    rp, ci, va, n = gko.gen.stencil_csr(kind, *dims)
    rng = np.random.default_rng(9)
    x, y = rng.standard_normal((n, 1)), rng.standard_normal((n, 1))
    ax, A = gpu_apply(gko, exec_, rp, ci, va, (n, n), x, "automatical")
    assert A.kernel() == "classical"
    assert np.array_equal(ax, ora.csr_spmv(rp, ci, va, x))
    # size-independent property: A(x + 2y) == Ax + 2Ay up to rounding
    ay, _ = gpu_apply(gko, exec_, rp, ci, va, (n, n), y, "merge_path")
    axy, _ = gpu_apply(gko, exec_, rp, ci, va, (n, n), x + 2 * y, "classical")
    assert np.allclose(axy, ax + 2 * ay, rtol=0, atol=1e-12 * 60)
    # A * 1 vanishes in the interior of a Laplacian-like stencil
    one, _ = gpu_apply(gko, exec_, rp, ci, va, (n, n), np.ones((n, 1)), "classical")
    assert np.count_nonzero(one) < n


@pytest.mark.parametrize("n,row_len", [(40, 30000), (3, 2304), (5000, 1), (2305, 0)])
def test_merge_path_rows_spanning_many_tiles(gko, exec_, ora, n, row_len):
    """Tiles of the merge-path kernel hold 2304 merge items: rows far longer than a tile (tiles
    without any row end, carries through several tiles), rows of exactly one tile, a tile of
    single-entry rows, and an all-empty matrix; every third row empty in between."""
    rng = np.random.default_rng(n)
    m = max(row_len, 7)
    lens = np.where(np.arange(n) % 3 == 1, 0, row_len).astype(np.int64)
    rp = np.concatenate([[0], np.cumsum(lens)]).astype(np.int32)
    ci = np.concatenate([np.sort(rng.choice(m, size=l, replace=False)) for l in lens] + [np.zeros(0, np.int64)]).astype(np.int32)
    va = rng.uniform(-1, 1, len(ci))
    b = rng.uniform(-1, 1, (m, 1))
    want = ora.csr_spmv(rp, ci, va, b)
    got, A = gpu_apply(gko, exec_, rp, ci, va, (n, m), b, "merge_path")
    assert A.kernel() == "merge_path"
    assert (np.abs(got - want) / entry_bound(rp, ci, va, b)).max() <= 1e-12
    c0 = rng.uniform(-1, 1, (n, 1))
    want = ora.csr_spmv(rp, ci, va, b, -0.5, 3.0, c0)
    got, _ = gpu_apply(gko, exec_, rp, ci, va, (n, m), b, "merge_path", -0.5, 3.0, c0)
    assert (np.abs(got - want) / entry_bound(rp, ci, va, b, -0.5, 3.0, c0)).max() <= 1e-12
    # twice through the same (planned) descriptor: carries of the first launch must not leak
    got2, _ = gpu_apply(gko, exec_, rp, ci, va, (n, m), b, "merge_path", -0.5, 3.0, c0)
    assert np.array_equal(got, got2)


@pytest.mark.parametrize("kind,dims", [("27pt", (16, 16, 16)), ("27pt", (24, 20, 12)), ("7pt", (32, 16, 8)),
                                       ("5pt", (64, 64, 1)), ("27pt", (17, 16, 16)), ("7pt", (16, 16, 18))])
@pytest.mark.parametrize("fmt", ["csr", "ell", "sellp"])
def test_spmm_lattice_tile_order_is_bit_identical(gko, exec_, ora, kind, dims, fmt):
    """Regular grids switch the SpMM kernel to the lattice-aware tile order (warps of a CTA own row
    tiles that are neighbours across grid lines / planes; spmm.cuh); grids that do not divide
    evenly keep the consecutive order.  Either way every (row, column) is summed in storage order:
    the result equals the oracle's bit for bit, with and without alpha / beta."""
    rp, ci, va, n = gko.gen.stencil_csr(kind, *dims)
    rng = np.random.default_rng(3)
    b = rng.standard_normal((n, 32))
    c0 = rng.standard_normal((n, 32))
    A = gko.matrix.Csr.from_arrays(exec_, (n, n), rp, ci, va)
    M = A if fmt == "csr" else A.convert_to(fmt)
    db = gko.matrix.Dense.from_numpy(exec_, b)
    dc = gko.matrix.Dense.create(exec_, (n, 32))
    M.apply(db, dc)
    assert np.array_equal(dc.to_numpy(), ora.csr_spmv(rp, ci, va, b))
    dc = gko.matrix.Dense.from_numpy(exec_, c0)
    M.apply(gko.matrix.Dense.scalar(exec_, -0.75), db, gko.matrix.Dense.scalar(exec_, 1.5), dc)
    assert np.array_equal(dc.to_numpy(), ora.csr_spmv(rp, ci, va, b, -0.75, 1.5, c0))
