"""ELL / SELL-P / COO / Hybrid on the GPU vs the oracle, through the C-ABI.
Conversions and all integer outputs: bit-exact.  ELL / SELL-P SpMV keep the oracle's
order: bit-exact.  COO re-associates partial rows: 1e-12 / 1e-5 relative to sum|a||b|."""
import numpy as np
import pytest
import torch

from test_gpu_csr import entry_bound, random_csr
from test_oracle_formats import sellp_defined_mask

pytestmark = pytest.mark.gpu
TOL = {np.float64: 1e-12, np.float32: 1e-5}


def npy(t):
    return t.detach().cpu().numpy()


def make(gko, exec_, shape, density, seed, dtype=np.float64, idtype=np.int32, skew=False):
    rp, ci, va = random_csr(shape[0], shape[1], density, seed, dtype, idtype, skew=skew)
    return rp, ci, va, gko.matrix.Csr.from_arrays(exec_, shape, rp, ci, va)


@pytest.mark.parametrize("idtype", [np.int32, np.int64])
@pytest.mark.parametrize("shape,density,skew", [((300, 211), 0.04, False), ((1000, 900), 0.01, True), ((64, 64), 0.5, False),
                                                ((5, 3), 0.9, False)])
def test_conversions_bit_exact(gko, exec_, ora, idtype, shape, density, skew):
    rp, ci, va, A = make(gko, exec_, shape, density, 42, idtype=idtype, skew=skew)
    n, m = shape
    E = A.convert_to("ell")
    width, stride, cols, vals = ora.csr_to_ell(rp, ci, va)
    assert (E.width, E.stride) == (width, stride)
    assert np.array_equal(npy(E.col_idxs), cols) and np.array_equal(npy(E.values), vals)
    for ss, sf in ((64, 1), (32, 4), (7, 3)):
        S = A.convert_to("sellp", slice_size=ss, stride_factor=sf)
        sets, lens, cols, vals = ora.csr_to_sellp(rp, ci, va, ss, sf)
        assert np.array_equal(npy(S.slice_sets).astype(np.uint64), sets)
        assert np.array_equal(npy(S.slice_lengths).astype(np.uint64), lens)
        d = sellp_defined_mask(n, ss, sets)
        assert np.array_equal(npy(S.col_idxs)[d], cols[d]) and np.array_equal(npy(S.values)[d], vals[d])
    C = A.convert_to("coo")
    assert np.array_equal(npy(C.row_idxs), ora.csr_to_coo(rp, ci, va)[0])
    S = gko.matrix.HybridStrategy
    for strat, kind, kw in ((S.automatic(), "automatic", {}), (S.column_limit(3), "column_limit", dict(param=3)),
                            (S.imbalance_limit(0.6), "imbalance_limit", dict(percent=0.6)),
                            (S.imbalance_limit(1.0), "imbalance_limit", dict(percent=1.0)),
                            (S.imbalance_bounded_limit(0.9, 0.05), "imbalance_bounded_limit", dict(percent=0.9, ratio=0.05)),
                            (S.minimal_storage_limit(8, rp.itemsize), "minimal_storage_limit", {})):
        H = A.convert_to("hybrid", strategy=strat)
        h = ora.csr_to_hybrid(rp, ci, va, m, kind, **kw)
        assert (H.ell.width, H.ell.stride) == (h["ell_width"], h["ell_stride"]), kind
        assert np.array_equal(npy(H.ell.col_idxs), h["ell_cols"]) and np.array_equal(npy(H.ell.values), h["ell_vals"])
        assert np.array_equal(npy(H.coo.row_idxs), h["coo_rows"]) and np.array_equal(npy(H.coo.col_idxs), h["coo_cols"])
        assert np.array_equal(npy(H.coo.values), h["coo_vals"])


@pytest.mark.parametrize("n", [0, 1, 5, 4096, 4097, 100_000, 17_000_001])
@pytest.mark.parametrize("dt", [torch.int32, torch.int64])
def test_prefix_sum_bit_exact(gko, exec_, ora, n, dt):
    rng = np.random.default_rng(n)
    a = rng.integers(0, 100, size=n).astype(np.int32 if dt == torch.int32 else np.int64)
    t = torch.from_numpy(a).to(exec_.device)
    gko.matrix.prefix_sum(exec_, t)
    assert np.array_equal(npy(t), ora.prefix_sum(a))


def test_ptrs_idxs_round_trip(gko, exec_, ora):
    rp, ci, va, A = make(gko, exec_, (3000, 100), 0.03, 3, skew=True)
    for I, dt in (("i32", torch.int32), ("i64", torch.int64)):
        ptrs = A.row_ptrs.to(dt)
        idxs = torch.empty(A.nnz, dtype=dt, device=exec_.device)
        gko.check(getattr(gko.lib, f"gkob200_convert_ptrs_to_idxs_{I}")(gko.current_stream(), ptrs.data_ptr(), 3000,
                                                                         idxs.data_ptr()))
        assert np.array_equal(npy(idxs), np.repeat(np.arange(3000), np.diff(rp)))
        back = torch.full((3001,), -7, dtype=dt, device=exec_.device)
        gko.check(getattr(gko.lib, f"gkob200_convert_idxs_to_ptrs_{I}")(gko.current_stream(), idxs.data_ptr(), A.nnz,
                                                                         3000, back.data_ptr()))
        assert np.array_equal(npy(back), rp)
        empty = torch.full((11,), -7, dtype=dt, device=exec_.device)
        gko.check(getattr(gko.lib, f"gkob200_convert_idxs_to_ptrs_{I}")(gko.current_stream(), None, 0, 10,
                                                                         empty.data_ptr()))
        assert not npy(empty).any()


@pytest.mark.parametrize("dtype", [np.float64, np.float32])
@pytest.mark.parametrize("idtype", [np.int32, np.int64])
@pytest.mark.parametrize("nrhs", [1, 3, 32, 33])
@pytest.mark.parametrize("fmt", ["ell", "sellp"])
def test_ell_sellp_bit_identical(gko, exec_, ora, dtype, idtype, nrhs, fmt):
    rp, ci, va, A = make(gko, exec_, (523, 311), 0.04, 5, dtype, idtype, skew=True)
    M = A.convert_to(fmt)
    rng = np.random.default_rng(2)
    b = rng.uniform(-1, 1, (311, nrhs)).astype(dtype)
    c0 = rng.uniform(-1, 1, (523, nrhs)).astype(dtype)
    tdt = torch.float64 if dtype == np.float64 else torch.float32
    for alpha, beta in ((None, None), (0.37, -1.5)):
        if fmt == "ell":
            w, st, cols, vals = ora.csr_to_ell(rp, ci, va)
            want = ora.ell_spmv(523, st, w, cols, vals, b, alpha, beta, c0 if alpha else None)
        else:
            sets, lens, cols, vals = ora.csr_to_sellp(rp, ci, va)
            want = ora.sellp_spmv(523, 64, sets, lens, cols, vals, b, alpha, beta, c0 if alpha else None)
        db = gko.matrix.Dense.from_numpy(exec_, b)
        dc = gko.matrix.Dense.from_numpy(exec_, c0)
        if alpha is None:
            M.apply(db, dc)
        else:
            M.apply(gko.matrix.Dense.scalar(exec_, alpha, tdt), db, gko.matrix.Dense.scalar(exec_, beta, tdt), dc)
        assert np.array_equal(dc.to_numpy(), want)
        # and identical to CSR's result (same storage order per row)
        assert np.array_equal(want, ora.csr_spmv(rp, ci, va, b, alpha, beta, c0 if alpha else None))


@pytest.mark.parametrize("dtype", [np.float64, np.float32])
@pytest.mark.parametrize("shape,density,skew", [((523, 311), 0.04, True), ((3, 5000), 0.9, False), ((40000, 300), 0.002, False),
                                                ((1, 1), 1.0, False)])
@pytest.mark.parametrize("nrhs", [1, 2])
def test_coo_within_tolerance(gko, exec_, ora, dtype, shape, density, skew, nrhs):
    rp, ci, va, A = make(gko, exec_, shape, density, 8, dtype, skew=skew)
    C = A.convert_to("coo")
    rows = np.repeat(np.arange(shape[0]), np.diff(rp)).astype(np.int32)
    rng = np.random.default_rng(4)
    b = rng.uniform(-1, 1, (shape[1], nrhs)).astype(dtype)
    c0 = rng.uniform(-1, 1, (shape[0], nrhs)).astype(dtype)
    tdt = torch.float64 if dtype == np.float64 else torch.float32
    db = gko.matrix.Dense.from_numpy(exec_, b)
    # apply (c = A b), advanced apply, apply2 (c += A b), advanced apply2
    dc = gko.matrix.Dense.from_numpy(exec_, c0)
    C.apply(db, dc)
    want = ora.coo_spmv2(rows, ci, va, b, np.zeros_like(c0))
    bound = entry_bound(rp, ci, va, b)
    assert (np.abs(dc.to_numpy().astype(np.float64) - want) / bound).max() <= TOL[dtype]
    dc = gko.matrix.Dense.from_numpy(exec_, c0)
    C.apply(gko.matrix.Dense.scalar(exec_, -0.5, tdt), db, gko.matrix.Dense.scalar(exec_, 2.0, tdt), dc)
    want = ora.coo_spmv2(rows, ci, va, b, (c0 * dtype(2.0)).astype(dtype), -0.5)
    bound = entry_bound(rp, ci, va, b, 0.5, 2.0, c0)
    assert (np.abs(dc.to_numpy().astype(np.float64) - want) / bound).max() <= TOL[dtype]
    dc = gko.matrix.Dense.from_numpy(exec_, c0)
    C.apply2(db, dc)
    want = ora.coo_spmv2(rows, ci, va, b, c0)
    bound = entry_bound(rp, ci, va, b, 1.0, 1.0, c0)
    assert (np.abs(dc.to_numpy().astype(np.float64) - want) / bound).max() <= TOL[dtype]
    # deterministic: a second run gives the same bits
    dc2 = gko.matrix.Dense.from_numpy(exec_, c0)
    C.apply2(db, dc2)
    assert np.array_equal(dc.to_numpy(), dc2.to_numpy())


@pytest.mark.parametrize("dtype", [np.float64, np.float32])
@pytest.mark.parametrize("limit", [0, 3, 16, 10 ** 6])
def test_hybrid_apply(gko, exec_, ora, dtype, limit):
    rp, ci, va, A = make(gko, exec_, (700, 650), 0.02, 9, dtype, skew=True)
    H = A.convert_to("hybrid", strategy=gko.matrix.HybridStrategy.column_limit(limit))
    rng = np.random.default_rng(5)
    b = rng.uniform(-1, 1, (650, 1)).astype(dtype)
    c0 = rng.uniform(-1, 1, (700, 1)).astype(dtype)
    tdt = torch.float64 if dtype == np.float64 else torch.float32
    h = ora.csr_to_hybrid(rp, ci, va, 650, "column_limit", limit)
    for alpha, beta in ((None, None), (1.25, 0.5)):
        e = ora.ell_spmv(700, h["ell_stride"], h["ell_width"], h["ell_cols"], h["ell_vals"], b, alpha, beta,
                         c0 if alpha else None)
        want = ora.coo_spmv2(h["coo_rows"], h["coo_cols"], h["coo_vals"], b, e, alpha)
        db, dc = gko.matrix.Dense.from_numpy(exec_, b), gko.matrix.Dense.from_numpy(exec_, c0)
        if alpha is None:
            H.apply(db, dc)
        else:
            H.apply(gko.matrix.Dense.scalar(exec_, alpha, tdt), db, gko.matrix.Dense.scalar(exec_, beta, tdt), dc)
        bound = entry_bound(rp, ci, va, b, alpha or 1.0, beta or 0.0, c0 if alpha else None)
        assert (np.abs(dc.to_numpy().astype(np.float64) - want) / bound).max() <= TOL[dtype]
        if limit >= A.max_row_nnz:  # empty COO part: pure ELL, bit-identical
            assert np.array_equal(dc.to_numpy(), want)


def test_stencil_formats_agree_and_cg_on_sellp(gko, exec_, ora):
    rp, ci, va, n = gko.gen.stencil_csr("27pt", 20, 21, 22)
    A = gko.matrix.Csr.from_arrays(exec_, (n, n), rp, ci, va)
    x = np.random.default_rng(6).standard_normal((n, 1))
    want = ora.csr_spmv(rp, ci, va, x)
    dx = gko.matrix.Dense.from_numpy(exec_, x)
    for fmt in ("ell", "sellp"):
        M = A.convert_to(fmt)
        dy = gko.matrix.Dense.create(exec_, (n, 1))
        M.apply(dx, dy)
        assert np.array_equal(dy.to_numpy(), want), fmt
    # automatic hybrid = min(sorted_len[n/3], n/1000): here 9 -> the other 18 columns go to COO
    H = A.convert_to("hybrid")
    assert H.ell.width == ora.hybrid_ell_width(rp, n, "automatic") == 9
    dy = gko.matrix.Dense.create(exec_, (n, 1))
    H.apply(dx, dy)
    assert (np.abs(dy.to_numpy() - want) / entry_bound(rp, ci, va, x)).max() <= 1e-12
    # from 30^3 rows on, automatic hybrid on a 27-pt stencil is pure ELL
    rp2, ci2, va2, n2 = gko.gen.stencil_csr("27pt", 30, 30, 30)
    H2 = gko.matrix.Csr.from_arrays(exec_, (n2, n2), rp2, ci2, va2).convert_to("hybrid")
    assert H2.ell.width == 27 and H2.coo.values.numel() == 0
    # CG + scalar Jacobi with the SELL-P operator: fused-dot path of the strided kernel
    S = A.convert_to("sellp")
    M = gko.preconditioner.Jacobi.build().with_max_block_size(1).on(exec_).generate(A)
    solver = (gko.solver.Cg.build().with_criteria(gko.stop.Iteration(400), gko.stop.ResidualNorm(1e-9))
              .with_generated_preconditioner(M).on(exec_).generate(S))
    b = np.random.default_rng(7).standard_normal(n)
    x_ref, it_ref, hist_ref, _ = ora.cg_solve(rp, ci, va, b, np.zeros(n), precond=1, inv_diag=1 / np.full(n, 26.0),
                                              max_iters=400, factor=1e-9)
    dxs = gko.matrix.Dense.create(exec_, (n, 1))
    solver.apply(gko.matrix.Dense.from_numpy(exec_, b), dxs)
    assert abs(solver.num_iterations - it_ref) <= 2
    assert np.abs(dxs.to_numpy()[:, 0] - x_ref).max() <= 1e-9 * np.abs(x_ref).max()


@pytest.mark.parametrize("dtype", [np.float64, np.float32])
@pytest.mark.parametrize("nrhs", [2, 32, 37, 64])
def test_spmm_stencil_all_formats_bit_identical(gko, exec_, ora, dtype, nrhs):
    """SpMM tile kernel (spmm.cuh): clean tiles (all rows of a warp tile equally long — the
    predicate-free path), ragged boundary tiles, a partial last tile, several column blocks."""
    rp, ci, va, n = gko.gen.stencil_csr("27pt", 13, 11, 9, value_dtype=dtype)
    A = gko.matrix.Csr.from_arrays(exec_, (n, n), rp, ci, va)
    rng = np.random.default_rng(12)
    b = rng.uniform(-1, 1, (n, nrhs)).astype(dtype)
    c0 = rng.uniform(-1, 1, (n, nrhs)).astype(dtype)
    tdt = torch.float64 if dtype == np.float64 else torch.float32
    db = gko.matrix.Dense.from_numpy(exec_, b)
    for alpha, beta in ((None, None), (-0.75, 2.0)):
        want = ora.csr_spmv(rp, ci, va, b, alpha, beta, c0 if alpha else None)
        for fmt in ("csr", "ell", "sellp", "hybrid"):
            M = A if fmt == "csr" else A.convert_to(fmt, strategy=gko.matrix.HybridStrategy.column_limit(27)) \
                if fmt == "hybrid" else A.convert_to(fmt)
            dc = gko.matrix.Dense.from_numpy(exec_, c0)
            if alpha is None:
                M.apply(db, dc)
            else:
                M.apply(gko.matrix.Dense.scalar(exec_, alpha, tdt), db, gko.matrix.Dense.scalar(exec_, beta, tdt), dc)
            assert np.array_equal(dc.to_numpy(), want), (fmt, alpha)


@pytest.mark.parametrize("fmt", ["csr", "ell", "sellp"])
def test_spmm_rows_longer_than_the_staging_tile(gko, exec_, ora, fmt):
    """Rows above 36 entries take the unstaged path of the SpMM kernel; mixed with short and
    empty rows in the same matrix."""
    rp, ci, va, A = make(gko, exec_, (700, 650), 0.02, 21, skew=True)
    M = A if fmt == "csr" else A.convert_to(fmt)
    rng = np.random.default_rng(5)
    b = rng.uniform(-1, 1, (650, 33))
    c0 = rng.uniform(-1, 1, (700, 33))
    db = gko.matrix.Dense.from_numpy(exec_, b)
    dc = gko.matrix.Dense.from_numpy(exec_, c0)
    M.apply(db, dc)
    assert np.array_equal(dc.to_numpy(), ora.csr_spmv(rp, ci, va, b))
    dc = gko.matrix.Dense.from_numpy(exec_, c0)
    M.apply(gko.matrix.Dense.scalar(exec_, 1.25), db, gko.matrix.Dense.scalar(exec_, -0.5), dc)
    assert np.array_equal(dc.to_numpy(), ora.csr_spmv(rp, ci, va, b, 1.25, -0.5, c0))
