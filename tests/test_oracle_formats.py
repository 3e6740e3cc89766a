"""Pins the format / conversion part of the plain-C oracle: against literals of the
reference's unit tests and, bit for bit, against the compiled reference (oracle/_ref)."""
import numpy as np
import pytest
import scipy.sparse as sp


def rand_csr(n, m, density, seed, dtype=np.float64, idtype=np.int32, skew=False):
    rng = np.random.default_rng(seed)
    a = sp.random(n, m, density=density, random_state=rng, format="lil")
    if skew:
        for r in rng.choice(n, 4, replace=False):
            a[r, rng.choice(m, size=m // 2, replace=False)] = 1.0
        for r in rng.choice(n, 9, replace=False):
            a[r, :] = 0
    a = a.tocsr()
    a.data = rng.uniform(-1, 1, a.nnz)
    a.eliminate_zeros()
    a.sort_indices()
    return a.indptr.astype(idtype), a.indices.astype(idtype), a.data.astype(dtype)


def sellp_defined_mask(n, slice_size, sets):
    mask = np.zeros(int(sets[-1]) * slice_size, dtype=bool)
    for s in range(len(sets) - 1):
        rows_here = min(slice_size, n - s * slice_size)
        for i in range(int(sets[s]), int(sets[s + 1])):
            mask[i * slice_size: i * slice_size + rows_here] = True
    return mask


# reference/test/matrix/csr_kernels.cpp:98-118 fixture {{1,3,2},{0,5,0}} and its conversion
# KATs (:660-1042: ConvertsToSellp / ConvertsToEll / ConvertsToHybrid / ConvertsToCoo)
RP, CI, VA = np.array([0, 3, 4], np.int32), np.array([0, 1, 2, 1], np.int32), np.array([1.0, 3.0, 2.0, 5.0])


def test_kat_csr_to_coo(ora):
    rows, cols, vals = ora.csr_to_coo(RP, CI, VA)
    assert list(rows) == [0, 0, 0, 1] and list(cols) == [0, 1, 2, 1] and list(vals) == [1, 3, 2, 5]


def test_kat_csr_to_ell(ora):
    # assert_equal_to_mtx(Ell): 3 stored per row, stride 2, column-major, padding (col -1? no:
    # the reference test checks c[4]=invalid_index, v[4]=0) — reference/test/matrix/csr_kernels.cpp
    width, stride, cols, vals = ora.csr_to_ell(RP, CI, VA)
    assert (width, stride) == (3, 2)
    assert list(cols) == [0, 1, 1, -1, 2, -1]
    assert list(vals) == [1, 5, 3, 0, 2, 0]


def test_kat_csr_to_sellp(ora):
    # ConvertsToSellp: slice_size 64, stride_factor 1: slice_sets {0,3}, slice_lengths {3}
    sets, lens, cols, vals = ora.csr_to_sellp(RP, CI, VA)
    assert list(sets) == [0, 3] and list(lens) == [3]
    assert len(cols) == 3 * 64
    assert [cols[0], cols[1], cols[64], cols[65], cols[128], cols[129]] == [0, 1, 1, -1, 2, -1]
    assert [vals[0], vals[1], vals[64], vals[65], vals[128], vals[129]] == [1, 5, 3, 0, 2, 0]
    # ConvertsToSellpWithSliceSizeAndStrideFactor (slice_size 2, stride_factor 2): lengths padded to 4
    sets, lens, cols, vals = ora.csr_to_sellp(RP, CI, VA, slice_size=2, stride_factor=2)
    assert list(sets) == [0, 4] and list(lens) == [4] and len(cols) == 8
    assert list(cols) == [0, 1, 1, -1, 2, -1, -1, -1]


def test_kat_csr_to_hybrid(ora):
    # ConvertsToHybridByColumn2: column_limit(2): ELL 2 cols stride 2, COO holds (0,2,2.0)
    h = ora.csr_to_hybrid(RP, CI, VA, 3, "column_limit", 2)
    assert (h["ell_width"], h["ell_stride"]) == (2, 2)
    assert list(h["ell_cols"]) == [0, 1, 1, -1] and list(h["ell_vals"]) == [1, 5, 3, 0]
    assert list(h["coo_rows"]) == [0] and list(h["coo_cols"]) == [2] and list(h["coo_vals"]) == [2.0]
    # automatic on this 2-row matrix: min(sorted[0]=1, int(2*0.001)=0) = 0 -> everything in COO
    h = ora.csr_to_hybrid(RP, CI, VA, 3, "automatic")
    assert h["ell_width"] == 0 and list(h["coo_rows"]) == [0, 0, 0, 1]


def test_kat_prefix_sum(ora):
    # reference/test/components/prefix_sum_kernels.cpp: exclusive scan, last entry = total
    a = np.array([3, 0, 5, 1, 9999], dtype=np.int32)
    assert list(ora.prefix_sum(a)) == [0, 3, 3, 8, 9]


@pytest.mark.parametrize("idtype", [np.int32, np.int64])
@pytest.mark.parametrize("shape,density,skew", [((300, 211), 0.04, False), ((257, 400), 0.02, True), ((64, 64), 0.5, False)])
def test_conversions_match_reference_bit_for_bit(ora, refimpl, idtype, shape, density, skew):
    rp, ci, va = rand_csr(shape[0], shape[1], density, 42, idtype=idtype, skew=skew)
    n, m = shape
    r = refimpl.ref_convert(rp, ci, va, m, "ell")
    width, stride, cols, vals = ora.csr_to_ell(rp, ci, va)
    assert (width, stride) == (r["width"], r["stride"])
    assert np.array_equal(cols, r["cols"]) and np.array_equal(vals, r["vals"])
    for ss, sf in ((64, 1), (32, 4), (7, 3)):
        r = refimpl.ref_convert(rp, ci, va, m, "sellp", slice_size=ss, stride_factor=sf)
        sets, lens, cols, vals = ora.csr_to_sellp(rp, ci, va, ss, sf)
        assert np.array_equal(sets, r["slice_sets"]) and np.array_equal(lens, r["slice_lengths"])
        # slots of rows >= n in the last, partial slice are never written by the reference
        # kernel (common/unified/matrix/csr_kernels.cpp:137-165): compare the defined ones
        defined = sellp_defined_mask(n, ss, sets)
        assert np.array_equal(cols[defined], r["cols"][defined]) and np.array_equal(vals[defined], r["vals"][defined])
    r = refimpl.ref_convert(rp, ci, va, m, "coo")
    rows, cols, vals = ora.csr_to_coo(rp, ci, va)
    assert np.array_equal(rows, r["rows"]) and np.array_equal(cols, r["cols"])
    for kind, kw in (("automatic", {}), ("column_limit", dict(param=3)), ("imbalance_limit", dict(percent=0.6)),
                     ("imbalance_bounded_limit", dict(percent=0.9, ratio=0.05)), ("minimal_storage_limit", {})):
        r = refimpl.ref_convert(rp, ci, va, m, "hybrid", hyb_kind=kind, hyb_param=kw.get("param", 0),
                                percent=kw.get("percent", 0.8), ratio=kw.get("ratio", 0.0001))
        h = ora.csr_to_hybrid(rp, ci, va, m, kind, **kw)
        assert (h["ell_width"], h["ell_stride"]) == (r["ell_width"], r["ell_stride"]), kind
        for k in ("ell_cols", "ell_vals", "coo_rows", "coo_cols", "coo_vals"):
            assert np.array_equal(h[k], r[k]), (kind, k)


@pytest.mark.parametrize("dtype", [np.float64, np.float32])
@pytest.mark.parametrize("nrhs", [1, 3])
def test_format_spmv_matches_reference_bit_for_bit(ora, refimpl, dtype, nrhs):
    rp, ci, va = rand_csr(300, 211, 0.05, 7, dtype, skew=True)
    rng = np.random.default_rng(1)
    b = rng.uniform(-1, 1, (211, nrhs)).astype(dtype)
    c0 = rng.uniform(-1, 1, (300, nrhs)).astype(dtype)
    width, stride, ecols, evals = ora.csr_to_ell(rp, ci, va)
    sets, lens, scols, svals = ora.csr_to_sellp(rp, ci, va)
    rows, ccols, cvals = ora.csr_to_coo(rp, ci, va)
    h = ora.csr_to_hybrid(rp, ci, va, 211, "column_limit", 5)
    for alpha, beta, c in ((None, None, None), (0.7, -1.3, c0)):
        want = {f: refimpl.ref_spmv(rp, ci, va, b, alpha=alpha, beta=beta, c=c, fmt=f,
                                    hybrid_limit=5 if f == "hybrid" else -1)[0] for f in ("ell", "sellp", "coo", "hybrid")}
        assert np.array_equal(ora.ell_spmv(300, stride, width, ecols, evals, b, alpha, beta, c), want["ell"])
        assert np.array_equal(ora.sellp_spmv(300, 64, sets, lens, scols, svals, b, alpha, beta, c), want["sellp"])
        # coo::spmv = fill/scale + spmv2 (reference/matrix/coo_kernels.cpp:62-89)
        base = np.zeros((300, nrhs), dtype) if c is None else (c * dtype(beta)).astype(dtype)
        assert np.array_equal(ora.coo_spmv2(rows, ccols, cvals, b, base, alpha), want["coo"])
        # Hybrid = ELL apply then COO apply2 (core/matrix/hybrid.cpp:133-160)
        e = ora.ell_spmv(300, h["ell_stride"], h["ell_width"], h["ell_cols"], h["ell_vals"], b, alpha, beta, c)
        assert np.array_equal(ora.coo_spmv2(h["coo_rows"], h["coo_cols"], h["coo_vals"], b, e, alpha), want["hybrid"])
