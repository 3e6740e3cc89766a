"""MatrixMarket / GINKGO-binary reader and writer of the C-ABI (host code) against the
reference's own gko::read_generic_raw / write_raw / write_binary_raw (oracle/_ref), on files of
every header combination the reference accepts (core/base/mtx_io.cpp:690-722) — the layouts
the reference's tests exercise in core/test/base/mtx_io.cpp."""
import numpy as np
import pytest

import oracle
from __graft_entry__ import load_package

needs_ref = pytest.mark.skipif(oracle.ref() is None, reason="oracle/_ref not built")


@pytest.fixture(scope="module")
def gko():
    return load_package()


def same(got, want, values_exact=True):
    (gs, gr, gc, gv), (ws, wr, wc, wv) = got, want
    assert tuple(gs) == tuple(ws)
    assert np.array_equal(gr, wr) and np.array_equal(gc, wc)
    assert np.array_equal(gv, wv) if values_exact else np.allclose(gv, wv, rtol=1e-6)


# the fixtures of core/test/base/mtx_io.cpp (ReadsDenseDoubleRealMtx, ReadsSparseRealMtx,
# ReadsSparseRealSymetricMtx, ReadsSparseRealSkewSymetricMtx, ReadsSparsePatternMtx, ...)
FILES = {
    "dense_real": ("%%MatrixMarket matrix array real general\n2 3\n1.0\n0.0\n3.0\n5.0\n2.0\n0.0\n",
                   [(0, 0, 1.0), (0, 1, 3.0), (0, 2, 2.0), (1, 0, 0.0), (1, 1, 5.0), (1, 2, 0.0)]),
    "dense_int": ("%%MatrixMarket matrix array integer general\n2 3\n1\n0\n3\n5\n2\n0\n",
                  [(0, 0, 1.0), (0, 1, 3.0), (0, 2, 2.0), (1, 0, 0.0), (1, 1, 5.0), (1, 2, 0.0)]),
    "sparse_real": ("%%MatrixMarket matrix coordinate real general\n2 3 4\n1 1 1.0\n2 2 5.0\n1 2 3.0\n1 3 2.0\n",
                    [(0, 0, 1.0), (0, 1, 3.0), (0, 2, 2.0), (1, 1, 5.0)]),
    "sparse_sym": ("%%MatrixMarket matrix coordinate real symmetric\n3 3 4\n1 1 1.0\n2 1 2.0\n3 3 6.0\n3 1 3.0\n",
                   [(0, 0, 1.0), (0, 1, 2.0), (0, 2, 3.0), (1, 0, 2.0), (2, 0, 3.0), (2, 2, 6.0)]),
    "sparse_skew": ("%%MatrixMarket matrix coordinate real skew-symmetric\n3 3 2\n2 1 2.0\n3 1 3.0\n",
                    [(0, 1, -2.0), (0, 2, -3.0), (1, 0, 2.0), (2, 0, 3.0)]),
    "sparse_pattern": ("%%MatrixMarket matrix coordinate pattern general\n2 3 4\n1 1\n2 2\n1 2\n1 3\n",
                       [(0, 0, 1.0), (0, 1, 1.0), (0, 2, 1.0), (1, 1, 1.0)]),
    "sparse_herm": ("%%MatrixMarket matrix coordinate real hermitian\n3 3 3\n1 1 1.0\n2 1 2.0\n3 2 -1.5\n",
                    [(0, 0, 1.0), (0, 1, 2.0), (1, 0, 2.0), (1, 2, -1.5), (2, 1, -1.5)]),
    "comments_case": ("%%matrixMARKET MATRIX Coordinate Integer General\n% a comment\n%another\n2 2 2\n2 1 -7\n1 2 4e0\n",
                      [(0, 1, 4.0), (1, 0, -7.0)]),
    "dense_sym": ("%%MatrixMarket matrix array real symmetric\n3 3\n1\n2\n3\n4\n5\n6\n",
                  [(0, 0, 1.0), (0, 1, 2.0), (0, 2, 3.0), (1, 0, 2.0), (1, 1, 4.0), (1, 2, 5.0), (2, 0, 3.0), (2, 1, 5.0),
                   (2, 2, 6.0)]),
    "dense_skew": ("%%MatrixMarket matrix array real skew-symmetric\n3 3\n2\n3\n5\n",
                   [(0, 1, -2.0), (0, 2, -3.0), (1, 0, 2.0), (1, 2, -5.0), (2, 0, 3.0), (2, 1, 5.0)]),
}


@pytest.mark.parametrize("name", sorted(FILES))
def test_reads_literal_files(gko, tmp_path, name):
    text, entries = FILES[name]
    p = tmp_path / f"{name}.mtx"
    p.write_text(text)
    size, r, c, v = gko.io.read_raw(p)
    assert list(zip(r.tolist(), c.tolist(), v.tolist())) == entries
    if oracle.ref() is not None:
        same((size, r, c, v), oracle.ref_mtx_read(p))


@needs_ref
@pytest.mark.parametrize("modifier", ["general", "symmetric", "skew-symmetric", "hermitian"])
@pytest.mark.parametrize("entry", ["real", "integer", "pattern"])
def test_random_coordinate_files_match_reference(gko, tmp_path, modifier, entry):
    rng = np.random.default_rng(hash((modifier, entry)) % 2 ** 32)
    n = 57
    pairs = {(int(a), int(b)) for a, b in rng.integers(0, n, (400, 2))}
    if modifier != "general":
        pairs = {(max(a, b), min(a, b)) for a, b in pairs}
        if modifier == "skew-symmetric":
            pairs = {(a, b) for a, b in pairs if a != b}
    lines = [f"%%MatrixMarket matrix coordinate {entry} {modifier}", "% generated", f"{n} {n} {len(pairs)}"]
    for a, b in sorted(pairs, key=lambda t: rng.random()):
        val = "" if entry == "pattern" else (f" {rng.integers(-9, 10)}" if entry == "integer" else f" {rng.standard_normal():.17g}")
        lines.append(f"{a + 1} {b + 1}{val}")
    p = tmp_path / "m.mtx"
    p.write_text("\n".join(lines) + "\n")
    same(gko.io.read_raw(p, index_dtype=np.int64), oracle.ref_mtx_read(p))
    # narrower types: same entries, values rounded to fp32
    size, r, c, v = gko.io.read_raw(p, value_dtype=np.float32, index_dtype=np.int32)
    _, wr, wc, wv = oracle.ref_mtx_read(p)
    assert np.array_equal(r, wr) and np.array_equal(c, wc) and np.array_equal(v, wv.astype(np.float32))


@needs_ref
@pytest.mark.parametrize("entry", ["real", "pattern"])
def test_large_file_with_free_token_layout(gko, tmp_path, entry):
    """> 1 MB: the parallel parser cuts the body at token boundaries.  The grammar is a token
    stream — several entries per line, entries split across lines, tabs, blank lines — and the
    entries arrive unsorted."""
    rng = np.random.default_rng(11)
    n, nnz = 5000, 150_000
    key = rng.choice(n * n, nnz, replace=False)
    toks = []
    for k in key:
        toks += [str(k // n + 1), str(k % n + 1)] + ([] if entry == "pattern" else [f"{rng.standard_normal():.17g}"])
    seps = rng.choice([" ", "\n", "\t", "  ", "\n\n", " \n"], len(toks))
    body = "".join(t + s for t, s in zip(toks, seps))
    p = tmp_path / "big.mtx"
    p.write_text(f"%%MatrixMarket matrix coordinate {entry} general\n%c\n{n} {n} {nnz}\n" + body + "\n")
    assert p.stat().st_size > (1 << 20)
    same(gko.io.read_raw(p, index_dtype=np.int64), oracle.ref_mtx_read(p))
    # truncated in the middle of an entry: both readers fail
    q = tmp_path / "cut.mtx"
    q.write_text(p.read_text()[: p.stat().st_size * 2 // 3])
    with pytest.raises(gko.Error):
        gko.io.read_raw(q)
    with pytest.raises(RuntimeError):
        oracle.ref_mtx_read(q)


@needs_ref
@pytest.mark.parametrize("index32", [True, False])
@pytest.mark.parametrize("value32", [True, False])
def test_binary_files_both_directions(gko, tmp_path, index32, value32):
    rng = np.random.default_rng(3)
    n, m, nnz = 300, 211, 5000
    key = rng.choice(n * m, nnz, replace=False)
    rows, cols = key // m, key % m
    vals = rng.standard_normal(nnz).astype(np.float32 if value32 else np.float64)
    # written by the reference, read by us (file types I|L x S|D into every requested type)
    pr = tmp_path / "ref.bin"
    oracle.ref_mtx_write(pr, (n, m), rows, cols, vals, binary=True, index32=index32, value32=value32)
    want = oracle.ref_mtx_read(pr)
    same(gko.io.read_raw(pr, index_dtype=np.int64), want)
    size, r, c, v = gko.io.read_raw(pr, value_dtype=np.float32, index_dtype=np.int32)
    assert np.array_equal(r, want[1]) and np.array_equal(v, want[3].astype(np.float32))
    # written by us, read by the reference: byte-identical file
    po = tmp_path / "ours.bin"
    idt = np.int32 if index32 else np.int64
    gko.io.write_raw(po, (n, m), rows.astype(idt), cols.astype(idt), vals, layout=gko.io.BINARY)
    assert po.read_bytes() == pr.read_bytes()
    same(oracle.ref_mtx_read(po), want)


@needs_ref
def test_text_writer_matches_reference_bytes(gko, tmp_path):
    rng = np.random.default_rng(9)
    rows, cols = rng.integers(0, 40, 300), rng.integers(0, 50, 300)
    vals = rng.standard_normal(300) * 10.0 ** rng.integers(-12, 12, 300)
    vals[:5] = [0.0, 1.0, -2.5, 1e100, 123456789.0]
    pr, po = tmp_path / "ref.mtx", tmp_path / "ours.mtx"
    oracle.ref_mtx_write(pr, (40, 50), rows, cols, vals)
    gko.io.write_raw(po, (40, 50), rows.astype(np.int64), cols.astype(np.int64), vals)
    assert po.read_text() == pr.read_text()
    # full precision round trip through our own writer / reader
    gko.io.write_raw(po, (40, 50), rows.astype(np.int64), cols.astype(np.int64), vals, precision=17)
    size, r, c, v = gko.io.read_raw(po, index_dtype=np.int64)
    order = np.lexsort((np.arange(300), cols, rows))
    assert np.array_equal(r, rows[order]) and np.array_equal(c, cols[order]) and np.array_equal(v, vals[order])


@needs_ref
def test_array_writer_matches_reference_bytes(gko, tmp_path):
    """gko::write of a Dense (array layout): column-major with explicit zeros."""
    rng = np.random.default_rng(4)
    key = rng.choice(7 * 5, 12, replace=False)
    rows, cols, vals = key // 5, key % 5, rng.standard_normal(12)
    pr, po = tmp_path / "ref.mtx", tmp_path / "ours.mtx"
    oracle.ref_mtx_write(pr, (7, 5), rows, cols, vals, array=True)
    gko.io.write_raw(po, (7, 5), rows.astype(np.int64), cols.astype(np.int64), vals, layout=gko.io.ARRAY)
    assert po.read_text() == pr.read_text()
    size, r, c, v = gko.io.read_raw(po, index_dtype=np.int64)
    assert size == (7, 5) and len(v) == 35          # the array layout reads every entry back, zeros included
    same((size, r, c, v), oracle.ref_mtx_read(po))


@pytest.mark.parametrize("text,needle", [
    ("%%MatrixMarket matrix coordinate complex general\n1 1 1\n1 1 1.0 2.0\n", "complex"),
    ("%%MatrixMarket tensor coordinate real general\n1 1 1\n1 1 1.0\n", "header"),
    ("%%MatrixMarket matrix coordinate real general\n2 2 3\n1 1 1.0\n2 2\n", "entry"),
    ("%%MatrixMarket matrix coordinate real general\n2 2\n", "size"),
    ("GINKGOXX" + "\0" * 24, "magic"),
    ("GINK", "header"),
])
def test_errors_like_the_reference(gko, tmp_path, text, needle):
    p = tmp_path / "bad.mtx"
    p.write_bytes(text.encode())
    with pytest.raises(gko.Error) as e:
        gko.io.read_raw(p)
    assert needle in str(e.value)
    if oracle.ref() is not None:
        with pytest.raises(RuntimeError):
            oracle.ref_mtx_read(p)
    with pytest.raises(gko.Error):
        gko.io.read_raw(tmp_path / "does_not_exist.mtx")


def test_index_overflow_is_rejected(gko, tmp_path):
    p = tmp_path / "wide.mtx"
    p.write_text("%%MatrixMarket matrix coordinate pattern general\n3000000000 5 1\n2999999999 1\n")
    with pytest.raises(gko.Error):
        gko.io.read_raw(p, index_dtype=np.int32)
    size, r, c, v = gko.io.read_raw(p, index_dtype=np.int64)
    assert size == (3000000000, 5) and r.tolist() == [2999999998]


def _binary(n_rows, n_cols, nnz_header, triplets):
    import struct
    out = b"GINKGODI" + struct.pack("<QQQ", n_rows, n_cols, nnz_header)
    for r, c, v in triplets:
        out += struct.pack("<iid", r, c, v)
    return out


@pytest.mark.parametrize("payload,needle", [
    # header sizes are not trusted: a huge nnz with a tiny body fails at the first missing entry
    (b"%%MatrixMarket matrix coordinate real general\n2 2 4000000000000\n1 1 1.0\n", "entry 1"),
    (b"%%MatrixMarket matrix array real general\n3000000 3000000\n1.0\n", "entry 1"),
    # coordinates outside the declared matrix must never reach device assembly
    (b"%%MatrixMarket matrix coordinate real general\n2 2 2\n1 1 1.0\n3 1 2.0\n", "outside"),
    (b"%%MatrixMarket matrix coordinate real general\n2 2 1\n0 1 1.0\n", "outside"),
    (b"%%MatrixMarket matrix coordinate pattern general\n2 2 1\n1 5\n", "outside"),
    # binary: 32 + n * rec wraps around 2^64 for this n; and out-of-range entries
    (_binary(2, 2, (2 ** 64) // 16 + 1, [(0, 0, 1.0)]), "entry"),
    (_binary(2, 2, 1, [(2, 0, 1.0)]), "outside"),
    (_binary(2, 2, 1, [(0, -1, 1.0)]), "outside"),
])
def test_hostile_sizes_and_coordinates_fail_cleanly(gko, tmp_path, payload, needle):
    p = tmp_path / "hostile.mtx"
    p.write_bytes(payload)
    with pytest.raises(gko.Error) as e:
        gko.io.read_raw(p)
    assert needle in str(e.value)
