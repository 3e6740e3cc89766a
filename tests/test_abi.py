"""CPU-side checks of the drop-in boundary: the C-ABI library loads and exports every
symbol that include/*.h declares (no compute calls here)."""
import ctypes
import os
import re
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    syms = set()
    for header in ("gko_b200.h", "gko_b200_solver.h"):
        out = subprocess.run(["gcc", "-E", "-P", "-I", os.path.join(ROOT, "include"), os.path.join(ROOT, "include", header)],
                             check=True, capture_output=True, text=True).stdout
        syms |= set(re.findall(r"\b(gkob200_[a-z0-9_]+)\s*\(", out))
    return sorted(syms)


def test_header_is_plain_c():
    # the boundary must be consumable from C: compile a TU that only includes the header
    src = '#include "gko_b200.h"\nint main(void){return 0;}\n'
    subprocess.run(["gcc", "-std=c99", "-fsyntax-only", "-I", os.path.join(ROOT, "include"), "-x", "c", "-"],
                   input=src, text=True, check=True)


def test_library_exports_every_declared_symbol(gko):
    lib = ctypes.CDLL(gko.LIB_PATH)
    syms = declared_symbols()
    assert len(syms) > 60
    missing = [s for s in syms if not hasattr(lib, s)]
    assert not missing, missing


def test_no_cpu_fallback_in_product():
    # the product package must not import or load anything from oracle/
    pkg = os.path.join(ROOT, "repo-8852-ginkgo_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                text = open(os.path.join(dirpath, f)).read()
                assert "liboracle" not in text and "import oracle" not in text, f


def test_pick_strategy_is_pure_host(gko):
    # regular matrix -> classical row-block kernel, skewed -> merge path
    assert gko.lib.gkob200_csr_pick_strategy(8_000_000, 213_847_192, 27, 27 * 128) == 0
    assert gko.lib.gkob200_csr_pick_strategy(10_000_000, 100_000_000, 100_000, 150_000) == 1
    assert gko.lib.gkob200_csr_pick_strategy(0, 0, 0, 0) == 0


def test_merge_path_workspace_covers_carries_and_plan(gko):
    # layout (csr_spmv.cu): [carry_row: n_tiles + 1][plan: n_tiles + 2][carry_val: n_tiles + 1], one tile =
    # 128 threads x 9 merge items; the size is a pure host function, monotone in rows + entries
    import ctypes as C
    f = gko.lib.gkob200_csr_spmv_workspace_bytes
    f.restype = C.c_size_t
    tile = 128 * 9
    prev = 0
    for n_rows, nnz in [(1, 0), (1, 50), (5000, 20_000), (10_000_000, 95_363_401)]:
        n_tiles = -(-(n_rows + nnz) // tile)
        for vb in (4, 8):
            got = f(C.c_int64(n_rows), C.c_int64(nnz), C.c_int64(1), C.c_int(vb))
            assert got >= (n_tiles + 1) * 8 + (n_tiles + 2) * 8 + (n_tiles + 1) * vb
            assert got <= (n_tiles + 3) * (16 + vb) + 4096
        assert got >= prev
        prev = got
