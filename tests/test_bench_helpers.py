"""Host-side pieces of the benchmark harness (no GPU): the reference arm must not map the product
library and must use every host thread even under torchrun's OMP_NUM_THREADS=1; the numpy
restatement bench_dist.py checks the full-size distributed apply against must equal the oracle's
distributed apply bit for bit."""
import json
import os
import subprocess
import sys

import numpy as np
import pytest

import oracle

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.parametrize("extra", [[], ["--gpus", "2"], ["--gpus", "4", "--scaling", "strong"]])
def test_reference_arm_runs_without_the_product_library(extra):
    code = ("import sys, json; sys.argv = ['bench.py', '--impl', 'reference', '--grid', '16', '--steps', '1', "
            f"'--warmup', '0'] + {extra!r}; import bench; bench.main(); "
            "maps = open('/proc/self/maps').read(); assert 'libgko_b200' not in maps, 'product library mapped'")
    env = dict(os.environ, OMP_NUM_THREADS="1", RANK="0", WORLD_SIZE=str(extra[1]) if extra else "1")
    out = subprocess.run([sys.executable, "-c", code], cwd=ROOT, capture_output=True, text=True, env=env, timeout=300)
    assert out.returncode == 0, out.stderr[-2000:]
    line = json.loads(out.stdout.strip().splitlines()[-1])
    assert line["impl"] == "reference" and line["value"] > 0 and line["gpu_launches"] == 0
    if oracle.ref() is not None:
        assert line["cpu_baseline"]["kind"] == "reference"
        assert line["cpu_baseline"]["cores"] == len(os.sched_getaffinity(0))   # not torchrun's 1


def test_reference_arm_other_ranks_do_nothing():
    env = dict(os.environ, RANK="1", WORLD_SIZE="2")
    out = subprocess.run([sys.executable, "bench.py", "--impl", "reference", "--gpus", "2", "--grid", "16"], cwd=ROOT,
                         capture_output=True, text=True, env=env, timeout=120)
    assert out.returncode == 0 and out.stdout.strip() == ""


@pytest.mark.parametrize("kind,dims,parts", [("7pt", (5, 4, 6), 3), ("7pt", (6, 5, 8), 8), ("27pt", (5, 6, 4), 2),
                                             ("27pt", (4, 4, 7), 7), ("7pt", (3, 3, 2), 1)])
def test_expected_rows_is_the_distributed_apply(kind, dims, parts):
    sys.path.insert(0, ROOT)
    import bench_dist
    nx, ny, nz = dims
    rp, ci, va, n = oracle.gen_stencil_csr(kind, nx, ny, nz)
    rows = np.repeat(np.arange(n, dtype=np.int64), np.diff(rp))
    part = oracle.Partition.uniform(parts, n)
    allp = [oracle.dist_build_local_nonlocal(rows, ci.astype(np.int64), va, part, p) for p in range(parts)]
    x = bench_dist.test_vector(np.arange(n, dtype=np.int64))
    want = oracle.dist_apply(allp, part, x[:, None])[:, 0]
    for p in range(parts):
        lo, hi = int(part.bounds[p]), int(part.bounds[p + 1])
        got = bench_dist.expected_rows(kind, nx, ny, nz, lo, hi, chunk=37)
        assert np.array_equal(got, want[lo:hi]), p
