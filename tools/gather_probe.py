"""Upper bound for a CSR SpMV on the config-3 matrix: stream (col, val) and gather b[col], nothing else.
Prints the probe variants next to the shipped merge-path kernel (tools/probe/gather_probe.cu)."""
import ctypes as C, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from __graft_entry__ import load_package
gko = load_package()
_so = os.path.join(ROOT, "tools", "probe", "libgather_probe.so")
if not os.path.exists(_so):
    sys.exit("tools/probe/libgather_probe.so is missing: run `make -C tools/probe` (or __graft_entry__.build())")
lib = C.CDLL(_so)
lib.gather_probe.argtypes = [C.c_void_p, C.c_int, C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
exec_ = gko.CudaExecutor.create(0)
dev = exec_.device
NAMES = {0: "stream only 9x4", 1: "gather 9 items x 4 CTAs", 2: "gather 9x6", 3: "gather 9x8", 4: "gather 4x8",
         5: "gather 16x3", 6: "gather 9x2"}


def timed(fn, reps=20):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) * 1e3 / reps


for n in [int(a) for a in sys.argv[1:]] or [4_000_000, 10_000_000]:
    rp, ci, va = gko.gen.powerlaw_csr(n)
    nnz = len(ci)
    A = gko.matrix.Csr.from_arrays(exec_, (n, n), rp, ci, va)
    x = gko.matrix.Dense.create(exec_, (n, 1)); y = gko.matrix.Dense.create(exec_, (n, 1))
    x.t.copy_(torch.randn(n, 1, dtype=torch.float64, device=dev))
    cols = torch.from_numpy(ci).to(dev); vals = torch.from_numpy(va).to(dev)
    out = torch.empty(nnz // 4 + 4096, dtype=torch.float64, device=dev)
    stream = torch.cuda.current_stream().cuda_stream
    alg = A.spmv_bytes(1)
    us = timed(lambda: A.apply(x, y))
    print(f"n={n} nnz={nnz}: shipped merge-path {us:.1f} us = {alg/us/1e3:.0f} GB/s", flush=True)
    for mode, name in NAMES.items():
        def run():
            rc = lib.gather_probe(stream, mode, nnz, cols.data_ptr(), vals.data_ptr(), x.t.data_ptr(), out.data_ptr())
            assert rc == 0
        us = timed(run)
        print(f"  probe {name:26s} {us:8.1f} us  {us*1e6/nnz:5.2f} ps/entry  (algorithmic bytes of the SpMV / t = {alg/us/1e3:.0f} GB/s)", flush=True)
    del A, x, y, cols, vals, out
