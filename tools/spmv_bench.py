#!/usr/bin/env python
"""Native SpMV benchmark driver with the JSON schema of the reference's benchmark/spmv/spmv.cpp
(:64-290): reads a JSON array of test cases from stdin, each `{"filename": "<MatrixMarket or
GINKGO-binary file>"}` (this driver also accepts `{"stencil": "27pt|7pt|5pt", "size": N}`), and
for every requested format adds

    "spmv": { "<format>": {"storage": bytes, "max_relative_norm2": e (with --detailed),
                           "time": seconds, "repetitions": n, "completed": true} },
    "optimal": {"spmv": "<fastest format>"}

and prints the array to stdout, so the output can be compared with ginkgo-data results.
Extra keys (not in the reference): "bandwidth_gbs" (algorithmic bytes / time), "kernel".

  echo '[{"stencil": "27pt", "size": 100}]' | python tools/spmv_bench.py --formats csr,ell,sellp,hybrid,coo --detailed
"""
import argparse
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np  # noqa: E402
import torch  # noqa: E402

from __graft_entry__ import load_package  # noqa: E402


def storage_bytes(M):
    total = 0
    for name in ("row_ptrs", "col_idxs", "values", "row_idxs", "slice_sets", "slice_lengths"):
        t = getattr(M, name, None)
        if isinstance(t, torch.Tensor):
            total += t.numel() * t.element_size()
    for part in ("ell", "coo"):
        if hasattr(M, part):
            total += storage_bytes(getattr(M, part))
    return total


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--formats", default="csr,ell,sellp,hybrid,coo")
    ap.add_argument("--nrhs", type=int, default=1)
    ap.add_argument("--detailed", action="store_true")
    ap.add_argument("--repetitions", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=2)
    ap.add_argument("--fp32", action="store_true")
    args = ap.parse_args()
    gko = load_package()
    exec_ = gko.CudaExecutor.create(0)
    dtype = np.float32 if args.fp32 else np.float64
    tdt = torch.float32 if args.fp32 else torch.float64
    cases = json.load(sys.stdin)
    if not isinstance(cases, list):
        sys.exit("input has to be a JSON array of test cases")
    for case in cases:
        try:
            if "filename" in case:
                A = gko.io.read(exec_, case["filename"], value_dtype=dtype)
            elif "stencil" in case:
                g = int(case["size"])
                rp, ci, va, n = gko.gen.stencil_csr(case["stencil"], g, g, g, value_dtype=dtype)
                A = gko.matrix.Csr.from_arrays(exec_, (n, n), rp, ci, va)
            else:
                raise ValueError('test case needs "filename" (or "stencil" + "size")')
            case.setdefault("spmv", {})
            case.setdefault("optimal", {})
            n, m = A.size
            print(f"Matrix is of size ({n}, {m})", file=sys.stderr)
            rng = torch.Generator(device=exec_.device).manual_seed(42)
            b = gko.matrix.Dense.create(exec_, (m, args.nrhs), tdt)
            b.t.copy_(torch.rand(m, args.nrhs, dtype=tdt, device=exec_.device, generator=rng) * 2 - 1)
            x = gko.matrix.Dense.create(exec_, (n, args.nrhs), tdt)
            answer = None
            if args.detailed:
                answer = gko.matrix.Dense.create(exec_, (n, args.nrhs), tdt)
                A.convert_to("coo").apply(b, answer)
            best = None
            for fmt in args.formats.split(","):
                res = case["spmv"].setdefault(fmt, {})
                try:
                    M = A if fmt == "csr" else A.convert_to(fmt)
                    res["storage"] = storage_bytes(M)
                    if args.detailed:
                        M.apply(b, x)
                        num = torch.linalg.vector_norm(x.t - answer.t, dim=0)
                        den = torch.linalg.vector_norm(answer.t, dim=0)
                        res["max_relative_norm2"] = float((num / den).max().item())
                    for _ in range(args.warmup):
                        M.apply(b, x)
                    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                    torch.cuda.synchronize()
                    e0.record()
                    for _ in range(args.repetitions):
                        M.apply(b, x)
                    e1.record()
                    torch.cuda.synchronize()
                    t = e0.elapsed_time(e1) * 1e-3 / args.repetitions
                    res.update({"time": t, "repetitions": args.repetitions, "completed": True,
                                "bandwidth_gbs": M.spmv_bytes(args.nrhs) / t / 1e9, "kernel": M.kernel()})
                    if best is None or t < best[1]:
                        best = (fmt, t)
                        case["optimal"]["spmv"] = fmt
                except Exception as e:  # noqa: BLE001  (the reference records the failure and goes on)
                    res["completed"] = False
                    res["error"] = str(e)
                    print(f"Error when processing test case {case.get('filename', case)}: {e}", file=sys.stderr)
        except Exception as e:  # noqa: BLE001
            print(f"Error setting up matrix data, what(): {e}", file=sys.stderr)
    json.dump(cases, sys.stdout, indent=4)
    print()


if __name__ == "__main__":
    main()
