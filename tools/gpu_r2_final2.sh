#!/bin/bash
# final single-GPU validation after the merge-path / SpMM rewrites: whole GPU suite, smoke, both bench arms, SpMM timing
set -u
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -x -q -m gpu > gpurun_out/r2_final2_gputests.log 2>&1; echo "gpu tests rc=$?"; tail -3 gpurun_out/r2_final2_gputests.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
SECONDS=0; python bench.py > gpurun_out/r2_final2_bench_n1.json 2> gpurun_out/r2_final2_bench_n1.err; echo "bench rc=$? wall=${SECONDS}s"
SECONDS=0; python bench.py --impl reference > gpurun_out/r2_final2_ref_n1.json 2> gpurun_out/r2_final2_ref_n1.err; echo "ref rc=$? wall=${SECONDS}s"
python tools/spmm_time.py 256 fp32 32 csr,ell,sellp
python tools/spmm_time.py 200 fp64 32 csr
python tools/gather_probe.py 10000000 2>&1 | tail -9
