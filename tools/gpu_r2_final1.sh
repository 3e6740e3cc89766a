#!/bin/bash
# final single-GPU validation of round 2: whole GPU suite, smoke, bench, ncu of the new ELL / COO kernels
set -u
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -x -q -m gpu > gpurun_out/r2_final_gputests.log 2>&1; echo "gpu tests rc=$?"; tail -3 gpurun_out/r2_final_gputests.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
SECONDS=0; python bench.py > gpurun_out/r2_final_bench_n1.json 2> gpurun_out/r2_final_bench_n1.err; echo "bench rc=$? wall=${SECONDS}s"
SECONDS=0; python bench.py --impl reference > gpurun_out/r2_final_ref_n1.json 2> gpurun_out/r2_final_ref_n1.err; echo "ref rc=$? wall=${SECONDS}s"
cap() { local name=$1 rx=$2; shift 2
  python tools/run_spmv.py "$@" > gpurun_out/plain_$name.log 2>&1 &&
  ncu --set full --clock-control none --import-source on -k "regex:$rx" -s 3 -c 1 -o gpurun_out/r02_$name python tools/run_spmv.py "$@" > gpurun_out/ncu_$name.log 2>&1
  echo "$name rc=$? $(tail -1 gpurun_out/plain_$name.log)"; }
cap ell2d 'ell_spmv_tma2d' 27pt 200 ell 5
cap coobulk 'coo_spmv2' 27pt 200 coo 5
cap sellp14 'sellp_spmv_tma' 27pt 200 sellp 5
