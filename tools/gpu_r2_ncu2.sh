#!/bin/bash
# round-2 ncu evidence, part 2: current merge-path, SELL-P, ELL, COO kernels
set -u
mkdir -p gpurun_out
cap() {  # name regex args...
  local name=$1 rx=$2; shift 2
  python tools/run_spmv.py "$@" > gpurun_out/plain_$name.log 2>&1 &&
  ncu --set full --clock-control none --import-source on -k "regex:$rx" -s 3 -c 1 -o gpurun_out/r02_$name \
      python tools/run_spmv.py "$@" > gpurun_out/ncu_$name.log 2>&1
  echo "$name rc=$? $(tail -1 gpurun_out/plain_$name.log)"
}
cap merge 'csr_spmv_merge<' powerlaw 10000000 csr 5
cap sellp 'sellp_spmv_tma' 27pt 200 sellp 5
cap ell 'ell_spmv_tma' 27pt 200 ell 5
cap coo 'coo_spmv2' 27pt 200 coo 5
