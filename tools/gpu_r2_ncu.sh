#!/bin/bash
# round-2 ncu evidence: launch list of the bench command, full captures of the dominant kernel
# (row-block CSR on configs[1]), of the current merge-path kernel (C3) and of cg_update.
set -u
mkdir -p gpurun_out
python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-config-block > gpurun_out/r2_plain_bench.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 900 --csv --log-file gpurun_out/r02_bench_launches_ncu.csv \
    python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-config-block > gpurun_out/r2_ncu_bench.log 2>&1
echo "launch list rc=$?"
python tools/run_spmv.py 27pt 200 csr 5 > gpurun_out/r2_plain_spmv.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:csr_spmv_rowblock_tma -s 3 -c 1 -o gpurun_out/r02_csr_rowblock \
    python tools/run_spmv.py 27pt 200 csr 5 > gpurun_out/r2_ncu_spmv.log 2>&1
echo "rowblock rc=$?"
python tools/run_spmv.py powerlaw 10000000 csr 5 > gpurun_out/r2_plain_mp.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:csr_spmv_merge -s 3 -c 1 -o gpurun_out/r02_csr_merge \
    python tools/run_spmv.py powerlaw 10000000 csr 5 > gpurun_out/r2_ncu_mp.log 2>&1
echo "merge rc=$?"
tail -2 gpurun_out/r2_plain_spmv.log gpurun_out/r2_plain_mp.log
