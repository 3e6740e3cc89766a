#!/bin/bash
set -u
export LD_LIBRARY_PATH=shim/_build:oracle/_ref/lib:repo-8852-ginkgo_b200
timeout 600 shim/_build/test_dropin > gpurun_out/r2_dropin.log 2>&1; echo "dropin rc=$?"; grep -c " ok$" gpurun_out/r2_dropin.log; grep -v " ok$" gpurun_out/r2_dropin.log | head -20
timeout 300 python -m pytest tests/test_gpu_krylov.py tests/test_gpu_dropin.py -x -q -m gpu 2>&1 | tail -3
python tools/run_spmv.py powerlaw 10000000 csr 5 > gpurun_out/plain_merge.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k 'regex:csr_spmv_merge$' -s 3 -c 1 -o gpurun_out/r02_merge \
    python tools/run_spmv.py powerlaw 10000000 csr 5 > gpurun_out/ncu_merge.log 2>&1
echo "merge capture rc=$?"; tail -2 gpurun_out/ncu_merge.log
