#!/bin/bash
set -u
mkdir -p gpurun_out
echo "== dist tests"
timeout 900 python -m pytest tests/test_gpu_dist.py -x -q -m gpu > gpurun_out/r2_dist_tests_n2.log 2>&1
echo "dist tests rc=$?"; tail -3 gpurun_out/r2_dist_tests_n2.log
echo "== bench N=2 configs[3] weak"
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29712 \
  bench.py --gpus 2 --steps 5 --warmup 3 > gpurun_out/r2_bench_n2_weak.json 2> gpurun_out/r2_bench_n2_weak.err
echo "rc=$?"; grep bench_dist gpurun_out/r2_bench_n2_weak.err | cut -c1-300
