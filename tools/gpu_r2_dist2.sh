#!/bin/bash
# 2-GPU validation of the distributed path: parity tests on all three exchange paths, then
# bench_dist (configs[3] weak at N=2) with its in-run parity checks.
set -u
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name,clocks.max.sm --format=csv > gpurun_out/r2_gpus.txt 2>&1
echo "== dist tests" 
timeout 900 python -m pytest tests/test_gpu_dist.py -x -q -m gpu > gpurun_out/r2_dist_tests_n2.log 2>&1
echo "dist tests rc=$?"; tail -5 gpurun_out/r2_dist_tests_n2.log
echo "== bench N=2 small grid"
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29711 \
  bench.py --gpus 2 --grid 128 --steps 2 --warmup 3 > gpurun_out/r2_bench_n2_g128.json 2> gpurun_out/r2_bench_n2_g128.err
echo "rc=$?"; tail -3 gpurun_out/r2_bench_n2_g128.err; cat gpurun_out/r2_bench_n2_g128.json
echo "== bench N=2 configs[3] weak"
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29712 \
  bench.py --gpus 2 --steps 5 --warmup 3 > gpurun_out/r2_bench_n2_weak.json 2> gpurun_out/r2_bench_n2_weak.err
echo "rc=$?"; tail -3 gpurun_out/r2_bench_n2_weak.err; cat gpurun_out/r2_bench_n2_weak.json
echo "== bench N=2 weak, fallback path (GKOB200_FUSED_HALO=0)"
GKOB200_FUSED_HALO=0 timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29713 \
  bench.py --gpus 2 --steps 5 --warmup 3 --no-verify > gpurun_out/r2_bench_n2_weak_nofuse.json 2> gpurun_out/r2_bench_n2_weak_nofuse.err
echo "rc=$?"; cat gpurun_out/r2_bench_n2_weak_nofuse.json
echo "== bench N=1"
timeout 600 python bench.py --steps 5 --warmup 3 > gpurun_out/r2_bench_n1.json 2> gpurun_out/r2_bench_n1.err
echo "rc=$?"; cat gpurun_out/r2_bench_n1.json
