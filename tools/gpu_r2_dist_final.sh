#!/bin/bash
# final multi-GPU check of round 2 with the shipped kernels: configs[3] weak (+ strong) at $1 GPUs, distributed tests
set -u
N=${1:-2}
mkdir -p gpurun_out
run_bench() {  # n name extra...
  local n=$1 name=$2; shift 2
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((29800 + RANDOM % 100)) \
    bench.py --gpus $n --steps 5 --warmup 3 "$@" > gpurun_out/$name.json 2> gpurun_out/$name.err
  echo "$name rc=$?"; grep "bench_dist\]" gpurun_out/$name.err | cut -c1-260; cut -c1-200 gpurun_out/$name.json
}
run_bench $N r2f_weak_n$N
run_bench $N r2f_strong_n$N --scaling strong --no-verify
if [ "$N" = 2 ]; then
  timeout 600 python -m pytest tests/test_gpu_dist.py -q -m gpu > gpurun_out/r2f_dist_tests_n$N.log 2>&1; echo "dist tests rc=$?"; tail -3 gpurun_out/r2f_dist_tests_n$N.log
  timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29790 bench.py --gpus $N --impl reference --steps 2 --warmup 1 > gpurun_out/r2f_ref_n$N.json 2> gpurun_out/r2f_ref_n$N.err; echo "ref rc=$?"; cut -c1-300 gpurun_out/r2f_ref_n$N.json
fi
