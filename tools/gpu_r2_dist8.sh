#!/bin/bash
# 8-GPU run: multi-rank parity tests on all exchange paths (log kept under profiles/), A/B
# breakdown, configs[3] weak N=8/4/2 and strong N=8/4/2.
set -u
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name,clocks.max.sm --format=csv > gpurun_out/r2_gpus8.txt 2>&1
nvidia-smi topo -m >> gpurun_out/r2_gpus8.txt 2>&1
run_bench() {  # n name extra...
  local n=$1 name=$2; shift 2
  timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((29800 + RANDOM % 100)) \
    bench.py --gpus $n --steps 5 --warmup 3 "$@" > gpurun_out/$name.json 2> gpurun_out/$name.err
  echo "$name rc=$?"; grep "bench_dist\]" gpurun_out/$name.err | cut -c1-260; cut -c1-200 gpurun_out/$name.json
}
echo "== weak N=8 (with --verify-full)"; run_bench 8 r2_weak_n8 --verify-full
echo "== dist tests world 8 (3 paths) + world 4 fused"
timeout 900 python -m pytest tests/test_gpu_dist.py -q -m gpu -k "torchrun and (8- or 4-fused)" > gpurun_out/r2_dist_tests_n8.log 2>&1
echo "dist tests rc=$?"; tail -4 gpurun_out/r2_dist_tests_n8.log
echo "== A/B N=8"
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29791 tools/dist_ab.py 2>&1 | grep "^AB\|rror" | tee gpurun_out/r2_dist_ab_n8.log
echo "== weak N=4, N=2"; run_bench 4 r2_weak_n4; run_bench 2 r2_weak_n2
echo "== strong N=4, N=2"; run_bench 4 r2_strong_n4 --scaling strong --no-verify; run_bench 2 r2_strong_n2 --scaling strong --no-verify
echo "== weak N=8 fallback paths"
GKOB200_FUSED_HALO=0 run_bench 8 r2_weak_n8_nofuse --no-verify
