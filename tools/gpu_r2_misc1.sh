#!/bin/bash
set -u
mkdir -p gpurun_out
python - <<'PY' > gpurun_out/r2_l2_attrs.txt 2>&1
import ctypes as C
rt = C.CDLL("libcudart.so.12")
for name, a in (("MaxPersistingL2CacheSize", 108), ("MaxAccessPolicyWindowSize", 109), ("L2CacheSize", 38), ("MultiProcessorCount", 16)):
    v = C.c_int(0); rc = rt.cudaDeviceGetAttribute(C.byref(v), a, 0); print(name, v.value, "rc", rc)
PY
cat gpurun_out/r2_l2_attrs.txt
nproc; free -g | head -2
echo "== reference arm N=8 (CPU only)"
OMP_NUM_THREADS=1 RANK=0 WORLD_SIZE=8 timeout 900 python bench.py --impl reference --gpus 8 --steps 3 --warmup 1 > gpurun_out/r2_ref_n8.json 2> gpurun_out/r2_ref_n8.err; echo "rc=$?"; cut -c1-400 gpurun_out/r2_ref_n8.json; tail -2 gpurun_out/r2_ref_n8.err
echo "== reference arm N=1"
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r2_ref_n1.json 2> gpurun_out/r2_ref_n1.err; echo "rc=$?"; cut -c1-300 gpurun_out/r2_ref_n1.json
echo "== strong N=1 (512^3 on one GPU through the distributed classes)"
timeout 900 python bench.py --gpus 1 --dist --scaling strong --steps 3 --warmup 3 --no-verify > gpurun_out/r2_strong_n1.json 2> gpurun_out/r2_strong_n1.err; echo "rc=$?"; cut -c1-300 gpurun_out/r2_strong_n1.json; tail -3 gpurun_out/r2_strong_n1.err
