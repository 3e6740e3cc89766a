timeout 900 python -m pytest tests/test_gpu_csr.py tests/test_gpu_formats.py -x -q -m gpu 2>&1 | tail -2
for v in 1 0; do echo "SPMM_VECTOR=$v fp32"; echo '[{"stencil": "27pt", "size": 256}]' | GKOB200_SPMM_VECTOR=$v python tools/spmv_bench.py --formats csr,ell --nrhs 32 --fp32 2>/dev/null | python -c "
import json,sys
for c in json.load(sys.stdin):
    for f,r in c['spmv'].items(): print(f, round(r['time']*1e3,3),'ms', round(r.get('bandwidth_gbs',0)),'GB/s')
"; done
for v in 1 0; do echo "SPMM_VECTOR=$v fp64 200^3"; echo '[{"stencil": "27pt", "size": 200}]' | GKOB200_SPMM_VECTOR=$v python tools/spmv_bench.py --formats csr,sellp --nrhs 32 2>/dev/null | python -c "
import json,sys
for c in json.load(sys.stdin):
    for f,r in c['spmv'].items(): print(f, round(r['time']*1e3,3),'ms', round(r.get('bandwidth_gbs',0)),'GB/s')
"; done
