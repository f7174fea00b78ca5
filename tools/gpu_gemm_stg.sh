#!/bin/bash
OUT=gpurun_out/${1:-gemmstg}; mkdir -p $OUT
export PYTHONUNBUFFERED=1
for k in 512 256 128; do echo "RBA_TC_STG_MAXK=$k"; RBA_TC_STG_MAXK=$k python tools/bench_gemm.py 2>/dev/null | python -c "
import sys, json
for l in sys.stdin:
    r = json.loads(l); print('  %-12s M=%-8d N=%-6d K=%-5d tc %.3f ms %.0f TF/s %.0f GB/s' % (r['name'], r['M'], r['N'], r['K'], r.get('tc_ms', -1), r.get('tc_tflops', -1), r.get('tc_GBs', -1)))
"; done | tee $OUT/gemm_stg.txt
