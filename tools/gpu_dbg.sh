#!/bin/bash
OUT=gpurun_out/${1:-d1}; mkdir -p $OUT
for d in 0 1 2; do echo "== RBA_TC_DEBUG=$d"; RBA_TC_DEBUG=$d RBA_TC_BN256=0 timeout 300 python tools/bench_gemm.py 2>/dev/null | python -c "
import json,sys
for l in sys.stdin:
    r=json.loads(l); print(r['name'], 'tc %.3f ms %.0f TF'%(r.get('tc_ms',0), r.get('tc_tflops',0)))
"; done
