#!/bin/bash
OUT=gpurun_out/${1:-d3}; mkdir -p $OUT
for k in 256 512 100000; do echo "== RBA_TC_STG_MAXK=$k"; RBA_TC_STG_MAXK=$k timeout 300 python tools/bench_gemm.py 2>/dev/null | python -c "
import json,sys
for l in sys.stdin:
    r=json.loads(l); print(r['name'], 'tc %.3f ms %.0f TF'%(r.get('tc_ms',0), r.get('tc_tflops',0)))
"; RBA_TC_STG_MAXK=$k timeout 600 python bench.py --no-cpu-baseline --steps 5 2>/dev/null | python -c "
import json,sys
r=json.loads(sys.stdin.read()); print('bench %.2f img/s'%r['value'])
"; done
