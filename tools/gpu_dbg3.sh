#!/bin/bash
OUT=gpurun_out/${1:-d4}; mkdir -p $OUT
for d in 0 2; do echo "== RBA_TC_DEBUG=$d"; RBA_TC_DEBUG=$d timeout 300 python tools/bench_gemm.py 2>/dev/null | python -c "
import json,sys
for l in sys.stdin:
    r=json.loads(l); print(r['name'], 'tc %.3f ms %.0f TF %.0f GB/s'%(r.get('tc_ms',0), r.get('tc_tflops',0), r.get('tc_GBs',0)))
"; done
RBA_PROFILE_SEQ=1 timeout 600 python tools/profile_forward.py --batch 8 > $OUT/profile_seq_b8.txt 2>&1
