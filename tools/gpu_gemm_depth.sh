#!/bin/bash
# Does the operand ring depth (2 stages in the store-staged config vs 3) bound the K=512 shapes?  Stores skipped (DEBUG=1) or
# whole epilogue skipped (DEBUG=2) so that only the TMA + MMA pipeline is compared.
OUT=gpurun_out/${1:-gemmdepth}; mkdir -p $OUT
export PYTHONUNBUFFERED=1
for dbg in 0 1 2; do for k in 512 256; do echo "RBA_TC_DEBUG=$dbg RBA_TC_STG_MAXK=$k ($( [ $k = 512 ] && echo '2-stage staged' || echo '3-stage direct' ) for K=512)"; RBA_TC_DEBUG=$dbg RBA_TC_STG_MAXK=$k python tools/bench_gemm.py 2>/dev/null | python -c "
import sys, json
for l in sys.stdin:
    r = json.loads(l)
    if r['name'] in ('s2_qkv','s2_fc1','s0_fc2'): print('  %-12s tc %.3f ms %.0f TF/s' % (r['name'], r.get('tc_ms', -1), r.get('tc_tflops', -1)))
"; done; done | tee $OUT/gemm_depth.txt
