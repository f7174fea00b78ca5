#!/bin/bash
OUT=gpurun_out/${1:-gemmwait}; mkdir -p $OUT
export PYTHONUNBUFFERED=1
for rep in 1 2; do for d in 0 4; do echo "RBA_TC_DEBUG=$d (rep $rep)"; RBA_TC_DEBUG=$d python tools/bench_gemm.py 2>/dev/null | python -c "
import sys, json
for l in sys.stdin:
    r = json.loads(l); print('  %-12s tc %.3f ms %.0f TF/s' % (r['name'], r.get('tc_ms', -1), r.get('tc_tflops', -1)))
" | grep -E "s2_|s3_|s0_fc1"; done; done | tee $OUT/gemm_wait.txt
for d in 0 4; do echo "bench RBA_TC_DEBUG=$d"; RBA_TC_DEBUG=$d python bench.py --no-cpu-baseline --steps 10 --warmup 3 2>/dev/null | python -c "
import sys, json
r = json.loads(sys.stdin.read()); print('  value %.2f img/s ms/step %.2f clocks %s' % (r['value'], r['ms_per_step'], r['clocks']))"; done | tee -a $OUT/gemm_wait.txt
