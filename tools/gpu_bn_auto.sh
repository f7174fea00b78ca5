#!/bin/bash
OUT=gpurun_out/${1:-bnauto}; mkdir -p $OUT
export PYTHONUNBUFFERED=1
timeout 900 python -m pytest tests/test_kernels_gpu.py tests/test_model_gpu.py -m gpu -q --tb=short -x -k "gemm or golden" > $OUT/pytest.log 2>&1; echo "tests rc=$?"; tail -3 $OUT/pytest.log
for bn in 0 2 0 2; do echo "RBA_TC_BN256=$bn"; RBA_TC_BN256=$bn python bench.py --no-cpu-baseline --steps 10 --warmup 3 2>/dev/null | python -c "
import sys, json
r = json.loads(sys.stdin.read()); print('  value %.2f img/s ms/step %.2f clocks %s' % (r['value'], r['ms_per_step'], r['clocks']['sm_mhz']))"; done | tee $OUT/bn_auto.txt
