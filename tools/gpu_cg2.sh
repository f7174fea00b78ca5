#!/bin/bash
OUT=gpurun_out/cg2; mkdir -p $OUT
export PYTHONUNBUFFERED=1
timeout 180 python -m pytest tests/test_kernels_gpu.py -m gpu -q -x -k "gemm_tc" 2>&1 | tail -8 | tee $OUT/pytest_gemm.txt
grep -q "passed" $OUT/pytest_gemm.txt && ! grep -q "failed" $OUT/pytest_gemm.txt || { echo "gemm tests failed: stop"; exit 1; }
for cg in 1 0; do
  echo "== RBA_TC_CG2=$cg"
  RBA_TC_CG2=$cg timeout 300 python tools/bench_gemm.py 2>&1 | tail -10 | tee $OUT/bench_gemm_cg$cg.txt
done
