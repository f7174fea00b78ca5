#!/bin/bash
OUT=gpurun_out/${1:-wt5}; mkdir -p $OUT
export PYTHONUNBUFFERED=1
timeout 300 python -m pytest tests/test_kernels_gpu.py -q -x -k "window_attention_tensor_core" 2>&1 | tail -5 | tee $OUT/pytest_wattn.txt
{
for d in 0 3 4 7; do echo "RBA_WT_DEBUG=$d"; RBA_WT_DEBUG=$d python tools/bench_wattn_one.py 2 8 10 2>&1 | tail -1; done
echo "=== timeline full ==="; RBA_WT_TIMELINE=1 python tools/bench_wattn_one.py 2 8 1 2>&1 | tail -16
} | tee $OUT/ablation.txt
timeout 300 python tools/bench_wattn.py 8 2>&1 | tail -6 | tee $OUT/bench_wattn.txt
