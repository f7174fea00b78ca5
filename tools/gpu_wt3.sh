#!/bin/bash
OUT=gpurun_out/${1:-wt3}; mkdir -p $OUT
export PYTHONUNBUFFERED=1
timeout 300 python -m pytest tests/test_kernels_gpu.py -q -x -k "window_attention_tensor_core" 2>&1 | tail -5 | tee $OUT/pytest_wattn.txt
{
for d in 0 3 4 7; do echo "RBA_WT_DEBUG=$d"; RBA_WT_DEBUG=$d python tools/bench_wattn_one.py 2 8 10 2>&1 | tail -1; done
for st in 0 1 3; do python tools/bench_wattn_one.py $st 8 10 2>&1 | tail -1; done
} | tee $OUT/ablation.txt
