#!/bin/bash
OUT=gpurun_out/${1:-wt11}; mkdir -p $OUT
export PYTHONUNBUFFERED=1
timeout 300 python -m pytest tests/test_kernels_gpu.py -q -k "window_attention_tensor_core or tiled_qkv" 2>&1 | grep -E "passed|failed|FAILED|Error" | tee $OUT/pytest_wattn.txt
{
for d in 0 3 4 7; do echo "RBA_WT_DEBUG=$d"; RBA_WT_DEBUG=$d python tools/bench_wattn_one.py 2 8 10 2>&1 | tail -1; done
} | tee $OUT/ablation.txt
timeout 300 python tools/bench_wattn.py 8 2>&1 | tail -1 | tee $OUT/bench_wattn.txt
