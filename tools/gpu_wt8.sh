#!/bin/bash
OUT=gpurun_out/${1:-wt8}; mkdir -p $OUT
export PYTHONUNBUFFERED=1
for d in 96 32; do
echo "== M=64 / lane-offset experiment (RBA_WT_DEBUG=$d) =="
RBA_WT_DEBUG=$d timeout 300 python -m pytest tests/test_kernels_gpu.py -q -k "window_attention_tensor_core and tcgen05" 2>&1 | grep -E "passed|failed|FAILED" | tee -a $OUT/pytest_m64.txt
done
