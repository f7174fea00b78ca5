#!/bin/bash
OUT=gpurun_out/${1:-wt}; mkdir -p $OUT
export PYTHONUNBUFFERED=1
timeout 300 python -m pytest tests/test_kernels_gpu.py -q -x -k "window_attention_tensor_core" 2>&1 | tail -15 | tee $OUT/pytest_wattn.txt
timeout 300 python tools/bench_wattn.py 8 2>&1 | tail -8 | tee $OUT/bench_wattn.txt
