#!/bin/bash
OUT=gpurun_out/${1:-attn}; mkdir -p $OUT
export PYTHONUNBUFFERED=1
timeout 600 python -m pytest tests/test_kernels_gpu.py tests/test_model_gpu.py -m gpu -q --tb=short -x -k "attn or golden or layernorm" > $OUT/pytest.log 2>&1; echo "tests rc=$?"; tail -3 $OUT/pytest.log
python tools/profile_forward.py --batch 8 > $OUT/profile_b8.txt 2>&1; sed -n 3,9p $OUT/profile_b8.txt
