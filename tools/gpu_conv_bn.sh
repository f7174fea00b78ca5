#!/bin/bash
OUT=gpurun_out/${1:-convbn}; mkdir -p $OUT
export PYTHONUNBUFFERED=1
timeout 900 python -m pytest tests/test_kernels_gpu.py tests/test_model_gpu.py -m gpu -q --tb=short -x -k "conv or golden" > $OUT/pytest.log 2>&1; echo "tests rc=$?"; tail -3 $OUT/pytest.log
for bn in 0 2; do echo "RBA_TC_BN256=$bn"; RBA_TC_BN256=$bn python tools/profile_forward.py --batch 8 2>&1 | grep -E "total kernel|true, 0, false, false"; done | tee $OUT/conv_bn.txt
