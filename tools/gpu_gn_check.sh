#!/bin/bash
OUT=gpurun_out/${1:-gn}; mkdir -p $OUT
export PYTHONUNBUFFERED=1
timeout 600 python -m pytest tests/test_kernels_gpu.py tests/test_model_gpu.py -m gpu -q --tb=short -x -k "groupnorm or golden or invariance" > $OUT/pytest.log 2>&1; echo "tests rc=$?"; tail -3 $OUT/pytest.log
python tools/profile_forward.py --batch 8 > $OUT/profile_b8.txt 2>&1; grep -E "total kernel|gn_" $OUT/profile_b8.txt
