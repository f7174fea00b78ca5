#!/bin/bash
OUT=gpurun_out/${1:-fsc}; mkdir -p $OUT
export PYTHONUNBUFFERED=1
timeout 600 python -m pytest tests/test_kernels_gpu.py tests/test_model_gpu.py -m gpu -q --tb=short -x -k "einsum_score or fused or golden or energy or densehybrid or full_size" > $OUT/pytest_fs.log 2>&1; echo "fs tests rc=$?"; tail -3 $OUT/pytest_fs.log
python tools/fused_score_only.py 8 20 2>&1 | tail -1 | tee $OUT/time.txt
