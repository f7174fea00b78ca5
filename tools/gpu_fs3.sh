#!/bin/bash
OUT=gpurun_out/${1:-fs3}; mkdir -p $OUT
export PYTHONUNBUFFERED=1
timeout 300 python -m pytest tests/test_kernels_gpu.py -m gpu -q --tb=short -x -k "einsum_score" > $OUT/pytest_fs.log 2>&1; echo "fs kernel rc=$?"; tail -3 $OUT/pytest_fs.log
python tools/fused_score_only.py 8 10 2>&1 | tail -1
