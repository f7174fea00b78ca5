#!/bin/bash
OUT=gpurun_out/${1:-t1}; mkdir -p $OUT
export PYTHONUNBUFFERED=1
timeout 600 python -m pytest tests -m gpu -q --tb=short -x > $OUT/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -15 $OUT/pytest_gpu.log
