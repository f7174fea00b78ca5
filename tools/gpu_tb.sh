#!/bin/bash
# tests + bench
OUT=gpurun_out/${1:-tb}; mkdir -p $OUT
export PYTHONUNBUFFERED=1
timeout 1200 python -m pytest tests -m gpu -q --tb=short -x > $OUT/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -12 $OUT/pytest_gpu.log
timeout 900 python bench.py --no-cpu-baseline > $OUT/bench.json 2> $OUT/bench.err; echo "bench rc=$?"; cat $OUT/bench.json; tail -3 $OUT/bench.err
