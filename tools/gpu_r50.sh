#!/bin/bash
OUT=gpurun_out/${1:-r50}; mkdir -p $OUT
timeout 600 python -m pytest tests -m gpu -q -k "r50" -s 2>&1 | grep -E "^r50|passed|failed|FAILED" | tee $OUT/pytest_r50.txt
