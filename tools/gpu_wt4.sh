#!/bin/bash
OUT=gpurun_out/${1:-wt4}; mkdir -p $OUT
export PYTHONUNBUFFERED=1
{
echo "=== full ==="; RBA_WT_TIMELINE=1 python tools/bench_wattn_one.py 2 8 1 2>&1 | tail -16
echo "=== debug 7 (no math, no tile 2) ==="; RBA_WT_DEBUG=7 RBA_WT_TIMELINE=1 python tools/bench_wattn_one.py 2 8 1 2>&1 | tail -16
} | tee $OUT/timeline.txt
