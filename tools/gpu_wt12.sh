#!/bin/bash
OUT=gpurun_out/${1:-wt12}; mkdir -p $OUT
export PYTHONUNBUFFERED=1
{
echo "=== debug 7 timeline (tiled) ==="; RBA_WT_DEBUG=7 RBA_WT_TIMELINE=1 python tools/bench_wattn_one.py 2 8 1 2>&1 | tail -15
echo "=== debug 31 timeline (no MMAs at all) ==="; RBA_WT_DEBUG=31 RBA_WT_TIMELINE=1 python tools/bench_wattn_one.py 2 8 1 2>&1 | tail -15
} | tee $OUT/timeline.txt
