#!/bin/bash
OUT=gpurun_out/${1:-wt6}; mkdir -p $OUT
export PYTHONUNBUFFERED=1
{
for d in 7 15 23 31 8 16 24; do echo "RBA_WT_DEBUG=$d"; RBA_WT_DEBUG=$d python tools/bench_wattn_one.py 2 8 10 2>&1 | tail -1; done
} | tee $OUT/ablation.txt
