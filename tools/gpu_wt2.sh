#!/bin/bash
OUT=gpurun_out/${1:-wt2}; mkdir -p $OUT
export PYTHONUNBUFFERED=1
{
for d in 0 1 2 3 4 5 6 7; do echo "RBA_WT_DEBUG=$d"; RBA_WT_DEBUG=$d python tools/bench_wattn_one.py 2 8 10 2>&1 | tail -1; done
} | tee $OUT/ablation.txt
