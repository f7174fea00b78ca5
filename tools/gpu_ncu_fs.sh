#!/bin/bash
# ncu --set full of the fused score kernel (B=8), default configuration
OUT=gpurun_out/${1:-ncufs}; mkdir -p $OUT
ncu --set full --clock-control none --import-source on -k regex:rba_einsum_score -s 3 -c 1 -o $OUT/fused_score -f python tools/fused_score_only.py 8 3 > $OUT/ncu.log 2>&1
tail -3 $OUT/ncu.log
ls -la $OUT
