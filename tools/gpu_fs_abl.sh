#!/bin/bash
# Ablation timings of the fused kernel (B=8): which phase bounds it
OUT=gpurun_out/${1:-fsabl}; mkdir -p $OUT
export PYTHONUNBUFFERED=1
{
echo "default";            python tools/fused_score_only.py 8 20 2>&1 | tail -1
echo "RBA_FS_DEBUG=16 (no score phase: TMA + einsum GEMM + drain only)"; RBA_FS_DEBUG=16 python tools/fused_score_only.py 8 20 2>&1 | tail -1
for a in 1 2 4 8 15; do echo "RBA_FS_ABL=$a"; RBA_FS_ABL=$a python tools/fused_score_only.py 8 20 2>&1 | tail -1; done
} | tee $OUT/ablation.txt
