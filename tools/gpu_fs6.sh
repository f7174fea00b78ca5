#!/bin/bash
# third-generation fused score kernel: ablations (RBA_FS_DEBUG bits: 2 no score phase, 4 no sigmoid math, 8 no MMAs, 16 no epilogue math)
mkdir -p gpurun_out/fs6
for d in ${@:-0 2 4 8 12 16 28}; do echo "== variant 3 RBA_FS_DEBUG=$d"; RBA_FS_VARIANT=3 RBA_FS_DEBUG=$d timeout 120 python tools/fused_score_only.py 8 20 2>&1 | tail -1; done | tee gpurun_out/fs6/ablation.txt
