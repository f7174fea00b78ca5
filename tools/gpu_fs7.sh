#!/bin/bash
# third-generation fused score kernel: in-kernel timeline of CTA 0 (RBA_FS_TIMELINE)
mkdir -p gpurun_out/fs7
for d in ${@:-0 2}; do echo "== variant 3 RBA_FS_DEBUG=$d"; RBA_FS_TIMELINE=1 RBA_FS_VARIANT=3 RBA_FS_DEBUG=$d timeout 120 python tools/fused_score_only.py 8 1 2>&1 | tail -26; done | tee gpurun_out/fs7/timeline.txt
