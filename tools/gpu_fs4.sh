#!/bin/bash
mkdir -p gpurun_out/fs4
for d in 0 6; do
  echo "== RBA_FS_DEBUG=$d"; RBA_FS_TIMELINE=1 RBA_FS_DEBUG=$d timeout 120 python tools/fused_score_only.py 8 1 2>&1 | tail -44 | tee gpurun_out/fs4/timeline_dbg$d.txt
done
