#!/bin/bash
# second-generation fused score kernel: quick liveness check, parity tests, then the kernel timed alone
mkdir -p gpurun_out/fs3
export RBA_FS_VARIANT=2
timeout 120 python tools/fused_score_only.py 1 1 > gpurun_out/fs3/live.txt 2>&1; rc=$?
tail -3 gpurun_out/fs3/live.txt
if [ $rc -ne 0 ]; then echo "liveness check failed rc=$rc"; exit 1; fi
timeout 420 python -m pytest tests/test_kernels_gpu.py -m gpu -q -k "einsum_score" 2>&1 | tail -25 | tee gpurun_out/fs3/pytest.txt
for v in 2 1; do
  RBA_FS_VARIANT=$v timeout 120 python tools/fused_score_only.py 8 10 2>&1 | tail -1 | tee gpurun_out/fs3/time_v$v.txt
done
for d in 1 2 4 6; do
  echo "RBA_FS_DEBUG=$d"; RBA_FS_DEBUG=$d timeout 120 python tools/fused_score_only.py 8 10 2>&1 | tail -1 | tee gpurun_out/fs3/time_dbg$d.txt
done
for d in 0; do
  RBA_FS_TIMELINE=1 RBA_FS_DEBUG=$d timeout 120 python tools/fused_score_only.py 8 1 2>&1 | tail -44 > gpurun_out/fs3/timeline_dbg$d.txt
done
