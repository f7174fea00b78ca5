#!/bin/bash
# fused score kernel v2 A/B: compute warps (16 / 12) x cells per warp iteration (1 / 2) x reciprocal batching (pair / quad)
OUT=gpurun_out/${1:-fs2}; mkdir -p $OUT
export PYTHONUNBUFFERED=1
timeout 600 python -m pytest tests/test_kernels_gpu.py -q -x -k "score or einsum" 2>&1 | tail -3 | tee $OUT/pytest_score.txt
{
for lib in "" rba_b200/lib/librba_b200_cw12.so; do
  for nc in 1 2; do for rcp in 1 2; do
    echo "lib=${lib:-default(cw16)} NCELL=$nc RCP=$rcp"
    RBA_B200_LIB=$lib RBA_FS_NCELL=$nc RBA_FS_RCP=$rcp python tools/fused_score_only.py 8 20 2>&1 | tail -1
  done; done
done
echo "skeleton (RBA_FS_ABL=15) cw16 / cw12"
RBA_FS_ABL=15 python tools/fused_score_only.py 8 20 2>&1 | tail -1
RBA_B200_LIB=rba_b200/lib/librba_b200_cw12.so RBA_FS_ABL=15 python tools/fused_score_only.py 8 20 2>&1 | tail -1
echo "no score phase (RBA_FS_DEBUG=16)"
RBA_FS_DEBUG=16 python tools/fused_score_only.py 8 20 2>&1 | tail -1
} | tee $OUT/ab.txt
