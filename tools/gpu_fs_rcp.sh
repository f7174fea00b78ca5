#!/bin/bash
# A/B of the fused kernel's reciprocal batching modes (RBA_FS_RCP) + its parity tests
OUT=gpurun_out/${1:-fsrcp}; mkdir -p $OUT
export PYTHONUNBUFFERED=1
timeout 600 python -m pytest tests/test_kernels_gpu.py tests/test_model_gpu.py -m gpu -q --tb=short -x -k "einsum_score or fused or golden" > $OUT/pytest_fs.log 2>&1; echo "fs tests rc=$?"; tail -3 $OUT/pytest_fs.log
for m in 0 3 1 2; do echo "RBA_FS_RCP=$m"; RBA_FS_RCP=$m python tools/fused_score_only.py 8 20 2>&1 | tail -1; done | tee $OUT/rcp_modes.txt
