#!/bin/bash
OUT=gpurun_out/${1:-fspipe}; mkdir -p $OUT
export PYTHONUNBUFFERED=1
timeout 600 python -m pytest tests/test_kernels_gpu.py tests/test_model_gpu.py -m gpu -q --tb=short -x -k "einsum_score or fused or golden" > $OUT/pytest_fs.log 2>&1; echo "fs tests rc=$?"; tail -3 $OUT/pytest_fs.log
for m in 0 1; do echo "RBA_FS_PIPE=$m"; RBA_FS_PIPE=$m python tools/fused_score_only.py 8 20 2>&1 | tail -1; done | tee $OUT/pipe_modes.txt
