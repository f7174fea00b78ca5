#!/bin/bash
TAG=${1:-n1}
OUT=gpurun_out/$TAG
mkdir -p $OUT
export PYTHONUNBUFFERED=1
timeout 600 python -m pytest tests/test_model_gpu.py -m gpu -q -rA --tb=short > $OUT/pytest_m.log 2>&1; echo "model rc=$?"; grep -E "^(tiny|swin).*\{|passed|failed" $OUT/pytest_m.log | head -12
# stage-0 block 0: launches of gemm_tc_kernel in order: qkv(0) proj(1) fc1(2) fc2(3) ...
timeout 600 ncu --set full --clock-control none --import-source on -k regex:gemm_tc_kernel -s 4 -c 4 -o $OUT/prof_gemm_blk \
  python tools/profile_forward.py --batch 4 > $OUT/ncu_gemm.log 2>&1; echo "rc=$?"
ls -la $OUT
