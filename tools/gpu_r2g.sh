#!/bin/bash
mkdir -p gpurun_out/r2g
timeout 600 python -m pytest tests -m gpu -q -k "panoptic" 2>&1 | grep -E "passed|failed|FAILED|Error" | tee gpurun_out/r2g/pytest_panoptic.txt
bash tools/gpu_ncu_kernels.sh r2f_ncu "gemm_fc1_gelu gemm_fc2_bn256 gemm_qkv_planes conv3x3 layernorm groupnorm_apply"
