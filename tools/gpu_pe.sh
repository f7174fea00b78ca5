#!/bin/bash
# patch-embed rewrite (8 tokens per warp): kernel test, model goldens, per-kernel time
mkdir -p gpurun_out/pe
timeout 300 python -m pytest tests/test_kernels_gpu.py tests/test_model_gpu.py -m gpu -q -x --timeout 120 -k "patch_embed or golden or taps or full_size" 2>&1 | tail -4 | tee gpurun_out/pe/pytest.txt
timeout 200 python tools/profile_forward.py 2>&1 | grep -E "total kernel|patch_embed" | tee gpurun_out/pe/breakdown.txt
