#!/bin/bash
# row-based GroupNorm apply: kernel tests, model goldens, per-kernel time
mkdir -p gpurun_out/gn2
timeout 300 python -m pytest tests/test_kernels_gpu.py tests/test_model_gpu.py -m gpu -q -x --timeout 120 -k "groupnorm or golden or taps or full_size or batch_invariance" 2>&1 | tail -4 | tee gpurun_out/gn2/pytest.txt
timeout 200 python tools/profile_forward.py 2>&1 | grep -E "total kernel|gn_|patch_embed" | tee gpurun_out/gn2/breakdown.txt
