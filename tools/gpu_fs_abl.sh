#!/bin/bash
timeout 300 python -m pytest tests/test_kernels_gpu.py -m gpu -q --tb=short -x -k "einsum_score" 2>&1 | tail -1
for d in ${@:-0 1 2 4 8 6 7 15}; do echo -n "RBA_FS_ABL=$d: "; RBA_FS_ABL=$d python tools/fused_score_only.py 8 10 2>&1 | tail -1; done
