#!/bin/bash
# third-generation fused score kernel: timing against generation 1 and the ablations, then parity (short timeouts: a hang must
# not eat the GPU budget)
mkdir -p gpurun_out/fs5
for v in 3 1; do echo "== variant $v"; RBA_FS_VARIANT=$v timeout 40 python tools/fused_score_only.py 8 20 2>&1 | tail -1; done | tee gpurun_out/fs5/timing.txt
grep -q "ms/launch" gpurun_out/fs5/timing.txt || { echo "variant 3 did not finish"; exit 1; }
timeout 200 python -m pytest tests/test_kernels_gpu.py -m gpu -q -x -k "einsum_score" --timeout 60 2>&1 | tail -5 | tee gpurun_out/fs5/pytest.txt
for d in ${@:-32 1 4 8 12}; do echo "== variant 3 RBA_FS_DEBUG=$d"; RBA_FS_VARIANT=3 RBA_FS_DEBUG=$d timeout 40 python tools/fused_score_only.py 8 20 2>&1 | tail -1; done | tee -a gpurun_out/fs5/timing.txt
