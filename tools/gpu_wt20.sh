#!/bin/bash
# window attention: 4 softmax warps per TMEM lane quadrant (20 warps) vs 2 (12 warps)
OUT=gpurun_out/wt20; mkdir -p $OUT
export PYTHONUNBUFFERED=1
for parts in 6 4; do
  echo "== RBA_WT_PARTS=$parts"
  RBA_WT_PARTS=$parts timeout 300 python -m pytest tests/test_kernels_gpu.py -q -k "window_attention_tensor_core or tiled_qkv" 2>&1 | grep -E "passed|failed|FAILED|Error" | tee $OUT/pytest_wattn_p$parts.txt
  RBA_WT_PARTS=$parts timeout 300 python tools/bench_wattn.py 8 2>&1 | tail -1 | tee $OUT/bench_wattn_p$parts.txt
  for d in 0 3 4 7; do echo "RBA_WT_DEBUG=$d"; RBA_WT_PARTS=$parts RBA_WT_DEBUG=$d timeout 120 python tools/bench_wattn_one.py 2 8 10 2>&1 | tail -1; done | tee $OUT/ablation_p$parts.txt
done
