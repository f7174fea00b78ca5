#!/bin/bash
TAG=${1:-r1c}
OUT=gpurun_out/$TAG
mkdir -p $OUT
export PYTHONUNBUFFERED=1
echo "== profiler B=8 tc"; timeout 600 python tools/profile_forward.py --batch 8 > $OUT/profile_b8_tc.txt 2>&1; cat $OUT/profile_b8_tc.txt | head -30
echo "== ncu full: score kernel"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:rba_score_kernel -s 2 -c 1 -o $OUT/prof_score \
  python bench.py --batch 2 --steps 1 --warmup 0 --no-graph --no-cpu-baseline > $OUT/ncu_score.log 2>&1; echo "rc=$?"
echo "== ncu full: gemm_tc (stage0 fc1 + stage2 fc1) and window attention"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:gemm_tc_kernel -s 6 -c 2 -o $OUT/prof_gemm_s0 \
  python tools/profile_forward.py --batch 2 > $OUT/ncu_gemm.log 2>&1; echo "rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:gemm_tc_kernel -s 40 -c 3 -o $OUT/prof_gemm_s2 \
  python tools/profile_forward.py --batch 2 > $OUT/ncu_gemm2.log 2>&1; echo "rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:window_attn_kernel -s 1 -c 1 -o $OUT/prof_wattn \
  python tools/profile_forward.py --batch 2 > $OUT/ncu_wattn.log 2>&1; echo "rc=$?"
ls -la $OUT
