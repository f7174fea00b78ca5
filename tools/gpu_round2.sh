#!/bin/bash
# tcgen05 bring-up round: kernel tests for the TC backend first (bounded by timeout), then model tests, GEMM bench, bench.
TAG=${1:-r1b}
OUT=gpurun_out/$TAG
mkdir -p $OUT
export PYTHONUNBUFFERED=1
echo "== tc kernel tests"; timeout 300 python -m pytest tests/test_kernels_gpu.py -m gpu -q -rA --tb=short -k "tc or patch_embed or score_golden" > $OUT/pytest_tc.log 2>&1; echo "rc=$?"; tail -30 $OUT/pytest_tc.log
echo "== tc model tests"; timeout 600 python -m pytest tests/test_model_gpu.py -m gpu -q -rA --tb=short -k "tc_backend" > $OUT/pytest_tc_model.log 2>&1; echo "rc=$?"; grep -E "tc \{|passed|failed|Error" $OUT/pytest_tc_model.log | head -20
echo "== gemm bench"; timeout 600 python tools/bench_gemm.py > $OUT/bench_gemm.jsonl 2> $OUT/bench_gemm.err; echo "rc=$?"; cat $OUT/bench_gemm.jsonl; tail -3 $OUT/bench_gemm.err
echo "== bench tc B=8"; timeout 900 python bench.py --backend tc --no-cpu-baseline > $OUT/bench_tc.json 2> $OUT/bench_tc.err; echo "rc=$?"; cat $OUT/bench_tc.json; tail -3 $OUT/bench_tc.err
echo "== ncu launch list tc (B=1, eager)"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file $OUT/launches_tc.csv \
  python bench.py --backend tc --batch 1 --steps 1 --warmup 0 --no-graph --no-cpu-baseline > $OUT/ncu_bench.log 2>&1; echo "ncu rc=$?"
python tools/summarize_launches.py $OUT/launches_tc.csv > $OUT/launch_summary_tc.txt 2>&1; head -24 $OUT/launch_summary_tc.txt
