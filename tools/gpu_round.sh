#!/bin/bash
# One GPU-box visit: smoke, parity tests, bench, ncu launch list.  Everything lands in gpurun_out/.
# usage: tools/gpu_round.sh [tag]
TAG=${1:-r1}
OUT=gpurun_out/$TAG
mkdir -p $OUT
export PYTHONUNBUFFERED=1
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv > $OUT/gpu.txt 2>&1
nproc >> $OUT/gpu.txt; lscpu | grep "Model name" >> $OUT/gpu.txt
echo "== smoke"; timeout 600 python __graft_entry__.py smoke > $OUT/smoke.log 2>&1; echo "smoke rc=$?"; tail -3 $OUT/smoke.log
echo "== pytest gpu"; timeout 1500 python -m pytest tests -m gpu -q -rA --tb=short > $OUT/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -40 $OUT/pytest_gpu.log
echo "== bench B=1"; timeout 900 python bench.py --batch 1 --steps 5 --warmup 3 --no-cpu-baseline > $OUT/bench_b1.json 2> $OUT/bench_b1.err; echo "rc=$?"; cat $OUT/bench_b1.json; tail -3 $OUT/bench_b1.err
echo "== bench default (B=8)"; timeout 1200 python bench.py > $OUT/bench.json 2> $OUT/bench.err; echo "rc=$?"; cat $OUT/bench.json; tail -3 $OUT/bench.err
echo "== ncu launch list (B=1, eager)"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file $OUT/launches.csv \
  python bench.py --batch 1 --steps 1 --warmup 0 --no-graph --no-cpu-baseline > $OUT/ncu_bench.log 2>&1; echo "ncu rc=$?"
python tools/summarize_launches.py $OUT/launches.csv > $OUT/launch_summary.txt 2>&1; head -40 $OUT/launch_summary.txt
