#!/bin/bash
# Final evidence round: full gpu tests, bench (+reference arm), ncu launch list of the bench command, ncu full of the score kernel.
OUT=gpurun_out/${1:-final}; mkdir -p $OUT
export PYTHONUNBUFFERED=1
timeout 900 python -m pytest tests -m gpu -q --tb=short > $OUT/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -2 $OUT/pytest_gpu.log
timeout 600 python __graft_entry__.py smoke > $OUT/smoke.log 2>&1; tail -1 $OUT/smoke.log
timeout 900 python bench.py > $OUT/bench.json 2> $OUT/bench.err; echo "bench rc=$?"; cat $OUT/bench.json
timeout 900 python bench.py --impl reference --steps 2 --warmup 1 > $OUT/bench_reference.json 2> $OUT/bench_reference.err; cat $OUT/bench_reference.json
timeout 600 python tools/profile_forward.py --batch 8 > $OUT/profile_b8.txt 2>&1
timeout 600 python tools/bench_gemm.py > $OUT/bench_gemm.jsonl 2>/dev/null
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file $OUT/launches_bench.csv \
  python bench.py --steps 1 --warmup 1 --no-graph --no-cpu-baseline > $OUT/ncu_bench.log 2>&1; echo "ncu rc=$?"
python tools/summarize_launches.py $OUT/launches_bench.csv > $OUT/launch_summary_bench.txt 2>&1; head -12 $OUT/launch_summary_bench.txt
timeout 600 ncu --set full --clock-control none --import-source on -k regex:rba_score_mma -s 3 -c 1 -o $OUT/prof_score_b8 python tools/score_only.py 8 > $OUT/ncu_score.log 2>&1; echo "ncu score rc=$?"
