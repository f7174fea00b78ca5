#!/bin/bash
# Evidence round: full gpu tests, smoke, bench (+reference arm), torch-profiler breakdown, ncu launch list of the bench command,
# ncu --set full of the fused mask-einsum + score kernel (raw/details pages exported as text).
OUT=gpurun_out/${1:-evidence}; mkdir -p $OUT
export PYTHONUNBUFFERED=1
nvidia-smi --query-gpu=name,clocks.max.sm,clocks.sm,power.limit --format=csv > $OUT/gpu.txt 2>&1
python -c "import os; print('cpus', os.cpu_count())" >> $OUT/gpu.txt
timeout 1200 python -m pytest tests -m gpu -q --tb=short > $OUT/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -3 $OUT/pytest_gpu.log
timeout 600 python __graft_entry__.py smoke > $OUT/smoke.log 2>&1; tail -1 $OUT/smoke.log
timeout 900 python bench.py > $OUT/bench.json 2> $OUT/bench.err; echo "bench rc=$?"; cat $OUT/bench.json; tail -3 $OUT/bench.err
timeout 900 python bench.py --impl reference --steps 2 --warmup 1 > $OUT/bench_reference.json 2> $OUT/bench_reference.err; cat $OUT/bench_reference.json
timeout 600 python tools/profile_forward.py --batch 8 > $OUT/profile_b8.txt 2>&1; head -40 $OUT/profile_b8.txt
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file $OUT/launches_bench.csv \
  python bench.py --steps 1 --warmup 1 --no-graph --no-cpu-baseline > $OUT/ncu_bench.log 2>&1; echo "ncu rc=$?"
python tools/summarize_launches.py $OUT/launches_bench.csv > $OUT/launch_summary_bench.txt 2>&1; head -16 $OUT/launch_summary_bench.txt
timeout 600 ncu --set full --clock-control none --import-source on -k regex:rba_einsum_score -s 2 -c 1 -o $OUT/prof_fused_score python tools/fused_score_only.py 2 2 > $OUT/ncu_fs.log 2>&1; echo "ncu fused rc=$?"
ncu -i $OUT/prof_fused_score.ncu-rep --page raw --csv > $OUT/prof_fused_score_raw.csv 2>/dev/null
ncu -i $OUT/prof_fused_score.ncu-rep --page details > $OUT/prof_fused_score_details.txt 2>/dev/null
