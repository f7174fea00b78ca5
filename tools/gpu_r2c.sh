#!/bin/bash
mkdir -p gpurun_out/r2c
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r2c/pytest_gpu.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/r2c/pytest_gpu.log
tail -4 gpurun_out/r2c/pytest_gpu.log
timeout 900 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-gpu-baseline > gpurun_out/r2c/bench_b8.json 2> gpurun_out/r2c/bench_b8.err; echo "bench rc=$?"
python -c "
import json;d=json.load(open('gpurun_out/r2c/bench_b8.json'));print(d['value'],d['e2e']['value'],d['ms_per_step'],d['clocks'],d['roofline']['ms_per_launch'])"
timeout 600 python tools/profile_forward.py > gpurun_out/r2c/kernel_breakdown.txt 2>&1; tail -40 gpurun_out/r2c/kernel_breakdown.txt
