#!/bin/bash
# round 2, first GPU pass: full GPU test suite (incl. full-size goldens + live reference parity), bench with gpu_baseline,
# evaluate_ood.py unchanged on the GPU, MSDA vs the reference kernel.
mkdir -p gpurun_out/r2a
nvidia-smi --query-gpu=name,clocks.max.sm,power.limit --format=csv > gpurun_out/r2a/gpu.txt 2>&1
nproc >> gpurun_out/r2a/gpu.txt
timeout 1500 python -m pytest tests -m gpu -x -q -s > gpurun_out/r2a/pytest_gpu.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/r2a/pytest_gpu.log
tail -5 gpurun_out/r2a/pytest_gpu.log
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/r2a/bench_b8.json 2> gpurun_out/r2a/bench_b8.err; echo "bench rc=$?"
cat gpurun_out/r2a/bench_b8.json | head -c 6000
timeout 600 python tools/msda_vs_reference.py > gpurun_out/r2a/msda_vs_reference.log 2>&1; tail -4 gpurun_out/r2a/msda_vs_reference.log
timeout 900 python tools/run_evaluate_ood_gpu.py --arch swin_b_1dl --images 16 --out gpurun_out/r2a/evaluate_ood_swin_b_1dl.json > gpurun_out/r2a/evaluate_ood.log 2>&1; echo "evaluate_ood rc=$?"; tail -3 gpurun_out/r2a/evaluate_ood.log | head -c 3000
