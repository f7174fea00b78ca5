#!/bin/bash
OUT=gpurun_out/${1:-r4}; mkdir -p $OUT
export PYTHONUNBUFFERED=1
timeout 900 python -m pytest tests -m gpu -q --tb=short > $OUT/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -3 $OUT/pytest_gpu.log
timeout 600 python tools/torch_gpu_baseline.py 2 > $OUT/torch_gpu_baseline.txt 2>&1; tail -2 $OUT/torch_gpu_baseline.txt
timeout 600 python bench.py --model swin_l_1dl --no-cpu-baseline --steps 5 > $OUT/bench_swin_l.json 2> $OUT/bench_swin_l.err; python -c "
import json; r=json.load(open('$OUT/bench_swin_l.json')); print('swin_l: %.2f img/s e2e %.2f ms/step %.1f'%(r['value'], r['e2e']['value'], r['ms_per_step']))"
timeout 600 python __graft_entry__.py smoke 2>&1 | tail -1
