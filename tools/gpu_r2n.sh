#!/bin/bash
OUT=gpurun_out/r2n; mkdir -p $OUT
export PYTHONUNBUFFERED=1
timeout 900 python -m pytest tests -m gpu -q -x -k "gemm or conv or model or golden" 2>&1 | tail -3 | tee $OUT/pytest.txt
timeout 600 python tools/profile_forward.py > $OUT/kernel_breakdown.txt 2>&1; sed -n 3,12p $OUT/kernel_breakdown.txt
timeout 900 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-gpu-baseline > $OUT/bench_b8.json 2> $OUT/bench_b8.err; echo "bench rc=$?"
python -c "
import json; d=json.load(open('gpurun_out/r2n/bench_b8.json')); print(d['value'], d['e2e']['value'], d['ms_per_step'], d['clocks'])"
ls tools/gemm_shapes.py >/dev/null 2>&1 && timeout 300 python tools/gemm_shapes.py 2>&1 | tail -10 | tee $OUT/gemm_shapes.txt
