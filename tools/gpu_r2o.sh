#!/bin/bash
OUT=gpurun_out/r2o; mkdir -p $OUT
export PYTHONUNBUFFERED=1
timeout 1500 python -m pytest tests -m gpu -q -x 2>&1 | tail -3 | tee $OUT/pytest_gpu.log
for cg in 1 0; do
  echo "== RBA_TC_CG2=$cg"
  RBA_TC_CG2=$cg timeout 600 python tools/profile_forward.py > $OUT/kernel_breakdown_cg$cg.txt 2>&1; sed -n 3,10p $OUT/kernel_breakdown_cg$cg.txt
  RBA_TC_CG2=$cg timeout 900 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-gpu-baseline > $OUT/bench_b8_cg$cg.json 2> $OUT/bench_b8_cg$cg.err; echo "bench rc=$?"
  python -c "
import json; d=json.load(open('gpurun_out/r2o/bench_b8_cg$cg.json')); print(d['value'], d['e2e']['value'], d['ms_per_step'], d['clocks'])"
done
