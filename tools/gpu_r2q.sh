#!/bin/bash
OUT=gpurun_out/r2q; mkdir -p $OUT
export PYTHONUNBUFFERED=1
timeout 300 python -m pytest tests -m gpu -q -x -k "gemm or model_full or golden" 2>&1 | tail -3 | tee $OUT/pytest.log
for stg in 1 0; do
  echo "== RBA_TC_CG2_STG=$stg"
  RBA_TC_CG2_STG=$stg timeout 600 python tools/profile_forward.py > $OUT/kernel_breakdown_stg$stg.txt 2>&1; sed -n 3,9p $OUT/kernel_breakdown_stg$stg.txt
done
timeout 900 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-gpu-baseline > $OUT/bench_b8.json 2> $OUT/bench_b8.err; echo "bench rc=$?"
python -c "
import json; d=json.load(open('gpurun_out/r2q/bench_b8.json')); print(d['value'], d['e2e']['value'], d['ms_per_step'], d['clocks'])"
