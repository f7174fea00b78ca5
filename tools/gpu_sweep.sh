#!/bin/bash
TAG=${1:-s1}
OUT=gpurun_out/$TAG
mkdir -p $OUT
export PYTHONUNBUFFERED=1
RBA_PROFILE_SEQ=1 timeout 600 python tools/profile_forward.py --batch 8 > $OUT/profile_seq_b8.txt 2>&1; grep -c gemm_tc $OUT/profile_seq_b8.txt
for b in 1 2 4 16; do
  timeout 600 python bench.py --batch $b --steps 5 --warmup 3 --no-cpu-baseline > $OUT/bench_b$b.json 2> $OUT/bench_b$b.err
  python - <<PY
import json
r=json.load(open("$OUT/bench_b$b.json")); print("batch $b: %.2f img/s  e2e %.2f  ms/step %.1f"%(r["value"], r["e2e"]["value"], r["ms_per_step"]))
PY
done
