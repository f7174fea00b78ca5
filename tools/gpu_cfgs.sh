#!/bin/bash
# BASELINE.json configs[3] (Swin-L 1dl) and configs[4] (Swin-B full decoder, 3 levels, 10 heads) at 8x1024x2048 on one GPU
OUT=gpurun_out/${1:-cfgs}; mkdir -p $OUT
export PYTHONUNBUFFERED=1
for m in swin_l_1dl swin_b_full; do
  timeout 600 python bench.py --model $m --no-cpu-baseline --steps 5 --warmup 3 > $OUT/bench_$m.json 2> $OUT/bench_$m.err; echo "$m rc=$?"
  python - <<PY
import json
try:
    r=json.load(open("$OUT/bench_$m.json")); print("$m value %.2f img/s  e2e %.2f  ms/step %.1f  launches/step %d clocks %s"%(r["value"], r["e2e"]["value"], r["ms_per_step"], r["gpu_launches_per_step"], r["clocks"]))
except Exception as e:
    print("no result", e); print(open("$OUT/bench_$m.err").read()[-800:])
PY
done
