#!/bin/bash
OUT=gpurun_out/${1:-fs2}; mkdir -p $OUT
export PYTHONUNBUFFERED=1
python tools/fused_score_only.py 8 10 2>&1 | tail -1
timeout 900 python bench.py --no-cpu-baseline > $OUT/bench.json 2> $OUT/bench.err; python - <<PY
import json
r=json.load(open("$OUT/bench.json")); print("bench value %.2f img/s  e2e %.2f  ms/step %.1f  roofline kernel %.3f ms frac %.4f launches/step %d"%(r["value"], r["e2e"]["value"], r["ms_per_step"], r["roofline"]["ms_per_launch"], r["roofline"]["frac"], r["gpu_launches_per_step"]))
PY
tail -2 $OUT/bench.err
timeout 600 ncu --set full --clock-control none --import-source on -k regex:rba_einsum_score -s 2 -c 1 -o $OUT/prof_fused_score python tools/fused_score_only.py 2 2 > $OUT/ncu_fs.log 2>&1; echo "ncu rc=$?"
