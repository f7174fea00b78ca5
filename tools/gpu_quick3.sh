#!/bin/bash
TAG=${1:-q}
OUT=gpurun_out/$TAG
mkdir -p $OUT
export PYTHONUNBUFFERED=1
timeout 300 python -m pytest tests/test_kernels_gpu.py -m gpu -q --tb=short > $OUT/pytest_k.log 2>&1; echo "kernels rc=$?"; tail -12 $OUT/pytest_k.log
timeout 600 python -m pytest tests/test_model_gpu.py -m gpu -q -rA --tb=short > $OUT/pytest_m.log 2>&1; echo "model rc=$?"; grep -E "^(tiny|swin).*tc \{|passed|failed" $OUT/pytest_m.log | head -12
timeout 900 python bench.py --no-cpu-baseline > $OUT/bench.json 2> $OUT/bench.err; python - <<PY
import json
r=json.load(open("$OUT/bench.json")); print("bench value %.2f img/s  e2e %.2f  ms/step %.1f  score kernel %.3f ms frac %.4f"%(r["value"], r["e2e"]["value"], r["ms_per_step"], r["roofline"]["ms_per_launch"], r["roofline"]["frac"]))
PY
tail -2 $OUT/bench.err
timeout 600 python tools/profile_forward.py --batch 8 > $OUT/profile_b8.txt 2>&1; head -16 $OUT/profile_b8.txt | grep -v Warn
