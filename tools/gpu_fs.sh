#!/bin/bash
# fused einsum+score kernel: parity tests, then bench
OUT=gpurun_out/${1:-fs}; mkdir -p $OUT
export PYTHONUNBUFFERED=1
timeout 300 python -m pytest tests/test_kernels_gpu.py -m gpu -q --tb=short -x -k "einsum_score" > $OUT/pytest_fs.log 2>&1; echo "fs kernel rc=$?"; tail -25 $OUT/pytest_fs.log
timeout 600 python -m pytest tests/test_model_gpu.py -m gpu -q -rA --tb=short -k "fused_score or module_interface or batch_invariance" > $OUT/pytest_m.log 2>&1; echo "model rc=$?"; grep -E "fused \{|passed|failed|Error|error" $OUT/pytest_m.log | head -20
if [ "$2" == "bench" ]; then
timeout 900 python bench.py --no-cpu-baseline > $OUT/bench.json 2> $OUT/bench.err; python - <<PY
import json
r=json.load(open("$OUT/bench.json")); print("bench value %.2f img/s  e2e %.2f  ms/step %.1f  roofline kernel %.3f ms frac %.4f"%(r["value"], r["e2e"]["value"], r["ms_per_step"], r["roofline"]["ms_per_launch"], r["roofline"]["frac"]))
PY
tail -2 $OUT/bench.err
fi
