#!/bin/bash
TAG=${1:-q}
OUT=gpurun_out/$TAG
mkdir -p $OUT
export PYTHONUNBUFFERED=1
timeout 300 python -m pytest tests/test_kernels_gpu.py -m gpu -q --tb=short -x > $OUT/pytest_k.log 2>&1; echo "kernels rc=$?"; tail -3 $OUT/pytest_k.log
timeout 600 python -m pytest tests/test_model_gpu.py -m gpu -q -rA --tb=short > $OUT/pytest_m.log 2>&1; echo "model rc=$?"; grep -E "^(tiny|swin).*tc \{|passed|failed" $OUT/pytest_m.log | head -12
for v in 1 0; do
echo "== RBA_TC_BN256=$v"
RBA_TC_BN256=$v timeout 600 python tools/bench_gemm.py > $OUT/bench_gemm_$v.jsonl 2> $OUT/bench_gemm_$v.err; python - <<PY
import json
for l in open("$OUT/bench_gemm_$v.jsonl"):
    r=json.loads(l); print(r["name"], "tc %.0f TF %.3f ms %.0f GB/s"%(r.get("tc_tflops",0), r.get("tc_ms",0), r.get("tc_GBs",0)), "diff %.1e"%r.get("max_abs_diff",-1))
PY
RBA_TC_BN256=$v timeout 900 python bench.py --no-cpu-baseline > $OUT/bench_$v.json 2> $OUT/bench_$v.err; python - <<PY
import json
r=json.load(open("$OUT/bench_$v.json")); print("bench value %.2f img/s  e2e %.2f  ms/step %.1f"%(r["value"], r["e2e"]["value"], r["ms_per_step"]))
PY
done
timeout 600 python tools/profile_forward.py --batch 8 > $OUT/profile_b8.txt 2>&1; head -8 $OUT/profile_b8.txt | grep -v Warn
