#!/bin/bash
# round-2 validation after the fused-score rework: the whole GPU test suite, smoke, the driver's bench line
OUT=gpurun_out/r2h; mkdir -p $OUT
export PYTHONUNBUFFERED=1
timeout 1500 python -m pytest tests -m gpu -q -x 2>&1 | tail -6 | tee $OUT/pytest_gpu.log
python -c "import __graft_entry__ as g; g.smoke()" > $OUT/smoke.log 2>&1; echo "smoke rc=$?"; tail -1 $OUT/smoke.log
timeout 900 python bench.py --steps 10 --warmup 3 > $OUT/bench_b8.json 2> $OUT/bench_b8.err; echo "bench rc=$?"
RBA_FS_VARIANT=2 timeout 900 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-gpu-baseline > $OUT/bench_b8_score_v2.json 2> $OUT/bench_b8_score_v2.err; echo "bench v2 rc=$?"
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/r2h/bench_*.json')):
    try:
        d=json.load(open(f)); print(f.split('/')[-1], round(d['value'],2), round(d['e2e']['value'],2), round(d.get('ms_per_step',0),2), d['roofline']['ms_per_launch'], d['roofline']['frac'], (d.get('gpu_baseline') or {}).get('value'), (d.get('cpu_baseline') or {}).get('value'))
    except Exception as e: print(f, 'ERR', e)
PY
