#!/bin/bash
# final evidence set of round 2 (one GPU): tests, smoke, bench (+ reference arm), per-kernel breakdown
OUT=gpurun_out/r2t; mkdir -p $OUT
export PYTHONUNBUFFERED=1
timeout 900 python -m pytest tests -m gpu -q --timeout 300 2>&1 | tail -3 | tee $OUT/pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $OUT/smoke.log 2>&1; echo "smoke rc=$?"; tail -1 $OUT/smoke.log
timeout 600 python bench.py --steps 10 --warmup 3 > $OUT/bench_b8.json 2> $OUT/bench_b8.err; echo "bench rc=$?"
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > $OUT/bench_reference_arm.json 2> $OUT/bench_reference_arm.err; echo "ref arm rc=$?"
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/r2t/bench_*.json')):
    try:
        d=json.load(open(f)); r=d.get('roofline') or {}
        print(f.split('/')[-1], round(d['value'],2), round(d['e2e']['value'],2), round(d.get('ms_per_step',0),2), (d.get('clocks') or {}).get('sm_mhz'), (d.get('gpu_baseline') or {}).get('value'), (d.get('cpu_baseline') or {}).get('value'), r.get('ms_per_launch'), r.get('frac'))
    except Exception as e: print(f, 'ERR', e)
PY
timeout 300 python tools/profile_forward.py > $OUT/kernel_breakdown_swin_b_1dl.txt 2>&1; sed -n 3,5p $OUT/kernel_breakdown_swin_b_1dl.txt
