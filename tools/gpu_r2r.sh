#!/bin/bash
# round-2 closing evidence set (one GPU): tests, smoke, the driver's bench lines, other model configs, launch list
OUT=gpurun_out/r2r; mkdir -p $OUT
export PYTHONUNBUFFERED=1
timeout 1500 python -m pytest tests -m gpu -q 2>&1 | tail -3 | tee $OUT/pytest_gpu.log
python -c "import __graft_entry__ as g; g.smoke()" > $OUT/smoke.log 2>&1; echo "smoke rc=$?"; tail -1 $OUT/smoke.log
timeout 900 python bench.py --steps 10 --warmup 3 > $OUT/bench_b8.json 2> $OUT/bench_b8.err; echo "bench rc=$?"
timeout 900 python bench.py --impl reference --steps 3 --warmup 1 > $OUT/bench_reference_arm.json 2> $OUT/bench_reference_arm.err; echo "ref arm rc=$?"
for m in swin_l_1dl swin_b_full r50_1dl; do
  timeout 900 python bench.py --model $m --steps 10 --warmup 3 --no-cpu-baseline --no-gpu-baseline > $OUT/bench_$m.json 2> $OUT/bench_$m.err; echo "$m rc=$?"
done
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/r2r/bench_*.json')):
    try:
        d=json.load(open(f)); print(f.split('/')[-1], round(d['value'],2), round(d['e2e']['value'],2), round(d.get('ms_per_step',0),2), (d.get('clocks') or {}).get('sm_mhz'), (d.get('gpu_baseline') or {}).get('value'), (d.get('cpu_baseline') or {}).get('value'))
    except Exception as e: print(f, 'ERR', e)
PY
timeout 600 python tools/profile_forward.py > $OUT/kernel_breakdown_swin_b_1dl.txt 2>&1; sed -n 3,14p $OUT/kernel_breakdown_swin_b_1dl.txt
timeout 600 python tools/profile_forward.py --model swin_b_full > $OUT/kernel_breakdown_swin_b_full.txt 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file $OUT/launches_bench.csv python bench.py --steps 2 --warmup 1 --no-graph --no-cpu-baseline --no-gpu-baseline > $OUT/launches_bench.log 2>&1; echo "launch list rc=$?"
python tools/summarize_launches.py $OUT/launches_bench.csv > $OUT/launch_summary_bench.txt 2>&1; head -12 $OUT/launch_summary_bench.txt
