#!/bin/bash
# round-2 evidence set after the third-generation fused score kernel (one GPU): tests, smoke, bench lines, ncu capture, launch list
OUT=gpurun_out/r2s; mkdir -p $OUT
export PYTHONUNBUFFERED=1
timeout 900 python -m pytest tests -m gpu -q --timeout 300 2>&1 | tail -3 | tee $OUT/pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $OUT/smoke.log 2>&1; echo "smoke rc=$?"; tail -1 $OUT/smoke.log
timeout 600 python bench.py --steps 10 --warmup 3 > $OUT/bench_b8.json 2> $OUT/bench_b8.err; echo "bench rc=$?"
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > $OUT/bench_reference_arm.json 2> $OUT/bench_reference_arm.err; echo "ref arm rc=$?"
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/r2s/bench_*.json')):
    try:
        d=json.load(open(f)); r=d.get('roofline') or {}
        print(f.split('/')[-1], round(d['value'],2), round(d['e2e']['value'],2), round(d.get('ms_per_step',0),2), (d.get('clocks') or {}).get('sm_mhz'), (d.get('gpu_baseline') or {}).get('value'), (d.get('cpu_baseline') or {}).get('value'), r.get('ms_per_launch'), r.get('frac'))
    except Exception as e: print(f, 'ERR', e)
PY
timeout 300 ncu --set full --clock-control none --import-source on -k regex:rba_einsum_score -s 3 -c 1 -o $OUT/fused_score_v3 -f python tools/fused_score_only.py 8 3 > $OUT/ncu_fs.log 2>&1; tail -2 $OUT/ncu_fs.log
timeout 300 python tools/profile_forward.py > $OUT/kernel_breakdown_swin_b_1dl.txt 2>&1; sed -n 3,14p $OUT/kernel_breakdown_swin_b_1dl.txt
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file $OUT/launches_bench.csv python bench.py --steps 2 --warmup 1 --no-graph --no-cpu-baseline --no-gpu-baseline > $OUT/launches_bench.log 2>&1; echo "launch list rc=$?"
python tools/summarize_launches.py $OUT/launches_bench.csv > $OUT/launch_summary_bench.txt 2>&1; head -12 $OUT/launch_summary_bench.txt
