#!/bin/bash
OUT=gpurun_out/r2k; mkdir -p $OUT
export PYTHONUNBUFFERED=1
timeout 300 python -m pytest tests/test_jpeg_decode.py -q -s 2>&1 | grep -E "stream|passed|failed" | tee $OUT/pytest_jpeg.txt
timeout 900 python -m pytest tests -m gpu -q -x -k "model or msda or reference" 2>&1 | tail -4 | tee $OUT/pytest_model.txt
for form in new old; do
  if [ $form = old ]; then export RBA_MSDA_FORM1=1; else unset RBA_MSDA_FORM1; fi
  timeout 600 python tools/profile_forward.py --model swin_b_full 2>&1 | grep -E "total kernel|msda" | tee $OUT/breakdown_msda_$form.txt
done
unset RBA_MSDA_FORM1
timeout 900 python bench.py --model swin_b_full --steps 10 --warmup 3 --no-cpu-baseline --no-gpu-baseline > $OUT/bench_swin_b_full.json 2> $OUT/bench_swin_b_full.err; echo "rc=$?"
python -c "
import json; d=json.load(open('gpurun_out/r2k/bench_swin_b_full.json')); print(d['value'], d['e2e']['value'], d['ms_per_step'], d['clocks'])"
