#!/bin/bash
TAG=${1:-n2}
OUT=gpurun_out/$TAG
mkdir -p $OUT
python tools/score_only.py 8
timeout 600 ncu --set full --clock-control none --import-source on -k regex:rba_score -s 3 -c 1 -o $OUT/prof_score python tools/score_only.py 2 > $OUT/ncu.log 2>&1; echo rc=$?
timeout 600 ncu --set full --clock-control none --import-source on -k regex:window_attn_mma -s 2 -c 1 -o $OUT/prof_wattn python tools/profile_forward.py --batch 2 > $OUT/ncu2.log 2>&1; echo rc=$?
