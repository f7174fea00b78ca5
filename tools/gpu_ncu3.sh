#!/bin/bash
OUT=gpurun_out/${1:-n4}; mkdir -p $OUT
timeout 600 ncu --set full --clock-control none --import-source on -k regex:window_attn_mma -s 0 -c 1 -o $OUT/prof_wattn0 python tools/profile_forward.py --batch 8 > $OUT/ncu1.log 2>&1; echo rc=$?
timeout 600 ncu --set full --clock-control none --import-source on -k regex:window_attn_mma -s 6 -c 1 -o $OUT/prof_wattn2 python tools/profile_forward.py --batch 8 > $OUT/ncu2.log 2>&1; echo rc=$?
