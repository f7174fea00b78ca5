#!/bin/bash
OUT=gpurun_out/${1:-n3}; mkdir -p $OUT
timeout 900 ncu --set full --clock-control none --import-source on -k regex:gemm_tc_kernel -s 18 -c 4 -o $OUT/prof_gemm_s2 python tools/profile_forward.py --batch 8 > $OUT/ncu.log 2>&1; echo rc=$?
