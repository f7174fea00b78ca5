#!/bin/bash
# JPEG decode test + ncu --set full of the 3-level MSDeformAttn and the split-KV decoder attention (swin_b_full)
OUT=gpurun_out/r2j_ncu; mkdir -p $OUT gpurun_out/r2j
timeout 300 python -m pytest tests/test_jpeg_decode.py -q 2>&1 | tail -4 | tee gpurun_out/r2j/pytest_jpeg.txt
cap() {
  timeout 300 ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k "regex:$2" --launch-skip $3 -c 1 -f -o $OUT/$1 python tools/forward_once.py 2 swin_b_full > $OUT/$1.log 2>&1
  echo "$1 rc=$? $(ls -l $OUT/$1.ncu-rep 2>/dev/null | awk '{print $5}')"
}
cap msda_fused_3lvl 'msda_fused_kernel' 8
cap mha_split_full 'mha_split_kernel' 20
cap window_attn_tc_p4 'window_attn_tc_kernel' 30
