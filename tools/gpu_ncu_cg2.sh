#!/bin/bash
OUT=gpurun_out/r2p_ncu; mkdir -p $OUT
cap() {
  timeout 300 ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k "regex:$2" --launch-skip $3 -c 1 -f -o $OUT/$1 python tools/forward_once.py 2 > $OUT/$1.log 2>&1
  echo "$1 rc=$? $(ls -l $OUT/$1.ncu-rep 2>/dev/null | awk '{print $5}')"
}
cap gemm_fc1_gelu_pair 'gemm_tc_kernel<.int.256, .bool.0, .int.2, .bool.1, .bool.0, .bool.1>' 12
cap gemm_qkv_pair 'gemm_tc_kernel<.int.256, .bool.0, .int.0, .bool.1, .bool.0, .bool.1>' 12
cap gemm_fc2_pair 'gemm_tc_kernel<.int.256, .bool.0, .int.0, .bool.0, .bool.0, .bool.1>' 30
