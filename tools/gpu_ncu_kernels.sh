#!/bin/bash
# `ncu --set full` capture of ONE launch of every kernel family of the forward (B = 2, 1024 x 2048, second forward of the process)
OUT=gpurun_out/${1:-ncuk}; mkdir -p $OUT
ONLY=${2:-}
cap() {  # name, regex, launches to skip (lands in the 2nd / 3rd forward)
  if [ -n "$ONLY" ] && ! echo " $ONLY " | grep -q " $1 "; then return; fi
  timeout 300 ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k "regex:$2" --launch-skip $3 -c 1 -f -o $OUT/$1 python tools/forward_once.py 2 > $OUT/$1.log 2>&1
  echo "$1 rc=$? $(ls -la $OUT/$1.ncu-rep 2>/dev/null | awk '{print $5}')"
}
cap gemm_fc1_gelu   'gemm_tc_kernel<.int.128, .bool.0, .int.2, .bool.1, .bool.1>' 30
cap gemm_fc2_bn256  'gemm_tc_kernel<.int.256, .bool.0, .int.0, .bool.0, .bool.0>' 45
cap gemm_qkv_planes 'gemm_tc_kernel<.int.128, .bool.0, .int.0, .bool.1, .bool.1>' 35
cap conv3x3         'gemm_tc_kernel<.int.256, .bool.1, .int.0, .bool.0, .bool.0>' 5
cap layernorm       'layernorm_kernel<.int.4>' 50
cap groupnorm_apply 'gn_apply_kernel<.bool.1, .int.2>' 5
cap groupnorm_stats 'gn_stats_kernel' 13
cap msda_fused      'msda_fused_kernel' 8
cap mha_split       'mha_split_kernel' 3
cap patch_embed     'patch_embed_kernel' 1
cap window_attn_tc  'window_attn_tc_kernel' 34
cap fused_score     'rba_einsum_score_kernel' 1
