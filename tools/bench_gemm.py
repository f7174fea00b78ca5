"""GEMM micro-benchmark: both backends on the hot path's dominant shapes (useful FLOP = 2*M*N*K)."""
import json
import sys
import os
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from rba_b200 import ops

dev = torch.device("cuda", 0)
shapes = [  # (name, M, N, K)
    ("s0_qkv", 8 * 136224, 384, 128), ("s0_fc1", 8 * 131072, 512, 128), ("s0_fc2", 8 * 131072, 128, 512),
    ("s2_qkv", 8 * 9504, 1536, 512), ("s2_fc1", 8 * 8192, 2048, 512), ("s2_fc2", 8 * 8192, 512, 2048),
    ("s3_fc1", 8 * 2048, 4096, 1024), ("mask_einsum", 100, 131072, 256),
]
out = []
for name, M, N, K in shapes:
    a = torch.randn(M, K, device=dev)
    w = torch.randn(N, K, device=dev) / K ** 0.5
    ap, wp = ops.split_planes(a), ops.split_planes(w)
    c = torch.empty(M, N, device=dev)
    rec = {"name": name, "M": M, "N": N, "K": K}
    for be, bn in ((ops.RBA_GEMM_FFMA, "ffma"), (ops.RBA_GEMM_TC, "tc")):
        try:
            for _ in range(2):
                ops.gemm(ap, wp, out=c, backend=be)
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            reps = 5
            e0.record()
            for _ in range(reps):
                ops.gemm(ap, wp, out=c, backend=be)
            e1.record()
            torch.cuda.synchronize()
            ms = e0.elapsed_time(e1) / reps
            rec[bn + "_ms"] = ms
            rec[bn + "_tflops"] = 2.0 * M * N * K / ms / 1e9
            rec[bn + "_GBs"] = (4.0 * M * K + 4.0 * N * K + 4.0 * M * N) / ms / 1e6
        except Exception as ex:
            rec[bn + "_error"] = str(ex)[:200]
    if "tc_ms" in rec and "ffma_ms" in rec:
        c1 = ops.gemm(ap, wp, backend=ops.RBA_GEMM_FFMA)
        c2 = ops.gemm(ap, wp, backend=ops.RBA_GEMM_TC)
        rec["max_abs_diff"] = float((c1 - c2).abs().max())
    print(json.dumps(rec), flush=True)
    del a, w, ap, wp, c
