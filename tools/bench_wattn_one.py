"""One stage shape of the tcgen05 window attention, for ablations / ncu: python tools/bench_wattn_one.py [stage 0-3] [B] [reps]"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from rba_b200 import ops  # noqa: E402

stage = int(sys.argv[1]) if len(sys.argv) > 1 else 2
B = int(sys.argv[2]) if len(sys.argv) > 2 else 8
reps = int(sys.argv[3]) if len(sys.argv) > 3 else 10
H, W, C, heads = [(256, 512, 128, 4), (128, 256, 256, 8), (64, 128, 512, 16), (32, 64, 1024, 32)][stage]
dev = torch.device("cuda", 0)
nW = -(-H // 12) * -(-W // 12)
g = torch.Generator(device=dev).manual_seed(1)
qp = ops.split_planes(torch.randn(B * nW * 144, 3 * C, device=dev, generator=g))
table = torch.randn(529, heads, device=dev, generator=g)
tiled = os.environ.get("RBA_WT_ROWMAJOR") is None
if tiled:
    qp = ops.qkv_to_tiles(qp, heads)
for _ in range(3):
    ops.window_attn_tc(qp, table, B, H, W, C, heads, 12, 0, tiled=tiled)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(reps):
    ops.window_attn_tc(qp, table, B, H, W, C, heads, 12, 0, tiled=tiled)
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / reps
items = B * nW * heads
print(f"stage {stage} B={B}: {ms:.4f} ms, {items} items, {ms * 1e3 / (items / 148):.2f} us per item per SM")
