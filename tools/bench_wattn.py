"""Window attention alone at the Swin-B stage shapes of a 1024x2048 image (B images): tcgen05 + TMA kernel vs the mma.sync
kernel, ms per launch (CUDA events) and the total over the 24 blocks of the backbone.  python tools/bench_wattn.py [B]"""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from rba_b200 import ops  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 8
dev = torch.device("cuda", 0)
stages = [(256, 512, 128, 4, 2), (128, 256, 256, 8, 2), (64, 128, 512, 16, 18), (32, 64, 1024, 32, 2)]   # H, W, C, heads, blocks
res = {"B": B, "stages": []}
tot = {"tcgen05": 0.0, "tcgen05_rowmajor": 0.0, "mma_sync": 0.0}
for H, W, C, heads, blocks in stages:
    nW = -(-H // 12) * -(-W // 12)
    rows = B * nW * 144
    g = torch.Generator(device=dev).manual_seed(1)
    qkv = torch.randn(rows, 3 * C, device=dev, generator=g)
    qp = ops.split_planes(qkv)
    table = torch.randn(529, heads, device=dev, generator=g)
    row = {"H": H, "W": W, "C": C, "heads": heads, "blocks": blocks, "items": B * nW * heads,
           "algorithmic_bytes": rows * C * 4 * 4}            # q, k, v in + o out, 4 B per element (hi + lo planes)
    outs = {}
    qt = ops.qkv_to_tiles(qp, heads)
    tiled_fn = lambda q, *a: ops.window_attn_tc(qt, *a, tiled=True)  # noqa: E731
    for name, fn in (("tcgen05", tiled_fn), ("tcgen05_rowmajor", ops.window_attn_tc), ("mma_sync", ops.window_attn_planes)):
        for shift in (0, 6):
            for _ in range(3):
                o = fn(qp, table, B, H, W, C, heads, 12, shift)
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(10):
                o = fn(qp, table, B, H, W, C, heads, 12, shift)
            e1.record()
            torch.cuda.synchronize()
            ms = e0.elapsed_time(e1) / 10
            row[f"{name}_shift{shift}_ms"] = ms
            tot[name] += ms * blocks / 2
            outs[(name, shift)] = o
    d = max((outs[("tcgen05", s)][0].float() + outs[("tcgen05", s)][1].float() - outs[("mma_sync", s)][0].float() - outs[("mma_sync", s)][1].float()).abs().max().item() for s in (0, 6))
    row["max_abs_diff_between_kernels"] = d
    row["tcgen05_GBs"] = row["algorithmic_bytes"] / (row["tcgen05_shift0_ms"] * 1e-3) / 1e9
    res["stages"].append(row)
    print(row)
res["total_ms_24_blocks"] = tot
print(tot)
os.makedirs("gpurun_out", exist_ok=True)
json.dump(res, open("gpurun_out/bench_wattn.json", "w"), indent=1)
