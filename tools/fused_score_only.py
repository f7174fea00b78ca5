"""Launches the fused mask-einsum + score kernel alone (for ncu): python tools/fused_score_only.py [B] [reps]"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from rba_b200 import ops  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 2
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 5
dev = torch.device("cuda", 0)
Q, K, D, h, w = 100, 19, 256, 256, 512
g = torch.Generator(device=dev).manual_seed(2)
feat = torch.randn(B * h * w, D, device=dev, generator=g)
emb = torch.randn(B * Q, D, device=dev, generator=g) * (0.99 / D ** 0.5)
f_pl = tuple(t.view(B, h, w, D) for t in ops.split_planes(feat))
e_pl = tuple(t.view(B, Q, D) for t in ops.split_planes(emb))
bias = torch.full((B, Q), -0.54, device=dev)
logits = torch.randn(B, Q, K + 1, device=dev, generator=g)
for _ in range(reps):
    r = ops.einsum_score_fused(e_pl, f_pl, logits, (4 * h, 4 * w), bias=bias)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(reps):
    r = ops.einsum_score_fused(e_pl, f_pl, logits, (4 * h, 4 * w), bias=bias)
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / reps
print(f"B={B}: {ms:.3f} ms/launch = {ms / B * 1e3:.1f} us/img; rba mean {r.mean().item():.4f}")
