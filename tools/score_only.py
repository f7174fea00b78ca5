import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from rba_b200 import ops
dev = torch.device("cuda", 0)
B = int(sys.argv[1]) if len(sys.argv) > 1 else 2
g = torch.Generator().manual_seed(2)
masks = (torch.randn(B, 100, 256, 512, generator=g) * 0.99 - 0.54).to(dev)
logits = torch.randn(B, 100, 20, generator=g).to(dev)
for _ in range(3):
    r = ops.score_fused(masks, logits, (1024, 2048))
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(5):
    r = ops.score_fused(masks, logits, (1024, 2048))
e1.record(); torch.cuda.synchronize()
print("score ms/launch", e0.elapsed_time(e1) / 5, "B", B)
