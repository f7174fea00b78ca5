"""Informational baseline: the SAME PyTorch statement of the reference forward (oracle/rba_oracle.py = the reference's
own torch ops, op for op) executed by PyTorch/cuBLAS/cuDNN on the B200 in fp32 with default torch flags — a stand-in
for "the reference PyTorch/Detectron2 GPU path", which cannot travel to the GPU box.  Not part of the product, not
used by bench.py; prints images/s for comparison with bench.py's value."""
import os
import sys
import time
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "oracle"))
import rba_oracle as O
import rba_b200
from rba_b200 import weights

B = int(sys.argv[1]) if len(sys.argv) > 1 else 2
dev = torch.device("cuda", 0)
mc = rba_b200.config.swin_b_1dl()
sd = {k: v.to(dev) for k, v in weights.init_state_dict(mc, seed=0).items()}
g = torch.Generator().manual_seed(1)
imgs = [torch.randint(0, 256, (3, 1024, 2048), dtype=torch.uint8, generator=g).to(dev) for _ in range(B)]
with torch.no_grad():
    for _ in range(2):
        out = O.forward(sd, mc, imgs)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    n = 3
    for _ in range(n):
        out = O.forward(sd, mc, imgs)
    torch.cuda.synchronize()
    dt = (time.perf_counter() - t0) / n
print(f"torch-GPU restatement of the reference: batch {B} -> {dt*1e3:.1f} ms/step = {B/dt:.2f} img/s "
      f"(fp32, cudnn.allow_tf32={torch.backends.cudnn.allow_tf32}, matmul.allow_tf32={torch.backends.cuda.matmul.allow_tf32}), "
      f"peak mem {torch.cuda.max_memory_allocated()/2**30:.1f} GiB")
