"""Three eager forwards of Swin-B 1dl (B images of 1024x2048) for ncu captures of individual kernels:
   ncu --set full -k regex:<kernel> --launch-skip <n> -c 1 ... python tools/forward_once.py [B] [model]"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import rba_b200  # noqa: E402
from rba_b200 import weights  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 2
model = sys.argv[2] if len(sys.argv) > 2 else "swin_b_1dl"
mc = getattr(rba_b200.config, model)()
eng = rba_b200.Engine(mc, 0).load_state_dict(weights.init_state_dict(mc, seed=0))
eng.set_gemm_backend("tc")
img = torch.randint(0, 256, (B, 3, 1024, 2048), dtype=torch.uint8, device="cuda")
out = eng.alloc_outputs(B, 1024, 2048, rba=True)
for _ in range(3):
    eng.forward_into(img, out)
torch.cuda.synchronize()
print("ok", float(out["rba"].mean()))
