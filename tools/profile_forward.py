"""Per-kernel device time of one eager forward under real (warm, concurrent-clock) conditions via torch.profiler (CUPTI)."""
import argparse
import os
import sys
from collections import defaultdict

import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import rba_b200
from rba_b200 import weights

ap = argparse.ArgumentParser()
ap.add_argument("--batch", type=int, default=8)
ap.add_argument("--height", type=int, default=1024)
ap.add_argument("--width", type=int, default=2048)
ap.add_argument("--backend", default="tc")
ap.add_argument("--model", default="swin_b_1dl", choices=["swin_b_1dl", "swin_l_1dl", "swin_b_full"])
a = ap.parse_args()
dev = torch.device("cuda", 0)
mc = getattr(rba_b200.config, a.model)()
eng = rba_b200.Engine(mc, 0).load_state_dict(weights.init_state_dict(mc, seed=0))
eng.set_gemm_backend(a.backend)
img = torch.randint(0, 256, (a.batch, 3, a.height, a.width), dtype=torch.uint8, device=dev)
out = eng.alloc_outputs(a.batch, a.height, a.width, rba=True)
for _ in range(2):
    eng.forward_into(img, out)
torch.cuda.synchronize()
from torch.profiler import profile, ProfilerActivity
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    eng.forward_into(img, out)
    torch.cuda.synchronize()
tot = defaultdict(lambda: [0, 0.0])
seq = []
for ev in prof.events():
    if ev.device_type == torch.autograd.DeviceType.CUDA:
        name = ev.name.split("(")[0][:80]
        seq.append((ev.time_range.start, name, ev.device_time_total if hasattr(ev, "device_time_total") else ev.cuda_time_total))
        tot[name][0] += 1
        tot[name][1] += ev.device_time_total if hasattr(ev, "device_time_total") else ev.cuda_time_total
total = sum(v[1] for v in tot.values())
print(f"batch {a.batch} {a.height}x{a.width} backend {a.backend}: total kernel time {total/1e3:.2f} ms = {total/1e3/a.batch:.2f} ms/img")
for name, (n, us) in sorted(tot.items(), key=lambda kv: -kv[1][1])[:30]:
    print(f"{name:80s} {n:6d} {us/1e3:10.3f} ms {100*us/total:6.2f}%")

if os.environ.get("RBA_PROFILE_SEQ"):
    print("--- launch sequence (us) ---")
    for i, (t0, name, us) in enumerate(sorted(seq)):
        print(f"{i:4d} {us:10.1f} {name[:60]}")
