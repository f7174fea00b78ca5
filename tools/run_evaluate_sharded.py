"""BASELINE config 5 shape: OoD evaluation sharded over the GPUs of one box (one process per GPU), Swin-B 1dl or the full
3-level decoder, synthetic 1024 x 2048 images held in memory (so the number is the scoring pipeline, not PNG decode).
   torchrun --nproc-per-node N tools/run_evaluate_sharded.py [--arch swin_b_1dl|swin_b_full] [--images 128]
Every rank scores images rank, rank+N, ...; the histograms are all-reduced once; rank 0 prints img/s and the metrics."""
import argparse
import json
import os
import sys
import time

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import rba_b200  # noqa: E402
from rba_b200 import weights  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--arch", default="swin_b_1dl")
ap.add_argument("--images", type=int, default=128)
ap.add_argument("--out", default="gpurun_out/evaluate_sharded.json")
a = ap.parse_args()
local = int(os.environ.get("LOCAL_RANK", "0"))
world = int(os.environ.get("WORLD_SIZE", "1"))
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
if world > 1:
    fd = os.dup(1); os.dup2(2, 1)
    dist.init_process_group("nccl", device_id=dev)
    t = torch.zeros(1, device=dev); dist.all_reduce(t); torch.cuda.synchronize()
    sys.stdout.flush(); os.dup2(fd, 1); os.close(fd)
rank = dist.get_rank() if world > 1 else 0
mc = getattr(rba_b200.config, a.arch)()
model = rba_b200.MaskFormer(mc)
model.load_state_dict(weights.init_state_dict(mc, seed=0))
model.to(dev).eval()
g = torch.Generator().manual_seed(5)
base = [torch.randint(0, 256, (3, 1024, 2048), dtype=torch.uint8, generator=g) for _ in range(4)]
lab = torch.zeros(1024, 2048, dtype=torch.int64)
lab[300:500, 600:1200] = 1
lab[:64] = 255
items = [(base[i % 4], lab) for i in range(a.images)]        # in-memory dataset: (image CHW uint8, label HW)
ev = rba_b200.OODEvaluator(model)
ev.evaluate_dataset(items[: 16 * world], batch=8, workers=4)                 # warm-up: capture, pinned rings
torch.cuda.synchronize()
if world > 1:
    dist.barrier()
t0 = time.time()
m = ev.evaluate_dataset(items, batch=8, workers=4)
torch.cuda.synchronize()
if world > 1:
    dist.barrier()
dt = time.time() - t0
if rank == 0:
    res = {"arch": a.arch, "n_gpus": world, "images": a.images, "wall_s": dt, "images_per_s": a.images / dt, "metrics": m,
           "note": "in-memory uint8 images -> PinnedBatcher -> ScoreStream (H2D / forward overlap) -> device histogram; one all-reduce of the histogram"}
    os.makedirs(os.path.dirname(a.out), exist_ok=True)
    json.dump(res, open(a.out, "w"), indent=1)
    print(json.dumps(res))
if world > 1:
    dist.destroy_process_group()
