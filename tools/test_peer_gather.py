"""2+ GPU check of rba_b200.parallel.PeerGather (copy-engine gather of score maps to rank 0 through symmetric memory):
torchrun --nproc-per-node N tools/test_peer_gather.py  -> every step's (world, n, H, W) tensor on rank 0 equals what the ranks sent."""
import os
import sys

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from rba_b200.parallel import PeerGather  # noqa: E402

local = int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
rank, world = dist.get_rank(), dist.get_world_size()
shape = (2, 64, 96)
pg = PeerGather(shape, dev)
ok = True
for step in range(7):
    x = torch.full(shape, float(100 * step + rank), device=dev) + torch.arange(shape[-1], device=dev)
    out = pg.submit(x)
    x.zero_()                                   # the producer may overwrite its maps right away
    pg.wait()
    if rank == 0:
        torch.cuda.synchronize()
        for r in range(world):
            want = torch.full(shape, float(100 * step + r), device=dev) + torch.arange(shape[-1], device=dev)
            ok = ok and bool(torch.equal(out[r], want))
dist.barrier()
if rank == 0:
    print("PeerGather", "OK" if ok else "MISMATCH", "world", world)
dist.destroy_process_group()
