"""CPU, world_size 2 over gloo: the sharding + single all-gather logic of the N>1 path (rba_b200/parallel.py)."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from rba_b200 import parallel


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _fake_score(images):
    # deterministic stand-in for the engine: a function of the image only (images are independent units)
    return torch.stack([im.float().mean(0) * 0.5 - 1.0 for im in images])


def _worker(rank, world, port, n_items, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    g = torch.Generator().manual_seed(0)
    images = [torch.randint(0, 256, (3, 6, 10), dtype=torch.uint8, generator=g) for _ in range(n_items)]
    out = parallel.score_sharded(_fake_score, images, batch=2)
    q.put((rank, out))
    dist.barrier()
    dist.destroy_process_group()


def _worker_og(rank, world, port, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    og = parallel.OverlappedGather((2, 3, 4), torch.device("cpu"))
    outs = []
    for step in range(3):                                # three submissions through the two staging buffers
        local = torch.full((2, 3, 4), float(10 * step + rank))
        outs.append(og.submit(local).clone())
    og.wait()
    q.put((rank, torch.stack(outs)))
    dist.barrier()
    dist.destroy_process_group()


def test_overlapped_gather_gloo():
    world, port = 2, _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    ps = [ctx.Process(target=_worker_og, args=(r, world, port, q)) for r in range(world)]
    for p in ps:
        p.start()
    res = dict(q.get(timeout=120) for _ in range(world))
    for p in ps:
        p.join(timeout=60)
        assert p.exitcode == 0
    for rank in (0, 1):
        for step in range(3):
            want = torch.cat([torch.full((2, 3, 4), float(10 * step + r)) for r in range(world)])
            assert torch.equal(res[rank][step], want)


def _run(n_items):
    world, port = 2, _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    ps = [ctx.Process(target=_worker, args=(r, world, port, n_items, q)) for r in range(world)]
    for p in ps:
        p.start()
    res = dict(q.get(timeout=120) for _ in range(world))
    for p in ps:
        p.join(timeout=60)
        assert p.exitcode == 0
    return res


def test_sharded_equals_single_process_bitwise():
    for n in (5, 4, 1):
        g = torch.Generator().manual_seed(0)
        images = [torch.randint(0, 256, (3, 6, 10), dtype=torch.uint8, generator=g) for _ in range(n)]
        single = _fake_score(images)
        res = _run(n)
        for rank in (0, 1):
            assert torch.equal(res[rank], single), (n, rank)


def test_shard_indices_round_robin():
    assert parallel.shard_indices(10, 0, 8) == [0, 8]
    assert parallel.shard_indices(10, 7, 8) == [7]
    assert sorted(sum((parallel.shard_indices(64, r, 8) for r in range(8)), [])) == list(range(64))
    out = parallel.gather_scores(torch.ones(3, 2, 2), 3, rank=0, world=1)
    assert out.shape == (3, 2, 2)
