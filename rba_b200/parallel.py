"""Data-parallel scoring over the GPUs of one box.

Images are independent units: weights are replicated, image i goes to rank i % world, and the ONLY exchange on
the path is one all-gather of the per-image score maps (what evaluate_ood's OODEvaluator consumes on rank 0,
support.py:353-399).  One process per GPU (torchrun), NCCL for CUDA tensors; the same code runs over gloo on CPU
tensors so the host logic is testable without GPUs."""
import torch
import torch.distributed as dist


def shard_indices(n_items, rank, world):
    """Round-robin: item i -> rank i % world (SURVEY §8e 'image i -> rank i mod 8')."""
    return list(range(rank, n_items, world))


def gather_scores(local_scores, n_items, rank=None, world=None, group=None):
    """local_scores: (n_local, H, W) tensor for items shard_indices(n_items, rank, world), in that order.
    Returns the (n_items, H, W) tensor in the ORIGINAL item order on every rank.  One all_gather; ranks with one
    item fewer pad with a zero map that is dropped on reassembly."""
    if world is None:
        world = dist.get_world_size(group) if dist.is_initialized() else 1
    if rank is None:
        rank = dist.get_rank(group) if dist.is_initialized() else 0
    per = -(-n_items // world)                       # ceil
    H, W = local_scores.shape[-2:]
    buf = local_scores.new_zeros((per, H, W))
    n_local = len(shard_indices(n_items, rank, world))
    if local_scores.shape[0] != n_local:
        raise ValueError(f"rank {rank} holds {local_scores.shape[0]} maps, expected {n_local}")
    buf[:n_local].copy_(local_scores)
    if world == 1:
        return buf[:n_items]
    out = local_scores.new_empty((world * per, H, W))
    dist.all_gather_into_tensor(out, buf, group=group)
    out = out.view(world, per, H, W)
    order = torch.empty((n_items, H, W), dtype=out.dtype, device=out.device)
    for r in range(world):
        idx = shard_indices(n_items, r, world)
        if idx:
            order[torch.as_tensor(idx, device=out.device)] = out[r, :len(idx)]
    return order


def score_sharded(score_fn, images, rank=None, world=None, group=None, batch=8):
    """images: sequence of (3,H,W) tensors (same size).  score_fn(list_of_images) -> (n,H,W) score maps
    (e.g. rba_b200.MaskFormer.rba on this rank's GPU).  Returns all score maps in input order on every rank."""
    if world is None:
        world = dist.get_world_size(group) if dist.is_initialized() else 1
    if rank is None:
        rank = dist.get_rank(group) if dist.is_initialized() else 0
    mine = shard_indices(len(images), rank, world)
    outs = []
    for i in range(0, len(mine), batch):
        outs.append(score_fn([images[j] for j in mine[i:i + batch]]))
    if outs:
        local = torch.cat(outs)
    else:
        ref = images[0]
        local = torch.zeros((0, ref.shape[-2], ref.shape[-1]), dtype=torch.float32, device=ref.device)
    return gather_scores(local, len(images), rank, world, group)


class OverlappedGather:
    """The path's one collective, taken off the critical path: the all-gather of step i's score maps runs on a side stream
    while step i+1's forward runs on the compute stream.  `submit(local)` snapshots the (n_local, H, W) maps into one of two
    staging buffers (a 20 us device copy, so the producer may overwrite its output at once) and launches
    `all_gather_into_tensor` on the side stream; `wait()` makes the current stream wait for everything submitted;
    `submit` returns the (world * n_local, H, W) result tensor of that submission (valid after `wait()`, until the
    submission after next reuses it)."""

    def __init__(self, shape, device, dtype=torch.float32, group=None):
        self.group = group
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.stage = [torch.empty(shape, dtype=dtype, device=device) for _ in range(2)]
        self.out = [torch.empty((self.world * shape[0],) + tuple(shape[1:]), dtype=dtype, device=device) for _ in range(2)]
        self.stream = torch.cuda.Stream(device) if device.type == "cuda" else None
        self.ev_ready = [torch.cuda.Event(), torch.cuda.Event()] if self.stream else None
        self.ev_done = [torch.cuda.Event(), torch.cuda.Event()] if self.stream else None
        self.n = 0

    def submit(self, local):
        k = self.n & 1
        if self.stream is None:                       # CPU tensors (gloo tests): synchronous
            self.stage[k].copy_(local)
            if self.world > 1:
                dist.all_gather_into_tensor(self.out[k], self.stage[k], group=self.group)
            else:
                self.out[k].copy_(self.stage[k])
            self.n += 1
            return self.out[k]
        comp = torch.cuda.current_stream(local.device)
        if self.n >= 2:
            comp.wait_event(self.ev_done[k])          # the all-gather that last read stage[k]
        self.stage[k].copy_(local, non_blocking=True)
        self.ev_ready[k].record(comp)
        with torch.cuda.stream(self.stream):
            self.stream.wait_event(self.ev_ready[k])
            if self.world > 1:
                dist.all_gather_into_tensor(self.out[k], self.stage[k], group=self.group)
            else:
                self.out[k].copy_(self.stage[k], non_blocking=True)
            self.ev_done[k].record(self.stream)
        self.n += 1
        return self.out[k]

    def wait(self):
        if self.stream is not None:
            torch.cuda.current_stream(self.stage[0].device).wait_stream(self.stream)


class PeerGather:
    """Gather of the per-step score maps to ONE rank (what evaluate_ood's OODEvaluator consumes, support.py:353-399) without
    a collective kernel: every rank pushes its (n_local, H, W) maps straight into the root's buffer over NVLink with a
    copy-engine peer copy (the buffer is CUDA symmetric memory: torch.distributed._symmetric_memory, peer-mapped on every
    rank), on a side stream under the next step's forward.  No SM is taken from the forward (the NCCL all-gather kernel of
    `OverlappedGather` competes with the persistent one-CTA-per-SM kernels of the forward for SMs), and 1/world of the
    all-gather's bytes move.  Two slots; arrival and slot-free hand-shakes are symmetric-memory signals (stream ordered).

        pg = PeerGather((n_local, H, W), device)      # collective: every rank constructs it
        out = pg.submit(local)                         # every rank, every step; root gets (world, n_local, H, W) or None
        pg.wait()                                      # current stream waits for this rank's part (root: for all arrivals)

    `out` on the root is valid after `wait()` until the submission after next."""

    def __init__(self, shape, device, dtype=torch.float32, group=None, root=0):
        import torch.distributed._symmetric_memory as symm
        self.group = group if group is not None else dist.group.WORLD
        self.world, self.rank, self.root = dist.get_world_size(self.group), dist.get_rank(self.group), root
        self.shape = tuple(shape)
        full = (2, self.world) + self.shape
        self.buf = symm.empty(full, dtype=dtype, device=device)
        self.hdl = symm.rendezvous(self.buf, self.group)
        self.root_buf = self.buf if self.rank == root else self.hdl.get_buffer(root, full, dtype)
        self.stream = torch.cuda.Stream(device)
        self.stage = [torch.empty(self.shape, dtype=dtype, device=device) for _ in range(2)]   # snapshots: the producer may
        self.ev_ready = [torch.cuda.Event(), torch.cuda.Event()]                               # overwrite `local` at once
        self.ev_done = [torch.cuda.Event(), torch.cuda.Event()]
        self.n = 0
        self.hdl.barrier()

    def submit(self, local):
        k = self.n & 1
        comp = torch.cuda.current_stream(local.device)
        if self.n >= 2:
            comp.wait_event(self.ev_done[k])               # the peer copy that last read stage[k]
        self.stage[k].copy_(local, non_blocking=True)      # ~20 us on-device snapshot
        self.ev_ready[k].record(comp)                      # earlier results have been consumed by now (stream order)
        with torch.cuda.stream(self.stream):
            self.stream.wait_event(self.ev_ready[k])
            if self.rank == self.root:
                if self.n >= 2:                            # slot k (step n - 2) has been consumed: peers may overwrite it
                    for r in range(self.world):
                        if r != self.root:
                            self.hdl.put_signal(r, channel=k)
                self.root_buf[k, self.rank].copy_(self.stage[k], non_blocking=True)
                for r in range(self.world):                # arrivals of this step
                    if r != self.root:
                        self.hdl.wait_signal(r, channel=2 + k)
            else:
                if self.n >= 2:
                    self.hdl.wait_signal(self.root, channel=k)
                self.root_buf[k, self.rank].copy_(self.stage[k], non_blocking=True)    # copy-engine peer write over NVLink
                self.hdl.put_signal(self.root, channel=2 + k)
            self.ev_done[k].record(self.stream)
        self.n += 1
        return self.root_buf[k] if self.rank == self.root else None

    def wait(self):
        torch.cuda.current_stream(self.buf.device).wait_stream(self.stream)
