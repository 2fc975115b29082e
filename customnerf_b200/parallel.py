"""Ray-sharded data parallelism: one process per GPU, a full replica of the hash table /
MLPs / occupancy grid on every rank, rays split across ranks, and ONE all-reduce(sum) per step over a single flat
buffer holding every gradient (hash table + the three MLPs).  Nothing else is exchanged.

The reference's own DDP wrap is dead code (nerf/utils_init_nerf.py:76-78; init_process_group is never called).
"""
import os

import torch
import torch.distributed as dist


def init_from_env(backend=None):
    """(rank, local_rank, world_size); initialises torch.distributed when launched by torchrun."""
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1 and not dist.is_initialized():
        backend = backend or ("nccl" if torch.cuda.is_available() else "gloo")
        if backend == "nccl":
            torch.cuda.set_device(local_rank)
        dist.init_process_group(backend=backend)
    return rank, local_rank, world


def shard_rays(n_rays, rank, world_size, mode="interleaved"):
    """Indices of the rays owned by ``rank``.

    interleaved (default): ray i -> rank i % world.  Neighbouring pixels have similar sample counts, so a strided
    split balances the per-rank sample totals far better than contiguous image rows (SURVEY.md 8(e) caveat)."""
    if mode == "interleaved":
        return torch.arange(min(rank, n_rays), n_rays, world_size)
    per = (n_rays + world_size - 1) // world_size
    return torch.arange(min(n_rays, rank * per), min(n_rays, (rank + 1) * per))


class FlatGradSync:
    """All-reduce every gradient as one flat fp32 buffer (one collective per step)."""

    def __init__(self, params, group=None):
        self.params = list(params)
        self.group = group
        self.numel = sum(p.numel() for p in self.params)
        p0 = self.params[0]
        self.flat = torch.zeros(self.numel, dtype=torch.float32, device=p0.device)
        self.views, off = [], 0
        for p in self.params:
            self.views.append(self.flat[off:off + p.numel()].view_as(p))
            off += p.numel()

    def __call__(self, params=None):
        for p, v in zip(self.params, self.views):
            if p.grad is None:
                v.zero_()
            elif p.grad.data_ptr() != v.data_ptr():
                v.copy_(p.grad)
        if dist.is_initialized() and dist.get_world_size(self.group) > 1:
            dist.all_reduce(self.flat, op=dist.ReduceOp.SUM, group=self.group)
        for p, v in zip(self.params, self.views):
            p.grad = v            # gradients now alias the flat buffer: later steps accumulate straight into it
