"""Reference-view sharding across GPUs (one process per GPU) and the single result gather.

The reference processes one reference view per forward at batch 1 (`/root/reference/test.py:101-104,115`);
views are independent, so ranks take contiguous blocks of the view list and no collective sits on the data
path.  The only exchange is the gather of finished depth / confidence maps to rank 0.
Works with any `torch.distributed` backend (NCCL on the GPU box, gloo in the CPU tests).
"""
from __future__ import annotations

from typing import List, Optional, Sequence

import torch
import torch.distributed as dist


def shard_views(num_views: int, rank: int, world: int) -> range:
    """Contiguous block of reference-view indices owned by `rank` (sizes differ by at most one).
    Neighbouring reference views share source images, so contiguous blocks keep a rank's working set local."""
    if not (0 <= rank < world):
        raise ValueError(f"rank {rank} outside world of {world}")
    base, extra = divmod(num_views, world)
    start = rank * base + min(rank, extra)
    return range(start, start + base + (1 if rank < extra else 0))


def owner_of(view: int, num_views: int, world: int) -> int:
    for r in range(world):
        if view in shard_views(num_views, r, world):
            return r
    raise ValueError(view)


def gather_maps(local: torch.Tensor, counts: Sequence[int], dst: int = 0) -> Optional[torch.Tensor]:
    """Gather per-rank stacks `[n_r, H, W]` to `dst` as one `[sum(n_r), H, W]` tensor in view order.
    Ranks pad to the largest block so a single fixed-size gather suffices."""
    world, rank = dist.get_world_size(), dist.get_rank()
    n_max = max(counts)
    H, W = local.shape[-2:]
    buf = local.new_zeros((n_max, H, W))
    buf[:local.shape[0]] = local
    out: Optional[List[torch.Tensor]] = [torch.empty_like(buf) for _ in range(world)] if rank == dst else None
    dist.gather(buf, out, dst=dst)
    if rank != dst:
        return None
    return torch.cat([o[:c] for o, c in zip(out, counts)], dim=0)
