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


def all_gather_maps(local: torch.Tensor, counts: Sequence[int]) -> torch.Tensor:
    """Every rank gets every rank's depth maps `[sum(n_r), H, W]` in view order: geometric filtering of reference view
    i needs the depth maps of its source views, which other ranks may own.  One fixed-size
    all_gather of padded blocks."""
    world = dist.get_world_size()
    n_max = max(counts)
    H, W = local.shape[-2:]
    buf = local.new_zeros((n_max, H, W))
    buf[:local.shape[0]] = local
    out = [torch.empty_like(buf) for _ in range(world)]
    dist.all_gather(out, buf)
    return torch.cat([o[:c] for o, c in zip(out, counts)], dim=0)


def gather_points(points: torch.Tensor, colors: torch.Tensor, dst: int = 0):
    """The single gather of the fused point cloud (north_star): ranks hold different numbers of points
    `[n_r, 3]` float32 + colours `[n_r, 3]` uint8.  Counts are exchanged first, then one padded gather of 15-byte
    records (xyz as raw bytes + rgb) moves everything to `dst`; returns (points, colors) there and (None, None)
    elsewhere.  Order: rank 0's points, then rank 1's, ... (ranks own contiguous view blocks, so this is view order)."""
    world, rank = dist.get_world_size(), dist.get_rank()
    if points.dim() != 2 or points.shape[1] != 3 or tuple(colors.shape) != tuple(points.shape):
        raise ValueError("gather_points: points and colors must both be [N,3]")
    n = torch.tensor([points.shape[0]], dtype=torch.int64, device=points.device)
    all_n = [torch.zeros_like(n) for _ in range(world)]
    dist.all_gather(all_n, n)
    counts = [int(c.item()) for c in all_n]
    n_max = max(max(counts), 1)
    rec = torch.zeros((n_max, 15), dtype=torch.uint8, device=points.device)
    if points.shape[0]:
        rec[:points.shape[0], :12] = points.to(torch.float32).contiguous().view(torch.uint8).view(-1, 12)
        rec[:points.shape[0], 12:] = colors.to(torch.uint8)
    out = [torch.empty_like(rec) for _ in range(world)] if rank == dst else None
    dist.gather(rec, out, dst=dst)
    if rank != dst:
        return None, None
    allrec = torch.cat([o[:c] for o, c in zip(out, counts)], dim=0)
    pts = allrec[:, :12].contiguous().view(torch.float32).view(-1, 3)
    return pts, allrec[:, 12:].contiguous()
