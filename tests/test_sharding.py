"""Multi-process host logic (world_size 2, gloo, CPU): view sharding and the result gather."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from diffmvs_b200 import sharding


def test_shard_views_partitions_contiguously():
    for n in (1, 7, 49, 64):
        for world in (1, 2, 4, 8):
            seen = []
            for r in range(world):
                blk = sharding.shard_views(n, r, world)
                seen.extend(blk)
                assert len(blk) in (n // world, n // world + 1)
            assert seen == list(range(n))
    assert sharding.owner_of(48, 49, 8) == 7
    with pytest.raises(ValueError):
        sharding.shard_views(4, 2, 2)


def _worker(rank, world, port, num_views, ret):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    mine = sharding.shard_views(num_views, rank, world)
    # a rank's "depth map" for view v is the constant v (stand-in for the forward; no GPU here)
    local = torch.stack([torch.full((4, 6), float(v)) for v in mine]) if len(mine) else torch.zeros(0, 4, 6)
    counts = [len(sharding.shard_views(num_views, r, world)) for r in range(world)]
    out = sharding.gather_maps(local, counts, dst=0)
    if rank == 0:
        ret["shape"] = tuple(out.shape)
        ret["vals"] = out[:, 0, 0].tolist()
    else:
        assert out is None
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("num_views", [5, 8])
def test_gather_maps_two_ranks_gloo(num_views):
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    ret = mp.Manager().dict()
    mp.spawn(_worker, args=(2, port, num_views, ret), nprocs=2, join=True)
    assert ret["shape"] == (num_views, 4, 6)
    assert ret["vals"] == [float(v) for v in range(num_views)]


def _fusion_worker(rank, world, port, ret):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    num_views = 5
    mine = sharding.shard_views(num_views, rank, world)
    counts = [len(sharding.shard_views(num_views, r, world)) for r in range(world)]
    local = torch.stack([torch.full((3, 4), float(v)) for v in mine])
    every = sharding.all_gather_maps(local, counts)              # each rank now sees all depth maps
    assert every[:, 0, 0].tolist() == [float(v) for v in range(num_views)]
    # rank r fuses 2*r + 1 points (rank 1 more than rank 0: ragged), values encode (rank, index)
    n = 2 * rank + 1
    pts = torch.arange(n * 3, dtype=torch.float32).view(n, 3) + 100.0 * rank + 0.25
    col = (torch.arange(n * 3).view(n, 3) % 251 + rank).to(torch.uint8)
    p, c = sharding.gather_points(pts, col, dst=0)
    if rank == 0:
        ret["pts"], ret["col"] = p.tolist(), c.tolist()
    else:
        assert p is None and c is None
    # an empty contribution is legal
    p, c = sharding.gather_points(pts[:0] if rank == 0 else pts, col[:0] if rank == 0 else col, dst=0)
    if rank == 0:
        ret["n2"] = p.shape[0]
    dist.barrier()
    dist.destroy_process_group()


def test_fused_point_cloud_gather_two_ranks_gloo():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    ret = mp.Manager().dict()
    mp.spawn(_fusion_worker, args=(2, port, ret), nprocs=2, join=True)
    exp_pts, exp_col = [], []
    for rank in range(2):
        n = 2 * rank + 1
        exp_pts += (torch.arange(n * 3, dtype=torch.float32).view(n, 3) + 100.0 * rank + 0.25).tolist()
        exp_col += (torch.arange(n * 3).view(n, 3) % 251 + rank).tolist()
    assert ret["pts"] == exp_pts and ret["col"] == exp_col
    assert ret["n2"] == 3
