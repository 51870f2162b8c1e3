"""N > 1 host logic on CPU: world_size-2 gloo process group, frame sharding and the max-over-ranks
timing reduction bench.py uses (the data path itself has no collective)."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from animal_vision_b200 import sharding

SPECIES = ("Dog", "Cat", "HoneyBee")


def test_shards_partition_the_frame_range():
    for total in (0, 1, 7, 60, 61, 480):
        for world in (1, 2, 3, 4, 8):
            seen = []
            for r in range(world):
                b, e = sharding.shard_range(total, r, world)
                assert 0 <= b <= e <= total
                seen += list(range(b, e))
            assert seen == list(range(total)), (total, world)
            sizes = [sharding.shard_range(total, r, world) for r in range(world)]
            assert max(e - b for b, e in sizes) - min(e - b for b, e in sizes) <= 1
    with pytest.raises(ValueError):
        sharding.shard_range(10, 2, 2)


def test_species_round_robin_is_global():
    plan = [sharding.shard_plan(120, r, 2, SPECIES) for r in range(2)]
    flat = plan[0] + plan[1]
    assert [i for i, _ in flat] == list(range(120))
    assert all(sp == SPECIES[i % 3] for i, sp in flat)


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        total = 121
        b, e = sharding.shard_range(total, rank, world)
        # every rank reports its frame indices; gathered they must tile the range exactly once
        mine = torch.zeros(total, dtype=torch.int32)
        mine[b:e] = 1
        dist.all_reduce(mine)
        ok_cover = bool((mine == 1).all())
        t = sharding.max_over_ranks(10.0 + rank)          # slowest rank defines the step time
        q.put((rank, ok_cover, t, e - b))
    finally:
        dist.destroy_process_group()


def test_world_size_2_gloo():
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in range(world))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert [r[0] for r in res] == [0, 1]
    assert all(r[1] for r in res)
    assert all(r[2] == 11.0 for r in res)
    assert sum(r[3] for r in res) == 121
