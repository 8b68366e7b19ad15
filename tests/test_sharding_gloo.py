"""N > 1 host logic on CPU: world_size-2 gloo process group; frames are sharded round-robin with no
data-path collective, throughput = all units / max elapsed."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from popsift_b200 import shard


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close(); return p


def _worker(rank, world, port, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    mine = shard.frame_indices(13, rank, world)
    # pretend: every frame is 10 units and costs (rank+1) ms
    units, ms = shard.aggregate(10.0 * len(mine), (rank + 1) * len(mine))
    gathered = [None] * world
    dist.all_gather_object(gathered, mine)
    q.put((rank, mine, units, ms, gathered))
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_sharding_covers_all_frames_once():
    world, port = 2, _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    ps = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in ps:
        p.start()
    res = [q.get(timeout=120) for _ in range(world)]
    for p in ps:
        p.join(timeout=60)
        assert p.exitcode == 0
    res.sort()
    all_frames = sorted(res[0][1] + res[1][1])
    assert all_frames == list(range(13))                 # every frame exactly once
    assert res[0][1] == list(range(0, 13, 2)) and res[1][1] == list(range(1, 13, 2))
    for r in res:
        assert r[2] == 130.0                             # SUM of units
        assert r[3] == max(1 * 7, 2 * 6)                 # MAX of elapsed
        assert r[4] == [res[0][1], res[1][1]]


def test_single_process_is_identity():
    assert shard.frame_indices(5, 0, 1) == [0, 1, 2, 3, 4]
    assert shard.aggregate(3.0, 4.0) == (3.0, 4.0)
    assert [shard.slot_of(i, 4) for i in range(6)] == [0, 1, 2, 3, 0, 1]
