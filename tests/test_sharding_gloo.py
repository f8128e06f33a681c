"""CPU tier: the N>1 path (instance sharding, completion barrier, reductions, result gather) with world_size 2
over gloo -- the same helpers bench.py and the astar CLI use under torchrun with NCCL."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from deepcubea_b200.search import sharding


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close(); return p


def _worker(rank, world, port, n_items, out_q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        r, w, _ = sharding.world()
        mine = sharding.shard_indices(n_items, r, w)
        # stand-in for "solve instance i": a deterministic result per instance
        local = [(i, {"moves": [i % 12, (i * 7) % 12], "nodes": 1000 + i}) for i in mine]
        sharding.completion_barrier()
        merged = sharding.gather_results(local, n_items)
        nodes, secs = sharding.reduce_throughput(sum(x[1]["nodes"] for x in local), 1.0 + rank, torch.device("cpu"))
        # dynamic whole-instance queue: every instance is drawn exactly once, whichever rank gets there first
        import time
        drawn = []
        for i in sharding.InstanceQueue(n_items):
            drawn.append(i)
            time.sleep(0.002 * (rank + 1))
        sharding.completion_barrier()
        merged_q = sharding.gather_results([(i, i * i) for i in drawn], n_items)
        out_q.put((rank, mine, merged, nodes, secs, drawn, merged_q))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("n_items", [7, 8])
def test_two_rank_shard_gather_reduce(n_items):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, n_items, q)) for r in range(2)]
    for p in procs:
        p.start()
    got = {}
    for _ in range(2):
        rank, mine, merged, nodes, secs, drawn, merged_q = q.get(timeout=120)
        got[rank] = (mine, merged, nodes, secs, drawn, merged_q)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert got[0][0] == list(range(0, n_items, 2)) and got[1][0] == list(range(1, n_items, 2))
    merged = got[0][1]
    assert got[1][1] is None                                  # only rank 0 holds the merged list
    assert [m["nodes"] for m in merged] == [1000 + i for i in range(n_items)]     # input order restored
    assert [m["moves"] for m in merged] == [[i % 12, (i * 7) % 12] for i in range(n_items)]
    assert sorted(got[0][4] + got[1][4]) == list(range(n_items))                   # the queue hands out every instance exactly once
    assert got[0][5] == [i * i for i in range(n_items)] and got[1][5] is None
    for r in (0, 1):
        assert got[r][2] == sum(1000 + i for i in range(n_items))                  # SUM over ranks
        assert got[r][3] == 2.0                                                     # MAX over ranks


def test_single_process_paths():
    assert sharding.shard_indices(5, 0, 1) == [0, 1, 2, 3, 4]
    assert list(sharding.InstanceQueue(4)) == [0, 1, 2, 3]
    assert sharding.gather_results([(1, "b"), (0, "a")], 2) == ["a", "b"]
    assert sharding.reduce_throughput(10, 2.5, torch.device("cpu")) == (10.0, 2.5)
    with pytest.raises(ValueError):
        sharding.merge_sharded([[(0, "a")], [(0, "b")]], 2)
    with pytest.raises(ValueError):
        sharding.merge_sharded([[(0, "a")]], 2)
