"""CPU: the multi-rank host logic (env sharding, the one stats reduction) with gloo, world_size 2."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from traffic_env_b200.dist import STAT_KEYS, mean_episode_return, reduce_stats, shard_range


def test_shard_range_partitions_every_env_once():
    for total in (1, 7, 16384, 2 ** 20, 1000003):
        for world in (1, 2, 3, 4, 8):
            spans = [shard_range(total, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == total
            for a, b in zip(spans, spans[1:]):
                assert a[1] == b[0]
            sizes = [e - b for b, e in spans]
            assert max(sizes) - min(sizes) <= 1


def _worker(rank, world, port, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    stats = {k: (rank + 1) * (i + 1) for i, k in enumerate(STAT_KEYS)}
    stats["return_sum"] = -1.5 * (rank + 1)
    stats["disc_return_sum"] = -0.25 * (rank + 1)
    out = reduce_stats(stats)
    b, e = shard_range(1001, rank, world)
    t = torch.tensor([float(e - b)], dtype=torch.float64)
    dist.all_reduce(t)
    q.put((rank, out, t.item()))
    dist.destroy_process_group()


def test_reduce_stats_gloo_world2():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for rank, out, total in res:
        assert total == 1001.0
        for i, k in enumerate(STAT_KEYS):
            if k == "return_sum":
                assert out[k] == pytest.approx(-4.5)
            elif k == "disc_return_sum":
                assert out[k] == pytest.approx(-0.75)
            else:
                assert out[k] == 3 * (i + 1)
        assert mean_episode_return(out) == pytest.approx(-0.75 / out["episodes"])


def test_reduce_stats_without_process_group_is_identity():
    stats = {k: i for i, k in enumerate(STAT_KEYS)}
    assert reduce_stats(stats) == stats
