"""CPU, world_size 2 over gloo: the N>1 plumbing of bench.py (round-robin sharding, max-over-ranks timing)."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from mvstereovision3_b200 import shard


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    mine = shard.frames_for_rank(11, rank, world)
    t, u = shard.reduce_max_and_sum(dist, torch.device("cpu"), 10.0 + 5 * rank, len(mine))
    q.put((rank, mine, t, u))
    dist.barrier()
    dist.destroy_process_group()


def test_round_robin_two_ranks():
    assert shard.frames_for_rank(5, 0, 1) == [0, 1, 2, 3, 4]
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    ps = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in ps:
        p.start()
    res = sorted(q.get(timeout=120) for _ in ps)
    for p in ps:
        p.join(60)
        assert p.exitcode == 0
    assert res[0][1] == [0, 2, 4, 6, 8, 10] and res[1][1] == [1, 3, 5, 7, 9]
    assert sorted(res[0][1] + res[1][1]) == list(range(11))       # every frame exactly once
    for r in res:
        assert r[2] == 15.0 and r[3] == 11.0                       # max of times, sum of units on every rank
