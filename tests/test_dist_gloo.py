"""world_size-2 arg-min exchange on CPU (gloo): the N > 1 host path without GPUs."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from frenetix_motion_planner_b200.dist import ArgminExchange, shard_rows
    ex = ArgminExchange()
    rng = np.random.default_rng(42)
    total = rng.normal(10, 3, 1000)
    total[[17, 503]] = -5.0                        # exact tie across the two shards -> lowest row wins
    first, count = shard_rows(1000, world, rank)
    local = total[first:first + count]
    j = int(np.argmin(local))
    out = [ex.exchange(float(local[j]), first + j)]
    out.append(ex.exchange(float("inf"), -1) if rank == 0 else ex.exchange(3.5, 777))   # one rank has no candidate
    out.append(ex.exchange(float("inf"), -1))                                             # nobody has one
    q.put((rank, out))
    dist.destroy_process_group()


def test_argmin_exchange_two_ranks_gloo():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + os.getpid() % 2000
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = dict(q.get(timeout=120) for _ in procs)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert res[0] == res[1]                          # identical decision on every rank
    assert res[0][0] == (-5.0, 17, 0)
    assert res[0][1] == (3.5, 777, 1)
    assert res[0][2][1] == -1
