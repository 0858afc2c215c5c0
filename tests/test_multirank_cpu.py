"""CPU, world_size 2, gloo: the host-side logic of the multi-GPU path -- gene sharding and the one
gather of per-gene result records -- without any GPU (bench.py does the same over NCCL)."""
import os
import socket
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, genes_per_rank, q):
    import torch
    import torch.distributed as dist
    sys.path.insert(0, ROOT)
    from rvtests_b200 import engine, sharding
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    lo, hi = sharding.shard_range(world * genes_per_rank + 3, rank, world)
    # each rank fabricates the records of its shard (the GPU would compute them)
    rec = np.zeros(hi - lo, dtype=engine.RESULT_DTYPE)
    rec["Q"] = np.arange(lo, hi)
    rec["m_poly"] = rank
    gathered = sharding.gather_records(rec, dist)
    q.put((rank, lo, hi, gathered["Q"].tolist(), gathered["m_poly"].tolist()))
    dist.destroy_process_group()


def test_shard_ranges_cover_everything():
    sys.path.insert(0, ROOT)
    from rvtests_b200 import sharding
    for n in (0, 1, 7, 20000, 20003):
        for w in (1, 2, 3, 8):
            r = [sharding.shard_range(n, k, w) for k in range(w)]
            assert r[0][0] == 0 and r[-1][1] == n
            assert all(r[k][1] == r[k + 1][0] for k in range(w - 1))
            sizes = [b - a for a, b in r]
            assert max(sizes) - min(sizes) <= 1


def test_two_rank_gather_gloo():
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    world, gpr = 2, 5
    ps = [ctx.Process(target=_worker, args=(r, world, port, gpr, q)) for r in range(world)]
    for p in ps:
        p.start()
    outs = [q.get(timeout=120) for _ in ps]
    for p in ps:
        p.join(timeout=60)
        assert p.exitcode == 0
    n = world * gpr + 3
    for rank, lo, hi, Q, mp_ in outs:
        assert Q == list(range(n))          # every rank holds all records, in gene order
        assert mp_ == [0] * 7 + [1] * 6     # 13 genes: ranks own 7 and 6
