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


def _bolt_worker(rank, world, port, q):
    """SNP-sharded H-product of the BoltLMM null fit (BASELINE configs[4]): each rank holds a slice of the
    panel's SNPs, computes its partial of X X'v / M and the ONE sum over ranks gives the full product -- the host-side
    logic of rvt_bolt_fit_null_sharded, here in numpy over gloo."""
    import torch
    import torch.distributed as dist
    sys.path.insert(0, ROOT)
    from rvtests_b200 import sharding
    from oracle import bolt_oracle as BO
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    rng = np.random.default_rng(7)
    N, M, C, R = 240, 131, 2, 5
    G = rng.binomial(2, rng.uniform(0.05, 0.5, M)[:, None], size=(M, N)).astype(np.int8)
    G[rng.random((M, N)) < 0.02] = -1
    covar = np.column_stack([np.ones(N), rng.normal(size=N)])
    X, Z, _ = BO.prepare(G, covar, np.zeros(N))
    v = rng.normal(size=(N, R))
    lo, hi = sharding.snp_shard(M, rank, world)
    Xl = X[:, lo:hi]
    # local: X_l' (I - ZZ') v, then X_l (.)  -- the partial of the top rows; Z'X_l (.) the partial of the covariate rows
    xy = Xl.T @ v - (Xl.T @ Z) @ (Z.T @ v)
    part = np.vstack([Xl @ xy, (Z.T @ Xl) @ xy]) / M
    t = torch.from_numpy(part.copy())
    dist.all_reduce(t)                                   # the one collective per H-product
    full = BO.Fit(X, Z, np.zeros(N), mc_trials=3)
    want = full.Hx(0.0, v)
    got = t.numpy()
    q.put((rank, lo, hi, float(np.max(np.abs(got[:N] - want))), float(np.max(np.abs(got[N:] - Z.T @ want)))))
    dist.destroy_process_group()


def test_two_rank_bolt_snp_sharding_gloo():
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    ps = [ctx.Process(target=_bolt_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in ps:
        p.start()
    outs = sorted(q.get(timeout=120) for _ in ps)
    for p in ps:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert [(o[1], o[2]) for o in outs] == [(0, 66), (66, 131)]
    for _rank, _lo, _hi, e_top, e_bot in outs:
        assert e_top <= 1e-10 and e_bot <= 1e-10


def test_bolt_bench_panel_shards_hold_the_same_bytes_at_every_world_size():
    """bench.py --workload bolt synthesises the panel in 256-row chunks keyed by the GLOBAL row index: a rank's shard of the SNP
    rows is the same bytes whatever the number of ranks (the strong-scaling runs at 1, 2, 4, 8 GPUs fit the same cohort)."""
    import torch
    sys.path.insert(0, ROOT)
    from tools import bolt_bench
    from rvtests_b200 import sharding
    dev = torch.device("cpu")
    N, M = 1003, 700
    full = bolt_bench.synth_rows(torch, dev, N, 0, M)
    assert full.shape == (M, (N + 3) // 4) and full.dtype == torch.uint8
    for world in (2, 3, 8):
        parts = [bolt_bench.synth_rows(torch, dev, N, *sharding.snp_shard(M, r, world)) for r in range(world)]
        assert torch.equal(torch.cat(parts), full), world
    # padding bits of the last byte are 00, missing calls are code 01 at about 1 %
    codes = torch.stack([(full >> s) & 3 for s in (0, 2, 4, 6)], dim=2).reshape(M, -1)
    assert int(codes[:, N:].sum()) == 0
    frac_missing = float((codes[:, :N] == 1).float().mean())
    assert 0.005 < frac_missing < 0.02
