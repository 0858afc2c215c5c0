"""Multi-GPU plumbing of the path: genes are independent, so each rank owns a
contiguous shard of the gene list and the only collective is ONE gather of the fixed-size per-gene
result records (168 B each).  Backend-agnostic (`dist` = torch.distributed with NCCL on GPUs, gloo in
the CPU tests); no arithmetic here."""
from __future__ import annotations

import numpy as np


def shard_range(n_genes: int, rank: int, world: int):
    """contiguous, balanced (sizes differ by at most one) shard [lo, hi) of rank `rank`"""
    base, extra = divmod(n_genes, world)
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def gather_records(local: np.ndarray, dist, device=None) -> np.ndarray:
    """all ranks end up with the records of every gene, in gene order.  `local` is a structured
    numpy array (engine.RESULT_DTYPE); shards may differ in length by one (padded for the gather)."""
    import torch
    world = dist.get_world_size()
    n_local = torch.tensor([len(local)], dtype=torch.int64, device=device)
    counts = [torch.zeros_like(n_local) for _ in range(world)]
    dist.all_gather(counts, n_local)
    counts = [int(c.item()) for c in counts]
    cap = max(counts) if counts else 0
    item = local.dtype.itemsize
    buf = torch.zeros(cap * item, dtype=torch.uint8, device=device)
    if len(local):
        buf[: len(local) * item] = torch.from_numpy(local.view(np.uint8).reshape(-1).copy()).to(buf.device)
    out = [torch.zeros_like(buf) for _ in range(world)]
    dist.all_gather(out, buf)
    parts = [np.frombuffer(o.cpu().numpy().tobytes(), dtype=local.dtype)[:c] for o, c in zip(out, counts)]
    return np.concatenate(parts) if parts else local


def snp_shard(m_total: int, rank: int, world: int):
    """BoltLMM null fit (BASELINE configs[4]): the panel's SNP rows are sharded, [lo, hi) of rank `rank`"""
    return shard_range(m_total, rank, world)


class _DevView:
    """a device buffer of doubles as an object torch.as_tensor understands (__cuda_array_interface__, no copy)"""

    def __init__(self, ptr: int, count: int):
        self.__cuda_array_interface__ = {"shape": (int(count),), "typestr": "<f8", "data": (int(ptr), False), "version": 2}


def torch_allreduce(dist, device=None, staged: bool = False):
    """The rvt_allreduce_fn of rvt_bolt_fit_null_sharded on torch.distributed: (dev_ptr, count, cuda_stream) -> 0.
    NCCL: the device buffer is wrapped in place and summed by ncclAllReduce, ordered on the engine's stream (handed to the
    callback).  staged=True (gloo, the CPU-side tests with two ranks on one GPU): the
    buffer goes through host memory."""
    import torch

    def fn(ptr: int, count: int, stream: int) -> int:
        t = torch.as_tensor(_DevView(ptr, count), device=device)
        if staged:
            torch.cuda.synchronize()
            h = t.cpu()
            dist.all_reduce(h)
            t.copy_(h)
            torch.cuda.synchronize()
        elif stream:
            with torch.cuda.stream(torch.cuda.ExternalStream(stream)):   # ordered on the engine's own stream
                dist.all_reduce(t)
        else:
            dist.all_reduce(t)
        return 0

    return fn
