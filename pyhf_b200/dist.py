"""Multi-GPU: one process per GPU, toys / scan points / batch rows sharded by index.

The path shards trivially: every rank holds a replica of the plan
(<= 10 MB of tables) and works on the contiguous index range ``shard_range(n, rank, world)``.
Toy ``i`` draws from the Philox counter ``(seed, i)`` -- its *global* index -- so results do
not depend on the number of ranks.  The only collective is the final all-gather of the
per-toy test statistics (NCCL over NVLink on GPUs, gloo on CPU for the tests): 2 x ntoys fp64.
"""
from __future__ import annotations

import os

import torch

__all__ = ["shard_range", "init_from_env", "all_gather_1d", "toy_hypotest_distributed"]


def shard_range(n, rank, world):
    """(start, count) of rank's contiguous shard of n items; counts differ by at most one."""
    base, rem = divmod(int(n), int(world))
    count = base + (1 if rank < rem else 0)
    start = rank * base + min(rank, rem)
    return start, count


def init_from_env(backend=None):
    """torchrun-style rendezvous (RANK / WORLD_SIZE / LOCAL_RANK / MASTER_*)."""
    import torch.distributed as dist

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1 and not dist.is_initialized():
        if backend is None:
            backend = "nccl" if torch.cuda.is_available() else "gloo"
        if backend == "nccl":
            torch.cuda.set_device(local)
            dist.init_process_group(backend, device_id=torch.device("cuda", local))
        else:
            dist.init_process_group(backend)
    return rank, world, local


def all_gather_1d(t, n_total=None):
    """Concatenate every rank's 1-D shard in rank order (shards as produced by shard_range)."""
    import torch.distributed as dist

    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return t
    world = dist.get_world_size()
    n_local = torch.tensor([t.numel()], dtype=torch.int64, device=t.device)
    counts = [torch.zeros_like(n_local) for _ in range(world)]
    dist.all_gather(counts, n_local)
    counts = [int(c) for c in counts]
    pad = max(counts)
    buf = torch.zeros(pad, dtype=t.dtype, device=t.device)
    buf[: t.numel()] = t
    parts = [torch.empty_like(buf) for _ in range(world)]
    dist.all_gather(parts, buf)
    out = torch.cat([p[:c] for p, c in zip(parts, counts)])
    if n_total is not None and out.numel() != n_total:
        raise RuntimeError(f"gathered {out.numel()} items, expected {n_total}")
    return out


def toy_hypotest_distributed(plan, mu_test, data_obs, ntoys, seed=0, **kw):
    """``batched.toy_hypotest`` with the toys sharded over the ranks of the default process
    group; every rank returns the same CLs."""
    import torch.distributed as dist

    from . import batched

    if dist.is_available() and dist.is_initialized():
        rank, world = dist.get_rank(), dist.get_world_size()
    else:
        rank, world = 0, 1
    start, count = shard_range(ntoys, rank, world)
    return batched.toy_hypotest(plan, mu_test, data_obs, count, seed=seed, toy_offset=start,
                                gather=lambda t: all_gather_1d(t, ntoys), **kw)
