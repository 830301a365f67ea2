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

__all__ = ["shard_range", "init_from_env", "all_gather_1d", "all_gather_cols", "toy_hypotest_distributed",
           "patchset_hypotest_distributed"]


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


def all_gather_cols(t, n_total=None):
    """Concatenate every rank's [k, n_local] shard along the last axis, in rank order (shards as
    produced by shard_range).  ONE collective: with ``n_total`` given the shard sizes follow from
    shard_range and need no exchange; equal-size slots are padded to the largest shard."""
    import torch.distributed as dist

    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return t
    world = dist.get_world_size()
    squeeze = t.dim() == 1
    t2 = t.reshape(1, -1) if squeeze else t
    k, n_local = t2.shape
    if n_total is not None:
        counts = [shard_range(n_total, r, world)[1] for r in range(world)]
        if counts[dist.get_rank()] != n_local:
            raise RuntimeError(f"rank shard has {n_local} items, shard_range says {counts[dist.get_rank()]}")
    else:
        mine = torch.tensor([n_local], dtype=torch.int64, device=t.device)
        allc = torch.empty(world, dtype=torch.int64, device=t.device)
        dist.all_gather_into_tensor(allc, mine)
        counts = allc.tolist()
    pad = max(counts)
    if pad == n_local:
        buf = t2.contiguous()
    else:
        buf = torch.zeros(k, pad, dtype=t.dtype, device=t.device)
        buf[:, :n_local] = t2
    flat = torch.empty(world * k * pad, dtype=t.dtype, device=t.device)
    dist.all_gather_into_tensor(flat, buf.reshape(-1))
    out = flat.reshape(world, k, pad)
    if all(c == pad for c in counts):
        res = out.permute(1, 0, 2).reshape(k, world * pad)
    else:
        res = torch.cat([out[r, :, :c] for r, c in enumerate(counts)], dim=1)
    return res.reshape(-1) if squeeze else res


def all_gather_1d(t, n_total=None):
    """1-D convenience form of all_gather_cols."""
    out = all_gather_cols(t, n_total)
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
                                gather=lambda t: all_gather_cols(t, ntoys), **kw)


_PATCH_FIELDS = ("CLs_obs", "CLs_exp", "CLsb", "CLb", "teststat", "qmu", "qmuA", "muhat")


def patchset_hypotest_distributed(workspace, patches, mus=1.0, **kw):
    """``patchset.patchset_hypotest`` with the signal patches sharded over the ranks of the default
    process group by contiguous index (SURVEY.md section 8 rows e + f4: signal points are
    independent units).  Every rank builds its own group models from its shard; ONE all-gather of
    a ``[rows, n_patches]`` block returns every patch's results to every rank."""
    import numpy as np
    import torch.distributed as dist

    from . import patchset

    if isinstance(patches, dict):
        patches = patches["patches"]
    patches = list(patches)
    if dist.is_available() and dist.is_initialized():
        rank, world = dist.get_rank(), dist.get_world_size()
    else:
        rank, world = 0, 1
    K = len(patches)
    M = int(np.atleast_1d(np.asarray(mus, float)).size)
    start, count = shard_range(K, rank, world)
    if world == 1:
        return patchset.patchset_hypotest(workspace, patches, mus, **kw)
    widths = {"CLs_exp": 5 * M, "muhat": 1}
    rows = sum(widths.get(f, M) for f in _PATCH_FIELDS) + 3
    device = kw.get("device")
    if count:
        out = patchset.patchset_hypotest(workspace, patches[start:start + count], mus, **kw)
        cols = [out[f].reshape(count, -1) for f in _PATCH_FIELDS]
        extra = torch.tensor([[1.0 if m == "group" else 0.0, out["n_fits"] / count, out["n_failed"] / count]
                              for m in out["mode"]], dtype=cols[0].dtype, device=cols[0].device)
        block = torch.cat(cols + [extra], dim=1).T.contiguous()
    else:
        dev = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
        block = torch.zeros(rows, 0, dtype=torch.float64, device=dev)
    full = all_gather_cols(block, K).T  # [K, rows]
    res, pos = {}, 0
    for f in _PATCH_FIELDS:
        wdt = widths.get(f, M)
        res[f] = full[:, pos:pos + wdt]
        pos += wdt
    res["CLs_exp"] = res["CLs_exp"].reshape(K, M, 5)
    res["muhat"] = res["muhat"].reshape(K)
    res["mode"] = ["group" if v > 0.5 else "single" for v in full[:, pos].tolist()]
    res["n_fits"] = int(round(float(full[:, pos + 1].sum())))
    res["n_failed"] = int(round(float(full[:, pos + 2].sum())))
    res["mu"] = torch.as_tensor(np.atleast_1d(np.asarray(mus, float)), device=full.device)
    return res
