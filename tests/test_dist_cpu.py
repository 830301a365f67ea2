"""CPU, world_size 2 over gloo: sharding + the final all-gather of per-toy test statistics."""
import os
import socket
import sys

import pytest
import torch
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_shard_range_partitions():
    from pyhf_b200.dist import shard_range

    for n in (0, 1, 7, 100, 100001):
        for world in (1, 2, 3, 8):
            pieces = [shard_range(n, r, world) for r in range(world)]
            assert sum(c for _, c in pieces) == n
            pos = 0
            for s, c in pieces:
                assert s == pos
                pos += c
            counts = [c for _, c in pieces]
            assert max(counts) - min(counts) <= 1


def _worker(rank, world, port, n):
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank),
                      WORLD_SIZE=str(world), LOCAL_RANK=str(rank))
    import torch.distributed as dist

    from pyhf_b200.dist import all_gather_1d, init_from_env, shard_range

    r, w, _ = init_from_env("gloo")
    assert (r, w) == (rank, world)
    full = torch.arange(n, dtype=torch.float64) ** 0.5  # "test statistic of toy i"
    start, count = shard_range(n, rank, world)
    got = all_gather_1d(full[start:start + count].clone(), n)
    assert torch.equal(got, full)
    # the two hypotheses' statistics travel as one [2, n_local] block; also without n_total
    from pyhf_b200.dist import all_gather_cols

    two = torch.stack([full, -full])
    got2 = all_gather_cols(two[:, start:start + count].clone(), n)
    assert torch.equal(got2, two)
    assert torch.equal(all_gather_cols(two[:, start:start + count].clone()), two)
    # p-value arithmetic of EmpiricalDistribution.pvalue on the gathered array is rank-independent
    assert float((got >= 3.0).double().mean()) == float((full >= 3.0).double().mean())
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("n", [10, 1001])
def test_all_gather_two_ranks_gloo(n):
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    mp.spawn(_worker, args=(2, port, n), nprocs=2, join=True)
