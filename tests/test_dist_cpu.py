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


def _patch_worker(rank, world, port):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank),
                      WORLD_SIZE=str(world), LOCAL_RANK=str(rank))
    import json

    import numpy as np
    import torch.distributed as dist

    from pyhf_b200 import patchset as PS
    from pyhf_b200.dist import init_from_env, patchset_hypotest_distributed
    from test_patchset import OraclePlan

    init_from_env("gloo")
    PS.DevicePlan = OraclePlan  # CPU stand-in (oracle fits): this test is about the sharding
    c = json.load(open(os.path.join(ROOT, "tests", "golden", "patchset.json")))["synthetic"]
    patches = c["patches"][:3]  # ranks get 2 + 1 patches
    out = patchset_hypotest_distributed(c["workspace"], patches, c["mus"][:2], device="cpu")
    assert out["mode"] == ["group"] * 3 and out["n_failed"] == 0
    assert out["n_fits"] == (2 * 2 * 3 + 1) + (2 * 1 * 3 + 1)
    np.testing.assert_allclose(out["CLs_obs"].numpy(), np.array(c["cls_obs_tight"])[:3, :2], rtol=0, atol=1e-6)
    np.testing.assert_allclose(out["CLs_exp"].numpy(), np.array(c["cls_exp_tight"])[:3, :2], rtol=0, atol=1e-6)
    assert tuple(out["muhat"].shape) == (3,) and tuple(out["qmu"].shape) == (3, 2)
    # more ranks than patches: the empty shard takes part in the collective
    one = patchset_hypotest_distributed(c["workspace"], patches[:1], 1.0, device="cpu")
    assert abs(float(one["CLs_obs"][0, 0]) - c["cls_obs_tight"][0][1]) < 1e-6 and len(one["mode"]) == 1
    dist.barrier()
    dist.destroy_process_group()


def test_patchset_sharded_over_two_ranks_gloo():
    """signal patches sharded by index over 2 ranks, one all-gather of the per-patch results;
    checked against the reference's per-patch CLs (tests/golden/patchset.json)"""
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    mp.spawn(_patch_worker, args=(2, port), nprocs=2, join=True)
