"""SURVEY.md §8 rows e + f4 on several GPUs: K signal patches on the C3 background sharded over the
ranks (one all-gather of the per-patch results); wall clock of the whole call, max over ranks.

    python tools/patchset_dist.py --k 256                                  # one GPU
    python -m torch.distributed.run --nproc-per-node N --master-addr 127.0.0.1 tools/patchset_dist.py --k 256
"""
import argparse
import json
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from patchset_probe import workspace  # noqa: E402

from pyhf_b200 import dist as hd  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--k", type=int, default=256)
    ap.add_argument("--group", type=int, default=32)
    ap.add_argument("--reps", type=int, default=3)
    a = ap.parse_args()
    rank, world, local = hd.init_from_env()
    torch.cuda.set_device(local)
    ws, patches = workspace("c3", a.k)
    best, out = None, None
    for _ in range(a.reps):
        if world > 1:
            torch.distributed.barrier()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        out = hd.patchset_hypotest_distributed(ws, patches, 1.0, group_size=a.group)
        torch.cuda.synchronize()
        dt = torch.tensor([time.perf_counter() - t0], device="cuda")
        if world > 1:
            torch.distributed.all_reduce(dt, op=torch.distributed.ReduceOp.MAX)
        best = float(dt) if best is None else min(best, float(dt))
    if rank == 0:
        cls = out["CLs_obs"][:, 0]
        print(json.dumps({"n_gpus": world, "K": a.k, "group_size": a.group, "seconds": best,
                          "ms_per_signal_point": 1e3 * best / a.k, "n_fits": out["n_fits"],
                          "n_failed": out["n_failed"], "modes": sorted(set(out["mode"])),
                          "CLs_obs_checksum": float(cls.sum()), "CLs_obs_first": [float(v) for v in cls[:3]]}))
    if world > 1:
        torch.distributed.destroy_process_group()


if __name__ == "__main__":
    main()
