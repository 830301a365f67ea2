"""config 3 end to end: q~ toy-based CLs on C3 with the toys sharded over the ranks of a torchrun
job (one process per GPU); prints one JSON line on rank 0.
    torchrun --nproc-per-node N tools/toy_hypotest_dist.py --ntoys 100000"""
import argparse
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "baseline", "_ref"))
import numpy as np
import torch

ap = argparse.ArgumentParser()
ap.add_argument("--ntoys", type=int, default=100000)
ap.add_argument("--model", default="c3")
ap.add_argument("--mu", type=float, default=1.0)
ap.add_argument("--reps", type=int, default=1, help="timed repetitions; the JSON line reports the median")
args = ap.parse_args()

import pyhf
from pyhf_b200 import batched, dist, synth
from pyhf_b200.plan import extract_plan

rank, world, local = dist.init_from_env()
torch.cuda.set_device(local)
model = pyhf.Model(getattr(synth, f"spec_{args.model}")(), poi_name="mu", validate=False)
hp = extract_plan(model)
np.random.seed(0)
data = np.asarray(model.make_pdf(pyhf.tensorlib.astensor(hp.init)).sample((1,))[0])
plan = batched.DevicePlan(hp)
dist.toy_hypotest_distributed(plan, args.mu, data, args.ntoys, seed=3)  # warm-up at full size: workspaces are sized on first use
times = []
for rep in range(args.reps):
    if world > 1:
        torch.distributed.barrier()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    res = dist.toy_hypotest_distributed(plan, args.mu, data, args.ntoys, seed=7)
    torch.cuda.synchronize()
    dt = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device="cuda")
    if world > 1:
        torch.distributed.all_reduce(dt, op=torch.distributed.ReduceOp.MAX)
    times.append(float(dt[0]))
dt = torch.tensor([float(np.median(times))], dtype=torch.float64, device="cuda")
nfail = torch.tensor([res["n_failed"]], dtype=torch.int64, device="cuda")
if world > 1:
    torch.distributed.all_reduce(nfail)
prof = {}
dist.toy_hypotest_distributed(plan, args.mu, data, args.ntoys, seed=7, profile=prof)  # untimed: phase split
if rank == 0:
    print(json.dumps({"workload": f"{args.model} qtilde toy hypotest", "ntoys_per_hypothesis": args.ntoys,
                      "n_gpus": world, "seconds": float(dt[0]), "all_reps_s": [round(t, 4) for t in times], "toy_fits_per_s": 4 * args.ntoys / float(dt[0]),
                      "CLs": res["CLs"], "CLsb": res["CLsb"], "CLb": res["CLb"], "qobs": res["qobs"],
                      "failed_fits": int(nfail[0]), "phases_rank0_s": {k: round(v, 4) for k, v in prof.items()},
                      "q_sig_checksum": float(res["q_sig"].sum()), "q_bkg_checksum": float(res["q_bkg"].sum())}))
if world > 1:
    torch.distributed.destroy_process_group()
