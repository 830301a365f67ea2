"""C3 toy-fit probe for profiling: one qmu_tilde_batch over N toys"""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "baseline", "_ref"))
import numpy as np, torch, pyhf
from pyhf_b200 import batched, synth
from pyhf_b200.plan import extract_plan
n = int(sys.argv[1]) if len(sys.argv) > 1 else 20000
model = pyhf.Model(synth.spec_c3(), poi_name="mu", validate=False)
hp = extract_plan(model); plan = batched.DevicePlan(hp)
toys = plan.sample_toys(hp.init, n, seed=11)
batched.qmu_tilde_batch(plan, 1.0, toys)  # warm-up at full size (workspaces are sized on first use)
torch.cuda.synchronize(); t0 = time.time()
out = batched.qmu_tilde_batch(plan, 1.0, toys, verbose=int(os.environ.get("HF_VERBOSE", "0")))
torch.cuda.synchronize(); dt = time.time() - t0
st = torch.cat([out["fixed"].status, out["free"].status])
print(f"{n} toys {dt:.3f}s {2*n/dt:.0f} fits/s failed {(st!=0).sum().item()} status2 {(st==2).sum().item()} status1 {(st==1).sum().item()} iters mean {out['free'].niter.float().mean().item():.1f} max {out['free'].niter.max().item()}")
