import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "baseline", "_ref"))
import numpy as np, torch, pyhf
from pyhf_b200 import batched, synth
from pyhf_b200.plan import extract_plan
model = pyhf.Model(synth.spec_c3(), poi_name="mu", validate=False)
hp = extract_plan(model); plan = batched.DevicePlan(hp)
toys = plan.sample_toys(hp.init, 100000, seed=11)
out = batched.qmu_tilde_batch(plan, 1.0, toys)
for kind in ("fixed", "free"):
    st = out[kind].status
    bad = torch.nonzero(st != 0).flatten()
    print(kind, "failed", bad.tolist(), "niter", out[kind].niter[bad].tolist())
    np.savez(os.path.join(ROOT, "gpurun_out", f"fail_{kind}.npz"), toys=toys[bad].cpu().numpy(), x=out[kind].x[bad].cpu().numpy(), fun=out[kind].fun[bad].cpu().numpy(), idx=bad.cpu().numpy())
