"""C4 (P=300) fits: correctness (first-order optimality via the oracle gradient) and timing"""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "baseline", "_ref"))
import numpy as np, torch, pyhf
from pyhf_b200 import batched, synth
from pyhf_b200.plan import extract_plan
from oracle import hf_oracle as O
model = pyhf.Model(synth.spec_c4(), poi_name="mu", validate=False)
hp = extract_plan(model); plan = batched.DevicePlan(hp); op = O.Plan(hp.arrays())
n = int(sys.argv[1]) if len(sys.argv) > 1 else 64
toys = plan.sample_toys(hp.init, n, seed=5)
torch.cuda.synchronize(); t0 = time.time()
out = batched.qmu_tilde_batch(plan, 1.0, toys, verbose=int(os.environ.get("HF_VERBOSE", "0")))
torch.cuda.synchronize(); dt = time.time() - t0
st = torch.cat([out["fixed"].status, out["free"].status])
print(f"c4 {n} toys: {dt:.2f}s -> {2*n/dt:.1f} fits/s; failed {(st != 0).sum().item()}; iters {out['free'].niter.float().mean().item():.1f}; q mean {out['q'].mean().item():.4f}")
x = out["free"].x[:4].cpu().numpy(); d = toys[:4].cpu().numpy()
v, g = O.logpdf_grad(op, x, d); g = -2 * g
free = ~hp.fixed.astype(bool)
at_lo = (x <= hp.lo + 1e-12) & (g > 0); at_hi = (x >= hp.hi - 1e-12) & (g < 0)
pg = np.where(free[None] & ~at_lo & ~at_hi, g, 0.0)
print("max projected gradient of 2NLL at the GPU minima:", np.abs(pg).max(axis=1), " 2NLL:", -2 * v, out["free"].fun[:4].cpu().numpy())
