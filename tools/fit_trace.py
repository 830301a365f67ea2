"""trace of fits on a golden: python tools/fit_trace.py NAME [free|fixed] [toy_seed ntoys fit_index]
(HF_FIT_SOLO=0 selects the lock-step driver; verbose=3 traces fits 0 and 1, HF_FIT_TRACE_FIT one more)"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests")); sys.path.insert(0, os.path.join(ROOT, "baseline", "_ref"))
import numpy as np, torch
from conftest import Golden
from pyhf_b200.batched import DevicePlan
name = sys.argv[1]
which = sys.argv[2] if len(sys.argv) > 2 else "free"
g = Golden(os.path.join(ROOT, "tests", "golden", f"{name}.npz"))
hp = g.host_plan()
plan = DevicePlan(hp)
kw = {}
if which == "fixed":
    fx = hp.fixed.copy(); fx[hp.poi_index] = 1
    kw = dict(fixed=fx)
    if len(sys.argv) <= 3:
        init = hp.init.copy(); init[hp.poi_index] = g.meta["mu_test"]; kw["init"] = init
data = np.asarray(g["data"], float)
if len(sys.argv) > 3:
    seed, n, idx = int(sys.argv[3]), int(sys.argv[4]), int(sys.argv[5])
    toys = plan.sample_toys(hp.init, n, seed=seed)
    toys[0] = torch.as_tensor(data, device=toys.device)
    data = toys[idx].cpu().numpy()
    print("data", data.tolist())
r = plan.fit(data, verbose=3, **kw)
print("x", r.x.cpu().numpy().tolist(), "f", float(r.fun[0]), "status", int(r.status[0]), "niter", int(r.niter[0]))
key = "fixed" if which == "fixed" else "free"
if len(sys.argv) <= 3 and f"{key}_x_tight" in g:
    print("ref", g[f"{key}_x_tight"], float(g[f"{key}_fun_tight"]))
