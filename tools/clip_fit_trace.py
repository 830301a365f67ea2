"""trace of a fit on a golden (HF verbose=3): python tools/clip_fit_trace.py NAME [fixed]"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
from conftest import Golden
from pyhf_b200.batched import DevicePlan
name = sys.argv[1] if len(sys.argv) > 1 else "clip_bin0"
g = Golden(os.path.join(ROOT, "tests", "golden", f"{name}.npz"))
hp = g.host_plan()
plan = DevicePlan(hp)
kw = {}
if len(sys.argv) > 2:
    fx = hp.fixed.copy(); fx[hp.poi_index] = 1
    init = hp.init.copy(); init[hp.poi_index] = g.meta["mu_test"]
    kw = dict(init=init, fixed=fx)
r = plan.fit(g["data"], verbose=3, **kw)
print("x", r.x.cpu().numpy(), "f", float(r.fun[0]), "status", int(r.status[0]), "niter", int(r.niter[0]))
key = "fixed" if kw else "free"
print("ref", g[f"{key}_x_tight"], float(g[f"{key}_fun_tight"]))
