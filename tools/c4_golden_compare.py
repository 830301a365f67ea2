import sys, os
sys.path.insert(0, "."); sys.path.insert(0, "tests"); sys.path.insert(0, "baseline/_ref")
import numpy as np, torch
from conftest import synthetic_host_plan
from pyhf_b200.batched import DevicePlan, qmu_tilde_batch
z = np.load("tests/golden/c4_fits.npz")
hp = synthetic_host_plan("spec_c4"); plan = DevicePlan(hp)
out = qmu_tilde_batch(plan, 1.0, z["toys"])
np.set_printoptions(precision=9, linewidth=200)
for w in ("fixed", "free"):
    f = out[w].fun.cpu().numpy()
    print(w, "ours", f); print(w, "ours - default", f - z[f"toy_{w}_fun_default"]); print(w, "ours - tight", f - z[f"toy_{w}_fun_tight"])
