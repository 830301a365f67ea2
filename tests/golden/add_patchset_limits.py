"""Augment tests/golden/patchset.json with the reference's upper limits for the synthetic signal patches
1 and 4: ``pyhf.infer.intervals.upper_limits.upper_limit(data, model, scan,
return_results=True)`` (upper_limits.py:148-213) on the patched models, numpy backend, SLSQP at
tolerance 1e-12.

    python tests/golden/add_patchset_limits.py     (needs baseline/_ref, see oracle/make_ref.py)
"""
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "baseline", "_ref"))
import pyhf  # noqa: E402
from pyhf.infer.intervals.upper_limits import upper_limit  # noqa: E402

from pyhf_b200.patchset import apply_patch  # noqa: E402

path = os.path.join(HERE, "patchset.json")
cases = json.load(open(path))
c = cases["synthetic"]
scan = np.linspace(0.0, 4.0, 17)
pyhf.set_backend("numpy", pyhf.optimize.scipy_optimizer(tolerance=1e-12))
obs, exp = [], []
which = [1, 4]
for p in (c["patches"][k] for k in which):
    w = apply_patch(c["workspace"], p)
    m = w["measurements"][0]
    model = pyhf.Model({"channels": w["channels"], "parameters": m["config"]["parameters"]}, poi_name=m["config"]["poi"])
    observations = {o["name"]: o["data"] for o in w["observations"]}
    data = sum((list(observations[ch]) for ch in model.config.channels), []) + list(model.config.auxdata)
    o, e, _ = upper_limit(data, model, scan, return_results=True)
    obs.append(float(o))
    exp.append([float(v) for v in e])
c["limits"] = {"scan": scan.tolist(), "patches": which, "obs_limit_tight": obs, "exp_limits_tight": exp}
json.dump(cases, open(path, "w"), indent=1)
print(obs, exp)
