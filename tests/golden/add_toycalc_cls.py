"""Augment tests/golden/hello_toys.npz with the CLs of the reference's UNCHANGED toy-based
hypotest (ToyCalculator, infer/calculators.py:754-915) on the hello-world model under
np.random.seed(0), 300 toys per hypothesis, default and tight optimizer tolerance.  The toys are
drawn by numpy's global RNG in the reference's order (s+b batch first, then b-only); the b200
backend's exact-parity mode (toy_rng="numpy") draws the same toys.

    python tests/golden/add_toycalc_cls.py        (needs baseline/_ref, see oracle/make_ref.py)
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "baseline", "_ref"))
import pyhf  # noqa: E402

from pyhf_b200 import synth  # noqa: E402

NTOYS = 300
path = os.path.join(HERE, "hello_toys.npz")
z = dict(np.load(path, allow_pickle=False))
model = pyhf.Model(synth.spec_hello(), poi_name="mu")
data = [51.0, 48.0] + model.config.auxdata
for tag, kw in (("", {}), ("_tight", {"tolerance": 1e-12})):
    pyhf.set_backend("numpy", pyhf.optimize.scipy_optimizer(**kw))
    np.random.seed(0)
    obs, exp = pyhf.infer.hypotest(1.0, data, model, test_stat="qtilde", calctype="toybased", ntoys=NTOYS,
                                   track_progress=False, return_expected_set=True)
    z[f"toycalc_cls_obs{tag}"] = np.float64(obs)
    z[f"toycalc_cls_exp{tag}"] = np.array([float(e) for e in exp])
    print(tag, float(obs), [float(e) for e in exp])
z["toycalc_ntoys"] = np.int64(NTOYS)
pyhf.set_backend("numpy", "scipy")
np.savez_compressed(path, **z)
