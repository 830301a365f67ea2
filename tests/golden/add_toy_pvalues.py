"""Augment tests/golden/hello_toys.npz with the reference's empirical p-value arithmetic on the
stored toy test statistics (EmpiricalDistribution.pvalue and ToyCalculator.pvalues /
expected_pvalues, infer/calculators.py:563-700, 872-973), so that the device implementation of
that arithmetic can be checked without the reference at test time.

    python tests/golden/add_toy_pvalues.py        (needs baseline/_ref, see oracle/make_ref.py)
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, os.path.join(ROOT, "baseline", "_ref"))
import pyhf  # noqa: E402
from pyhf.infer.calculators import EmpiricalDistribution, ToyCalculator  # noqa: E402

pyhf.set_backend("numpy", "scipy")
path = os.path.join(HERE, "hello_toys.npz")
z = dict(np.load(path, allow_pickle=False))
sb = EmpiricalDistribution(pyhf.tensorlib.astensor(z["sig2_q_tight"]))
b = EmpiricalDistribution(pyhf.tensorlib.astensor(z["bkg2_q_tight"]))
calc = ToyCalculator.__new__(ToyCalculator)  # pvalues / expected_pvalues use no instance state
qobs = float(z["qobs_tight"])
z["toys2_pvalues"] = np.array([float(v) for v in calc.pvalues(qobs, sb, b)])
z["toys2_expected_pvalues"] = np.array([[float(v) for v in band] for band in calc.expected_pvalues(sb, b)])
z["toys2_expected_values_b"] = np.array([float(b.expected_value(n)) for n in (-2, -1, 0, 1, 2)])
np.savez_compressed(path, **z)
print(z["toys2_pvalues"], z["toys2_expected_pvalues"], z["toys2_expected_values_b"], sep="\n")
