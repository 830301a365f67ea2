"""K5 (hf_teststat) in all three modes and the batched test statistics built on it, against the
reference's own functions: qmu / qmu_tilde (infer/test_statistics.py:63-254), tmu / tmu_tilde
(:257-436) and q0 (:439-523), evaluated by the unmodified reference with a tight optimizer on the
observed data and 7 toys of two models (tests/golden/teststats.npz, stage_teststats of
tests/golden/make_golden.py).  Tolerance 1e-6 absolute (BASELINE.json north_star)."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ts_golden():
    import os

    from conftest import GOLDEN

    return np.load(os.path.join(GOLDEN, "teststats.npz"))


@pytest.mark.parametrize("mode,mu", [("qtilde", 1.0), ("q", 0.4), ("q0", 0.0), ("t", 1.0), ("ttilde", 0.3)])
def test_k5_modes_against_the_reference_arithmetic(mode, mu, cuda):
    """hf_teststat on synthetic fit results: clip at zero (test_statistics.py:56), zeroing where
    muhat > mu (q / q~, :57-60) or muhat < 0 (q0, :507-520), no zeroing for t (:38-60); ties
    muhat == mu and negative differences included"""
    from pyhf_b200 import batched

    rng = np.random.default_rng(5)
    n = 1000
    f_free = rng.uniform(10, 20, n)
    f_fixed = f_free + rng.normal(1.0, 1.5, n)  # ~25 % negative differences
    muhat = rng.uniform(-1.0, 2.0, n)
    muhat[:10] = mu      # ties: not zeroed (strict comparison in the reference)
    muhat[10:20] = 0.0
    q = batched.teststat(torch.as_tensor(f_fixed, device=cuda), torch.as_tensor(f_free, device=cuda),
                         torch.as_tensor(muhat, device=cuda), mu, mode).cpu().numpy()
    want = np.clip(f_fixed - f_free, 0.0, None)
    if mode in ("q", "qtilde"):
        want = np.where(muhat > mu, 0.0, want)
    elif mode == "q0":
        want = np.where(muhat < 0.0, 0.0, want)
    assert np.array_equal(q, want)


@pytest.mark.parametrize("model", ["hello", "mixed"])
@pytest.mark.parametrize("stat,mode,mu", [("qmu", "q", 1.0), ("qmu_tilde", "qtilde", 1.0), ("tmu", "t", 1.0),
                                           ("tmu_tilde", "ttilde", 1.0), ("q0", "q0", 0.0)])
def test_batched_test_statistics_match_the_reference(model, stat, mode, mu, ts_golden, cuda):
    from pyhf_b200.batched import DevicePlan, qmu_tilde_batch
    from pyhf_b200.plan import HostPlan

    z = ts_golden
    pre = f"{model}__plan__"
    hp = HostPlan.from_arrays({k[len(pre):]: z[k] for k in z.files if k.startswith(pre)})
    plan = DevicePlan(hp, cuda)
    rows = z[f"{model}__rows"]
    out = qmu_tilde_batch(plan, mu, rows, lo=z[f"{model}__{stat}_lo"], hi=z[f"{model}__{stat}_hi"], mode=mode)
    assert (out["fixed"].status == 0).all() and (out["free"].status == 0).all()
    q = out["q"].cpu().numpy()
    ref = z[f"{model}__{stat}"]
    # both zeroing branches are present in the reference values (so the comparison means something)
    # (q0 with the POI bounded at 0: muhat sits ON the bound, is not < 0, and q0 = fixed - free ~ 1e-13)
    if mode in ("q", "qtilde", "q0"):
        assert (ref < 1e-9).any() and (ref > 1e-3).any()
    assert np.abs(q - ref).max() < 1e-6, (stat, q, ref)
    muhat = out["muhat"].cpu().numpy()
    assert np.abs(muhat - z[f"{model}__{stat}_muhat"]).max() < 2e-4
