"""the asymptotic CLs scan as two batched fit launches
(85 distinct fits instead of the reference's 205), against the reference's own scan results
(tests/golden/scan_2bin_2channel.npz: pyhf.infer.intervals.upper_limits.upper_limit on
BASELINE config 2 with tolerance=1e-12) and, per scan point, against the hello-world goldens."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _plan(g, cuda):
    from pyhf_b200.batched import DevicePlan
    from pyhf_b200.compile import compile_spec

    s = g.meta.get("settings") or {}
    hp = compile_spec(g.spec(), poi_name=g.meta.get("poi_name", "mu"), modifier_settings=s.get("modifier_settings"))
    return DevicePlan(hp, cuda)


def test_config2_scan_in_two_launches(golden, cuda):
    from pyhf_b200 import batched

    g = golden("scan_2bin_2channel")
    plan = _plan(g, cuda)
    l0 = plan.launch_count
    out = batched.asymptotic_scan(plan, g["scan"], g["data"])
    assert out["n_fits"] == 85 and out["n_failed"] == 0
    assert plan.launch_count > l0
    cls_obs = out["CLs_obs"].cpu().numpy()
    cls_exp = out["CLs_exp"].cpu().numpy()
    # tolerance: CLs within 1e-6 of the tight reference (BASELINE north star), the
    # default-tolerance reference itself is only converged to ~3e-5
    np.testing.assert_allclose(cls_obs, g["cls_obs_tight"], rtol=0, atol=1e-6)
    np.testing.assert_allclose(cls_exp, g["cls_exp_tight"], rtol=0, atol=1e-6)
    np.testing.assert_allclose(cls_obs, g["cls_obs"], rtol=0, atol=1e-4)
    obs = batched.upper_limit_from_scan(g["scan"], cls_obs)
    exp = batched.upper_limit_from_scan(g["scan"], cls_exp)
    assert abs(obs - float(g["obs_limit_tight"])) < 1e-6
    np.testing.assert_allclose(exp, g["exp_limits_tight"], rtol=0, atol=1e-6)
    # mu = 0 is the degenerate point of calculators.py:432-445 (sqrt(qmu_A) = 0): CLs = 1, no NaN
    assert cls_obs[0] == 1.0 and np.all(cls_exp[0] == 1.0)


@pytest.mark.parametrize("name", ["hello", "val_1bin_normsys", "val_2bin_histosys", "val_2bin_2channel_couplednorm"])
def test_single_point_scan_matches_hypotest_goldens(name, golden, cuda):
    from pyhf_b200 import batched

    g = golden(name)
    ts = g.meta.get("test_stat", "qtilde")
    key = "cls_obs_qtilde_tight" if name == "hello" else "cls_obs_tight"
    if key not in g:
        pytest.skip("no tight CLs stored for this model")
    plan = _plan(g, cuda)
    out = batched.asymptotic_scan(plan, [g.meta["mu_test"]], g["data"], test_stat=ts)
    assert out["n_failed"] == 0
    assert abs(float(out["CLs_obs"][0]) - float(g[key])) < 1e-6
    ekey = "cls_exp_qtilde_tight" if name == "hello" else "cls_exp_tight"
    np.testing.assert_allclose(out["CLs_exp"][0].cpu().numpy(), np.asarray(g[ekey]).ravel(), rtol=0, atol=1e-6)


def test_clipped_normal_base_distribution(golden, cuda):
    """calc_base_dist="clipped_normal" (calculators.py:312-341): expected values are capped at
    the cutoff -sqrt(qmu_A), so the +2 sigma band saturates instead of using unphysical values"""
    from pyhf_b200 import batched

    g = golden("hello")
    plan = _plan(g, cuda)
    a = batched.asymptotic_scan(plan, [0.2, 1.0], g["data"])
    b = batched.asymptotic_scan(plan, [0.2, 1.0], g["data"], calc_base_dist="clipped_normal")
    np.testing.assert_allclose(a["CLs_obs"].cpu().numpy(), b["CLs_obs"].cpu().numpy(), rtol=1e-12)
    sqA = np.sqrt(b["qmuA"].cpu().numpy())
    from scipy.stats import norm

    for i in range(2):
        for j, nsig in enumerate((2, 1, 0, -1, -2)):
            t = max(nsig, -sqA[i])
            want = norm.cdf(-(t + sqA[i])) / norm.cdf(-t)
            assert abs(float(b["CLs_exp"][i, j]) - want) < 1e-12
