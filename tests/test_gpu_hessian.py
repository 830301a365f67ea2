"""Hessian / uncertainties / correlations.

iminuit (the only source of uncertainties in the reference, optimize/opt_minuit.py:136-152) is
not installed in this image, so there are no reference outputs to pin against: the Hessian is
checked against central differences of the ORACLE gradient, the covariance against the closed
form of a counting experiment, and the plugin return conventions against optimize/mixins.py:83-120,
202-209.  (parity with iminuit itself: unpinned.)"""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def _oracle_hessian(op, x, data, free, lo, hi):
    from oracle import hf_oracle as O

    P = len(x)
    H = np.zeros((P, P))
    for j in np.nonzero(free)[0]:
        h = 1e-4 * max(1.0, abs(x[j]))
        xp, xm = x.copy(), x.copy()
        if x[j] + h <= hi[j]:
            xp[j] += h
        if x[j] - h >= lo[j]:
            xm[j] -= h
        gp = -2.0 * O.logpdf_grad(op, xp[None], data[None])[1][0]
        gm = -2.0 * O.logpdf_grad(op, xm[None], data[None])[1][0]
        H[:, j] = (gp - gm) / (xp[j] - xm[j])
    H = 0.5 * (H + H.T)
    fx = ~free
    H[fx, :] = 0.0
    H[:, fx] = 0.0
    H[fx, fx] = 1.0
    return H


@pytest.mark.parametrize("name", ["hello", "mixed_code4p_code4", "val_2bin_2channel_couplednorm", "staterror_zero"])
def test_hessian_matches_oracle_differences(name, golden, cuda):
    from pyhf_b200.batched import DevicePlan

    g = golden(name)
    hp, op = g.host_plan(), g.oracle_plan()
    plan = DevicePlan(hp, cuda)
    data = np.asarray(g["data"], float)
    fit = plan.fit(data)
    x = fit.x[0].cpu().numpy()
    pts = np.stack([x, np.clip(0.5 * (x + hp.init), hp.lo, hp.hi)])
    H = plan.hessian(pts, data).cpu().numpy()
    free = ~hp.fixed.astype(bool)
    for k in range(2):
        want = _oracle_hessian(op, pts[k], data, free, hp.lo, hp.hi)
        scale = np.sqrt(np.outer(np.abs(np.diag(want)), np.abs(np.diag(want)))) + 1e-12
        assert np.max(np.abs(H[k] - want) / scale) < 1e-7, name
        assert np.array_equal(H[k], H[k].T)


def test_counting_experiment_uncertainty_closed_form(cuda):
    """n ~ Pois(mu s + b): sigma_mu = sqrt(n) / s at the minimum"""
    from pyhf_b200.batched import DevicePlan
    from pyhf_b200.compile import compile_spec

    s_, b_, n = 12.0, 50.0, 71.0
    spec = {"channels": [{"name": "c", "samples": [
        {"name": "bkg", "data": [b_], "modifiers": []},
        {"name": "sig", "data": [s_], "modifiers": [{"name": "mu", "type": "normfactor", "data": None}]}]}]}
    plan = DevicePlan(compile_spec(spec, poi_name="mu"), cuda)
    fit = plan.fit(np.array([n]))
    assert abs(float(fit.x[0, 0]) - (n - b_) / s_) < 1e-7
    cov, unc, corr = plan.covariance(fit.x[0], np.array([n]))
    assert abs(float(unc[0]) - np.sqrt(n) / s_) < 1e-7
    assert float(corr[0, 0]) == pytest.approx(1.0, abs=1e-12)


def test_plugin_return_uncertainties_and_correlations(golden, cuda):
    """mle.fit(..., return_uncertainties / return_correlations) under the plugin: layouts of
    optimize/mixins.py:83-120 (x -> [P, 2]) and :202-209 (x, corr, fun order)"""
    from conftest import have_pyhf

    if not have_pyhf():
        pytest.skip("reference pyhf not staged")
    import pyhf

    import pyhf_b200
    from oracle import hf_oracle as O

    g = golden("mixed_code4p_code4")
    pyhf.set_backend(pyhf_b200.b200_backend(), pyhf_b200.b200_optimizer())
    try:
        model = g.model()
        data = np.asarray(g["data"], float).tolist()
        res, corr, fun = pyhf.infer.mle.fit(data, model, return_uncertainties=True, return_correlations=True,
                                            return_fitted_val=True)
        hp, op = g.host_plan(), g.oracle_plan()
        res = np.asarray(pyhf.tensorlib.tolist(res))
        corr = np.asarray(pyhf.tensorlib.tolist(corr))
        assert res.shape == (hp.n_pars, 2) and corr.shape == (hp.n_pars, hp.n_pars)
        x, unc = res[:, 0], res[:, 1]
        assert abs(float(fun) - (-2.0 * O.logpdf(op, x[None], np.asarray(data)[None])[0])) < 1e-8
        free = ~hp.fixed.astype(bool) & (x > hp.lo) & (x < hp.hi)
        H = _oracle_hessian(op, x, np.asarray(data, float), ~hp.fixed.astype(bool), hp.lo, hp.hi)
        cov = 2.0 * np.linalg.inv(H[np.ix_(free, free)])
        np.testing.assert_allclose(unc[free], np.sqrt(np.diag(cov)), rtol=1e-6)
        assert np.all(unc[~free] == 0.0)
        d = np.sqrt(np.diag(cov))
        np.testing.assert_allclose(corr[np.ix_(free, free)], cov / np.outer(d, d), atol=1e-6)
    finally:
        pyhf.set_backend("numpy", "scipy")
