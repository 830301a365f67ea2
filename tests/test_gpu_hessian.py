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


@pytest.mark.parametrize("name", ["hello", "mixed_code4p_code4", "mixed_code0_code1", "val_2bin_2channel_couplednorm",
                                  "val_2bin_2channel_coupledhistosys", "val_2bin_2channel_coupledshapefactor",
                                  "val_1bin_lumi", "staterror_zero", "empty_signal_bin", "c3", "c4"])
def test_closed_form_hessian_matches_differences_of_the_gradient(name, golden, cuda):
    """hf_hessian_analytic (the curvature of the minimiser K3, one CTA per point, histosys block on
    DMMA) against hf_hessian (central differences of the analytic gradient): exact mode on the
    data, Fisher mode on the Asimov data of each point; points inside, beyond |alpha| = 1, and with
    the POI exactly on its bound 0."""
    from pyhf_b200.batched import DevicePlan

    if name == "c4":
        pytest.importorskip("pyhf")
    g = golden(name)
    hp = g.host_plan()
    plan = DevicePlan(hp, cuda)
    data = np.asarray(g["data"], float)
    rng = np.random.default_rng(7)
    free = ~hp.fixed.astype(bool)
    B = 6
    width = np.minimum(0.05 * (hp.hi - hp.lo), 0.4)
    pts = hp.init[None] + rng.normal(0, 1, (B, hp.n_pars)) * width[None] * free[None]
    pts = np.clip(pts, hp.lo + 0.02, hp.hi - 0.02)
    alpha = np.zeros(hp.n_pars, bool)
    alpha[hp.hs_par] = True
    alpha[hp.sf_par[hp.sf_kind != 0]] = True
    pts[1, alpha & free] = 1.3       # extrapolation branches
    pts[2, alpha & free] = -1.6
    pts[3, hp.poi_index] = 0.0       # POI on its bound: one-sided differences in the reference matrix
    pts[:, ~free] = hp.init[~free]
    if name == "empty_signal_bin":
        pts[3, hp.poi_index] = 0.3   # (lambda = 0 under n = 0 at mu = 0: the curvature of that bin is 0/0)
    kinked = hp.hs_code == 0 or (hp.sf_kind == 2).any()
    if kinked:                       # stay off the kink at alpha == 0 (the differences straddle it)
        pts[:, alpha & free] = np.where(np.abs(pts[:, alpha & free]) < 0.05, 0.2, pts[:, alpha & free])
    from oracle import hf_oracle as O

    op = g.oracle_plan()
    H_fd = plan.hessian(pts, data).cpu().numpy()
    from pyhf_b200._lib import HFError

    try:
        H_an = plan.hessian(pts, data, method="analytic").cpu().numpy()
    except HFError as exc:
        # outside the closed form's scope (include/hf_b200.h: per-bin parameters acting on several
        # bins, ...): the minimiser falls back to differences of the gradient there
        assert "does not qualify" in str(exc)
        assert name == "val_2bin_2channel_coupledshapefactor"
        return
    asimov = plan.expected_data(pts)
    F_fd = plan.hessian(pts, asimov).cpu().numpy()
    F_an = plan.hessian(pts, data, method="fisher").cpu().numpy()
    fx = ~free
    for got, fd, what in ((H_an, H_fd, "exact"), (F_an, F_fd, "fisher")):
        for k in range(B):
            assert np.isfinite(got[k]).all(), (name, what, k)
            assert np.array_equal(got[k], got[k].T)
            # (a) the numpy closed form of the oracle (log-derivative formulation, pinned against
            #     differences of the oracle gradient to 1e-10 in tests/test_oracle_hessian.py)
            want = O.hessian_2nll(op, pts[k], data, what)
            want[fx, :] = 0.0
            want[:, fx] = 0.0
            want[fx, fx] = 1.0
            # (entries are compared on the scale of their diagonal; a diagonal that is itself 1e-9 of the
            #  largest one -- C4's POI at alpha = 1.3, rates e^12 above the data -- sits at the rounding
            #  floor of the sums and is held to that floor)
            dg = np.abs(np.diag(want))
            d = np.sqrt(dg) + 1e-4 * np.sqrt(dg.max()) + 1e-12
            err = np.abs(got[k] - want) / np.outer(d, d)
            assert err.max() < 1e-9, (name, what, k, err.max(), np.unravel_index(err.argmax(), err.shape))
            # (b) central differences of the CUDA gradient (h = 1e-4: truncation ~1e-5 where third
            #     derivatives are large, one-sided at the POI bound)
            dg = np.abs(np.diag(fd[k]))
            d = np.sqrt(dg) + 1e-4 * np.sqrt(dg.max()) + 1e-12
            err = np.abs(got[k] - fd[k]) / np.outer(d, d)
            assert err.max() < (1e-3 if k == 3 else 2e-4), (name, what, k, err.max())
    # the Fisher matrix is positive semi-definite
    for k in range(B):
        w = np.linalg.eigvalsh(F_an[k][np.ix_(free, free)])
        assert w.min() > -1e-8 * max(1.0, w.max())


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
