"""CPU: pin the oracle (oracle/hf_oracle.py) against the reference.

(a) known answers held by the reference's own tests;
(b) outputs of the unmodified reference run in the build container (tests/golden/*.npz):
    logpdf, expected_data, autograd gradient, fits, q~ values.
"""
import numpy as np
import pytest

from conftest import CLIPPED, FIT_GOLDENS, SMALL
from oracle import hf_oracle as O


@pytest.mark.parametrize("name", SMALL + ["c3"])
def test_oracle_logpdf_expected_grad(name, golden):
    g = golden(name)
    pl = g.oracle_plan()
    v = O.logpdf(pl, g["pars"], g["data"])
    assert np.allclose(v, g["logpdf"], rtol=1e-13, atol=0)
    ed = O.expected_data(pl, g["pars"])
    assert np.allclose(ed, g["expected_data"], rtol=1e-13, atol=1e-300)
    v2, gr = O.logpdf_grad(pl, g["pars"], g["data"])
    assert np.allclose(v2, v, rtol=1e-15)
    hp = g.host_plan()
    err = np.abs(gr - g["grad"]) / (np.abs(g["grad"]).max(axis=1, keepdims=True) + 1e-300)
    finite = np.isfinite(g["logpdf"])  # ln L = -inf (rates clipped to 0 under n > 0): no gradient to compare
    if hp.hs_code == 0 or (hp.sf_kind == 2).any():
        finite[0] = False  # kink of code0/code1 at alpha == 0 (autograd picks abs'(0) = 0)
    fd = g["grad_fd"] if "grad_fd" in g else np.zeros(len(finite), bool)  # rows pinned by finite differences
    assert (finite & ~fd).sum() >= 3 and err[finite & ~fd].max() < 1e-12
    assert not (finite & fd).any() or err[finite & fd].max() < 1e-6


@pytest.mark.needs_pyhf
def test_oracle_c4_golden(golden):
    g = golden("c4")
    pl = g.oracle_plan()
    v, gr = O.logpdf_grad(pl, g["pars"], g["data"])
    assert np.allclose(v, g["logpdf"], rtol=1e-13, atol=0)
    assert abs(v[0] - (-4514.22633127)) < 1e-6  # SURVEY.md section 8(d) probe
    assert (np.abs(gr - g["grad"]) / np.abs(g["grad"]).max(axis=1, keepdims=True)).max() < 1e-12


def test_known_answers_of_reference_tests(golden):
    """tests/test_pdf.py:119-169: uncorrelated_background known pdf / expected_data values"""
    from pyhf_b200.plan import HostPlan

    g = golden("hello")
    pl = g.oracle_plan()
    # README / doctest: expected data at init
    ed = O.expected_data(pl, pl.init)[0]
    assert np.allclose(ed, [62.0, 63.0, 277.77777778, 55.18367347], rtol=1e-9)
    # finite-difference check of the analytic gradient (independent of autograd)
    x = g["pars"][2]
    _, gr = O.logpdf_grad(pl, x, g["data"])
    for j in range(pl.n_pars):
        h = 1e-6
        e = np.zeros(pl.n_pars)
        e[j] = h
        fd = (O.logpdf(pl, x + e, g["data"]) - O.logpdf(pl, x - e, g["data"]))[0] / (2 * h)
        assert abs(fd - gr[0, j]) < 1e-5 * max(1.0, abs(fd))
    # tensor edge cases (tests/test_tensor.py:272-321)
    assert O._poisson_logpdf(np.array([0.0]), np.array([0.0]))[0] == 0.0
    assert O._poisson_logpdf(np.array([1.0]), np.array([0.0]))[0] == -np.inf
    assert np.isclose(O._normal_logpdf(0.5, 0.0, 1.0), -0.5 * np.log(2 * np.pi) - 0.125)
    HostPlan.from_arrays(g.plan_arrays())  # round trip


@pytest.mark.parametrize("name", FIT_GOLDENS)
def test_oracle_tight_fit_matches_reference_fits(name, golden):
    g = golden(name)
    pl = g.oracle_plan()
    x, f = O.fit(pl, g["data"])
    assert f <= float(g["free_fun_default"]) + 1e-6
    if name in CLIPPED:  # kinked likelihood: never a worse minimum than the reference's
        xf, ff = O.fit(pl, g["data"], fixed_vals=[(pl.poi_index, g.meta["mu_test"])])
        assert f <= float(g["free_fun_tight"]) + 1e-6 and ff <= float(g["fixed_fun_tight"]) + 1e-6
        return
    assert abs(f - float(g["free_fun_tight"])) < 1e-6
    xf, ff = O.fit(pl, g["data"], fixed_vals=[(pl.poi_index, g.meta["mu_test"])])
    assert abs(ff - float(g["fixed_fun_tight"])) < 1e-6
    assert xf[pl.poi_index] == g.meta["mu_test"]


def test_hello_qmu_tilde_known_answer(golden):
    g = golden("hello")
    pl = g.oracle_plan()
    q, muhat = O.qmu_tilde(pl, 1.0, g["data"])
    assert abs(q - 3.938244920380498) < 1e-6  # tests/test_infer.py:635-637
    gt = golden("hello_toys")
    for i in range(4):
        q, _ = O.qmu_tilde(pl, 1.0, gt["bkg_samples"][i])
        assert abs(q - gt["bkg_q"][i]) < 1e-6
