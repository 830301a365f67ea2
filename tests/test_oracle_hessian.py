"""CPU: the closed-form Hessian of the oracle (oracle/hf_oracle.py::hessian_2nll, the checker of
hf_hess_analytic.cu) against central differences of the oracle's analytic gradient, exact and
Fisher mode, on models that cover every modifier type and both interpolation-code pairs."""
import numpy as np
import pytest

from oracle import hf_oracle as O


def _fd_hessian(op, x, data, h=1e-5):
    P = len(x)
    H = np.zeros((P, P))
    for j in range(P):
        hh = h * max(1.0, abs(x[j]))
        xp, xm = x.copy(), x.copy()
        xp[j] += hh
        xm[j] -= hh
        gp = -2.0 * O.logpdf_grad(op, xp[None], data[None])[1][0]
        gm = -2.0 * O.logpdf_grad(op, xm[None], data[None])[1][0]
        H[:, j] = (gp - gm) / (2.0 * hh)
    return 0.5 * (H + H.T)


@pytest.mark.parametrize("name", ["hello", "mixed_code4p_code4", "mixed_code0_code1", "val_1bin_lumi",
                                  "val_2bin_2channel_coupledshapefactor", "val_2bin_2channel_coupledhistosys",
                                  "staterror_zero"])
def test_closed_form_hessian_of_the_oracle(name, golden):
    g = golden(name)
    op, hp = g.oracle_plan(), g.host_plan()
    rng = np.random.default_rng(7)
    free = ~hp.fixed.astype(bool)
    width = np.minimum(0.05 * (hp.hi - hp.lo), 0.4)
    alpha = np.zeros(hp.n_pars, bool)
    alpha[hp.hs_par] = True
    alpha[hp.sf_par[hp.sf_kind != 0]] = True
    data = np.asarray(g["data"], float)
    for trial in range(3):
        x = hp.init + rng.normal(0, 1, hp.n_pars) * width * free
        x = np.clip(x, hp.lo + 0.02, hp.hi - 0.02)
        if trial == 1:
            x[alpha & free] = 1.4     # extrapolation branches
        x[alpha & free] = np.where(np.abs(x[alpha & free]) < 0.05, 0.2, x[alpha & free])  # off the code0/1 kink
        for mode in ("exact", "fisher"):
            d = data if mode == "exact" else O.expected_data(op, x[None])[0]
            want = _fd_hessian(op, x, d)
            got = O.hessian_2nll(op, x, data, mode)
            dd = np.sqrt(np.abs(np.diag(want))) + 1e-12
            err = (np.abs(got - want) / np.outer(dd, dd))[np.ix_(free, free)]
            assert err.max() < 2e-9, (name, trial, mode, err.max())


def test_multiplicative_parameter_on_zero():
    """the POI exactly at 0 (its bound in q~ fits): evaluated at 2^-100, every entry is the limit"""
    from conftest import Golden, GOLDEN
    import os

    g = Golden(os.path.join(GOLDEN, "mixed_code4p_code4.npz"))
    op, hp = g.oracle_plan(), g.host_plan()
    data = np.asarray(g["data"], float)
    x0 = hp.init.copy()
    x0[hp.poi_index] = 0.0
    x1 = x0.copy()
    x1[hp.poi_index] = 1e-9
    H0, H1 = O.hessian_2nll(op, x0, data), O.hessian_2nll(op, x1, data)
    assert np.isfinite(H0).all()
    d = np.sqrt(np.abs(np.diag(H1))) + 1e-12
    assert (np.abs(H0 - H1) / np.outer(d, d)).max() < 1e-7


def test_several_histosys_on_one_sample():
    """C4's structure in small: every background sample carries three histosys (additive deltas:
    the mixed second derivative of a rate in two of them is zero) next to normsys and staterror."""
    from pyhf_b200 import synth
    from pyhf_b200.compile import compile_spec

    hp = compile_spec(synth.synth_spec(2, 4, 2, 3, 2, 1, seed=5), poi_name="mu")
    op = O.Plan(hp.arrays())
    rng = np.random.default_rng(3)
    data = O.expected_data(op, hp.init[None])[0]
    data[: hp.n_bins] = rng.poisson(data[: hp.n_bins])
    free = ~hp.fixed.astype(bool)
    for trial in range(2):
        x = np.clip(hp.init + rng.normal(0, 0.3, hp.n_pars) * free, hp.lo + 0.02, hp.hi - 0.02)
        for mode in ("exact", "fisher"):
            d = data if mode == "exact" else O.expected_data(op, x[None])[0]
            want = _fd_hessian(op, x, d)
            got = O.hessian_2nll(op, x, data, mode)
            dd = np.sqrt(np.abs(np.diag(want))) + 1e-12
            err = (np.abs(got - want) / np.outer(dd, dd))[np.ix_(free, free)]
            assert err.max() < 1e-8, (trial, mode, err.max())


def test_histosys_delta_cancelling_the_nominal():
    """base = nominal + delta exactly 0 in one cell (the reference's coupled-histosys validation model
    at alpha = -1.6): the closed form stays finite"""
    from conftest import Golden, GOLDEN
    import os

    g = Golden(os.path.join(GOLDEN, "val_2bin_2channel_coupledhistosys.npz"))
    op, hp = g.oracle_plan(), g.host_plan()
    x = hp.init.copy()
    x[hp.hs_par] = -1.6
    H = O.hessian_2nll(op, x, np.asarray(g["data"], float))
    assert np.isfinite(H).all()
