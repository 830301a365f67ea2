"""Drop-in boundary: unchanged pyhf calls under pyhf.set_backend(b200_backend(), b200_optimizer())."""
import numpy as np
import pytest
import torch

from conftest import SMALL

pytestmark = [pytest.mark.gpu, pytest.mark.needs_pyhf]


@pytest.fixture
def b200(cuda):
    import pyhf

    import pyhf_b200

    pyhf.set_backend(pyhf_b200.b200_backend(seed=7), pyhf_b200.b200_optimizer())
    yield pyhf
    pyhf.set_backend("numpy", "scipy")


def test_readme_hello_world(b200, golden):
    pyhf = b200
    g = golden("hello")
    model = g.model()
    data = [51, 48] + model.config.auxdata
    cls_obs, cls_exp = pyhf.infer.hypotest(1.0, data, model, test_stat="qtilde", return_expected_set=True)
    exp = [float(c) for c in cls_exp]
    # (ii) 1e-6 against the reference stack at tolerance=1e-12 ...
    assert abs(float(cls_obs) - float(g["cls_obs_qtilde_tight"])) < 1e-6
    assert np.allclose(exp, g["cls_exp_qtilde_tight"], atol=1e-6, rtol=0)
    # ... and the published README numbers (default SLSQP tolerance: under-converged at ~3e-5 rel,
    # tight vs default reference differ by exactly that, see tests/golden/make_golden.py)
    assert abs(float(cls_obs) - 0.05251497) < 2e-6  # README.rst:63
    ref = [0.00260626, 0.01382005, 0.06445319, 0.23525641, 0.57303617]
    assert np.allclose(exp, ref, rtol=1e-4, atol=0)


def test_model_logpdf_goes_through_k1(b200, golden):
    pyhf = b200
    from pyhf_b200 import batched

    g = golden("mixed_code4p_code4")
    model = g.model()
    plan = batched.plan_for(model)
    before = plan.launch_count
    val = model.logpdf(g["pars"][1], g["data"])
    assert plan.launch_count > before, "Model.logpdf did not reach the fused kernel"
    assert tuple(val.shape) == (1,)
    assert abs(float(val[0]) - g["logpdf"][1]) < 1e-12 * abs(g["logpdf"][1])
    with pytest.raises(pyhf.exceptions.InvalidPdfParameters):
        model.logpdf(g["pars"][1][:-1], g["data"])
    with pytest.raises(pyhf.exceptions.InvalidPdfData):
        model.logpdf(g["pars"][1], g["data"][:-1])
    # Model.expected_data (pdf.py:886-906) goes through hf_expected_data ...
    before = plan.launch_count
    ed = model.expected_data(g["pars"][1])
    assert plan.launch_count > before, "Model.expected_data did not reach hf_expected_data"
    assert np.allclose(ed.cpu().numpy(), g["expected_data"][1], rtol=1e-12)
    ed_main = model.expected_data(g["pars"][1], include_auxdata=False)
    assert np.allclose(ed_main.cpu().numpy(), g["expected_data"][1][: model.config.nmaindata], rtol=1e-12)
    with pytest.raises(pyhf.exceptions.InvalidPdfParameters):
        model.expected_data(g["pars"][1][:-1])
    # ... and the generic op route (unfused, what make_pdf(...).expected_data() does) agrees
    ed2 = model.make_pdf(pyhf.tensorlib.astensor(g["pars"][1])).expected_data()
    assert np.allclose(ed2.cpu().numpy(), g["expected_data"][1], rtol=1e-12)
    pyhf.set_backend("numpy", "scipy")
    assert not hasattr(pyhf.pdf.Model.logpdf, "__wrapped__"), "hook must uninstall on backend change"
    assert not hasattr(pyhf.pdf.Model.expected_data, "__wrapped__")


@pytest.mark.parametrize("name", [n for n in SMALL if n.startswith("val_")])
def test_validation_cls(name, b200, golden):
    """the reference's own CLs goldens (tests/test_validation.py) through the plugin"""
    pyhf = b200
    g = golden(name)
    model = g.model()
    obs, exp = pyhf.infer.hypotest(g.meta["mu_test"], g["data"].tolist(), model,
                                   test_stat=g.meta["test_stat"], return_expected_set=True)
    # 1e-6 against the tight reference; the stored known answers come from the default-tolerance
    # SLSQP, which the reference's own test only pins to 1e-6 .. 3e-4 relative
    assert abs(float(obs) - float(g["cls_obs_tight"])) < 1e-6
    assert np.allclose([float(e) for e in exp], g["cls_exp_tight"], atol=1e-6, rtol=0)
    tol = max(g.meta["tolerance"], 1e-4)
    assert abs(float(obs) - float(g["known_obs"])) / float(g["known_obs"]) < tol


def test_fit_return_conventions(b200, golden):
    pyhf = b200
    g = golden("hello")
    model = g.model()
    data = g["data"].tolist()
    x = pyhf.infer.mle.fit(data, model)
    assert isinstance(x, torch.Tensor) and tuple(x.shape) == (model.config.npars,)
    x, f = pyhf.infer.mle.fit(data, model, return_fitted_val=True)
    assert f.dim() == 0
    assert abs(float(f) - float(g["free_fun_tight"])) < 1e-6
    x, f, res = pyhf.infer.mle.fixed_poi_fit(1.0, data, model, return_fitted_val=True, return_result_obj=True)
    assert res.success and float(x[model.config.poi_index]) == 1.0
    assert np.abs(x.cpu().numpy() - g["fixed_x_tight"]).max() < 1e-5
    with pytest.raises(pyhf.exceptions.Unsupported):
        pyhf.infer.mle.fit(data, model, method="L-BFGS-B")
    with pytest.raises(pyhf.exceptions.Unsupported):
        pyhf.optimizer.minimize(lambda p, d, m: p.sum(), data, model, model.config.suggested_init(),
                                model.config.suggested_bounds())
    with pytest.raises(ValueError):  # infer/mle.py:57-66 runs before the optimizer
        pyhf.infer.mle.fit(data, model, init_pars=[-5.0, 1.0, 1.0])


def test_unchanged_toycalculator(b200, golden):
    """ToyCalculator loop untouched: the first minimize() of each loop fits the whole toy batch"""
    pyhf = b200
    g = golden("hello")
    model = g.model()
    data = [51, 48] + model.config.auxdata
    opt = pyhf.optimizer
    cls_obs, cls_exp = pyhf.infer.hypotest(1.0, data, model, test_stat="qtilde", calctype="toybased",
                                           ntoys=300, track_progress=False, return_expected_set=True)
    assert opt._stats["batched_fits"] == 4 * 300, opt._stats
    assert opt._stats["cache_hits"] >= 4 * 299
    assert 0.0 <= float(cls_obs) <= 1.0
    # asymptotic value 0.0525; 300 toys -> MC error ~0.02
    assert abs(float(cls_obs) - 0.0525) < 0.08
    assert len(cls_exp) == 5


def test_unchanged_toycalculator_exact_parity(cuda, golden):
    """The reference's own toy-based hypotest, unchanged, under the plugin in exact-parity mode
    (b200_backend(toy_rng="numpy"): toys drawn as the reference's numpy backend draws them,
    tensor/numpy_backend.py:24-51, in the order of calculators.py:797-817): the same
    np.random.seed(0) gives the reference's CLs_obs and expected band (300 + 300 toys, 1 200 toy fits
    in the CUDA minimiser)."""
    import pyhf

    import pyhf_b200

    g = golden("hello")
    gt = golden("hello_toys")
    pyhf.set_backend(pyhf_b200.b200_backend(toy_rng="numpy"), pyhf_b200.b200_optimizer())
    try:
        model = g.model()
        data = [51.0, 48.0] + model.config.auxdata
        np.random.seed(0)
        n = int(gt["toycalc_ntoys"])
        obs, exp = pyhf.infer.hypotest(1.0, data, model, test_stat="qtilde", calctype="toybased", ntoys=n,
                                       track_progress=False, return_expected_set=True)
        assert pyhf.optimizer._stats["batched_fits"] == 4 * n
    finally:
        pyhf.set_backend("numpy", "scipy")
    assert abs(float(obs) - float(gt["toycalc_cls_obs_tight"])) < 1e-6
    assert np.allclose([float(e) for e in exp], gt["toycalc_cls_exp_tight"], atol=1e-6, rtol=0)
    assert abs(float(obs) - float(gt["toycalc_cls_obs"])) < 1e-6   # default-tolerance reference: same counts


def test_user_data_rows_are_not_mistaken_for_a_toy_batch(b200, golden):
    """a row of the user's own data matrix is fitted on its own (only storage the backend's
    samplers created is recognised as a toy batch)"""
    pyhf = b200
    g = golden("hello")
    model = g.model()
    mine = torch.as_tensor(np.tile(g["data"], (5, 1)), device="cuda")
    before = dict(pyhf.optimizer._stats)
    pyhf.infer.mle.fit(mine[2], model)
    assert pyhf.optimizer._stats["batched_fits"] == before["batched_fits"]
    assert pyhf.optimizer._stats["single_fits"] == before["single_fits"] + 1


def test_config2_asymptotic_scan_upper_limit(b200, golden):
    """BASELINE config 2: 41-point q~ CLs scan through the UNCHANGED
    pyhf.infer.intervals.upper_limits.upper_limit (205 fits through b200_optimizer; exercises the
    DeviceArray negative-step slicing of upper_limits.py:16-18), and the same scan as ONE batched
    K3 call with per-row fixed POI values"""
    pyhf = b200
    g = golden("scan_2bin_2channel")
    model = g.model()
    data = g["data"].tolist()
    scan = g["scan"]
    obs, exp, (sc, res) = pyhf.infer.intervals.upper_limits.upper_limit(data, model, scan, return_results=True)
    assert abs(float(obs) - float(g["obs_limit_tight"])) < 1e-6
    assert np.allclose([float(e) for e in exp], g["exp_limits_tight"], atol=1e-6, rtol=0)
    assert abs(float(obs) - 1.0299091) < 5e-6  # SURVEY.md Appendix F probe (default-tolerance reference)
    cls_obs = np.array([float(r[0]) for r in res])
    assert np.abs(cls_obs - g["cls_obs_tight"]).max() < 1e-6
    # batched: all 41 fixed-POI fits in one call equal the per-point fits pyhf just made
    from pyhf_b200 import batched

    plan = batched.plan_for(model)
    hp = plan.host
    init = np.tile(hp.init, (len(scan), 1))
    init[:, hp.poi_index] = scan
    fixed = hp.fixed.copy()
    fixed[hp.poi_index] = 1
    rb = plan.fit(np.asarray(data), init=init, fixed=fixed)
    free = plan.fit(np.asarray(data))
    q = np.maximum(rb.fun.cpu().numpy() - float(free.fun[0]), 0.0)
    q[scan < float(free.x[0, hp.poi_index])] = 0.0
    for i in (8, 20, 40):
        qi = float(pyhf.infer.test_statistics.qmu_tilde(scan[i], data, model, model.config.suggested_init(),
                                                        model.config.suggested_bounds(), model.config.suggested_fixed()))
        assert abs(qi - q[i]) < 1e-6
