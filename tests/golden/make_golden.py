#!/usr/bin/env python
"""Generate the golden fixtures under tests/golden/ by running the UNMODIFIED reference.

Runs only in the build container (needs /root/reference staged by oracle/make_ref.py); the
outputs (*.npz) are committed, the reference is not.  Usage:

    python tests/golden/make_golden.py small      # hello + reference validation models (~1 min)
    python tests/golden/make_golden.py toys       # hello-world toy goldens (test_infer.py:548-637)
    python tests/golden/make_golden.py c3         # C3 logpdf/grad + a few reference fits (minutes)
    python tests/golden/make_golden.py c4         # C4 logpdf/grad (no reference fits: >20 CPU-min each)

Every file stores: the model spec (JSON) or its synthetic-generator name, the plan tables
extracted by pyhf_b200.plan (so GPU-box tests need no pyhf), inputs (pars batch, data) and the
reference outputs (logpdf, expected_data, autograd gradient through the unmodified pyhf graph,
fit results at default and tight SLSQP tolerance, test statistics, CLs).
"""
import ast
import json
import os
import sys
import time

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "baseline", "_ref"))

import pyhf  # noqa: E402  (the reference)
import torch  # noqa: E402

from pyhf_b200 import synth  # noqa: E402
from pyhf_b200.backend import TorchCarrierBackend  # noqa: E402
from pyhf_b200.plan import extract_plan  # noqa: E402

REF_TESTS = "/root/reference/tests"


# ------------------------------------------------------------------ reference fixtures by AST
def load_validation_fixtures():
    """exec the ``source_* / spec_* / expected_result_*`` fixture bodies of the reference's
    tests/test_validation.py (the module itself cannot be imported here: it needs jax+uproot)."""
    path = os.path.join(REF_TESTS, "test_validation.py")
    tree = ast.parse(open(path).read())
    from pathlib import Path
    ns = {"json": json, "pyhf": pyhf, "np": np, "Path": Path, "__builtins__": __builtins__}
    funcs = {}
    for node in tree.body:
        if isinstance(node, ast.FunctionDef) and node.name.split("_")[0] in ("source", "spec", "expected"):
            node.decorator_list = []
            mod = ast.Module(body=[node], type_ignores=[])
            exec(compile(mod, path, "exec"), ns)
            funcs[node.name] = ns[node.name]
    cwd = os.getcwd()
    os.chdir("/root/reference")  # fixtures open validation/data/*.json relatively
    out = {}
    try:
        for name, fn in funcs.items():
            if not name.startswith("source_"):
                continue
            key = name[len("source_"):]
            src = fn()
            spec = funcs[f"spec_{key}"](src)
            exp = funcs[f"expected_result_{key}"]()
            out[key] = (src, spec, exp)
    finally:
        os.chdir(cwd)
    return out


# name -> (mu_test, test_stat, tolerance) : tests/test_validation.py:661-745
VALIDATION_CFG = {
    "1bin_shapesys": (1.0, "q", 1e-6),
    "1bin_shapesys_q0": (0.0, "q0", 3e-4),
    "1bin_lumi": (1.0, "q", 4e-6),
    "1bin_normsys": (1.0, "q", 3e-6),
    "2bin_histosys": (1.0, "q", 8e-5),
    "2bin_2channel": (1.0, "q", 1e-6),
    "2bin_2channel_couplednorm": (1.0, "q", 1e-6),
    "2bin_2channel_coupledhistosys": (1.0, "q", 1e-6),
    "2bin_2channel_coupledshapefactor": (1.0, "q", 2.5e-6),
}


def source_data(src, model):
    if "bindata" in src:
        d = src["bindata"]["data"]
    else:
        d = []
        for c in model.config.channels:
            d += src["channels"][c]["bindata"]["data"]
    return list(d) + list(model.config.auxdata)


# ------------------------------------------------------------------ helpers
def autograd(model, pars, data):
    tb = TorchCarrierBackend()
    pyhf.set_backend(tb)
    try:
        gs = []
        for p in pars:
            t = torch.tensor(p, dtype=torch.float64, requires_grad=True)
            model.logpdf(t, torch.tensor(np.asarray(data, np.float64)))[0].backward()
            gs.append(t.grad.numpy().copy())
    finally:
        pyhf.set_backend("numpy", "scipy")
    return np.array(gs)


def par_batch(model, B, seed):
    rng = np.random.default_rng(seed)
    init = np.array(model.config.suggested_init(), float)
    bounds = np.array(model.config.suggested_bounds(), float)
    fixed = np.array(model.config.suggested_fixed(), bool)
    width = np.minimum(0.25 * (bounds[:, 1] - bounds[:, 0]), 1.5)
    pars = init[None] + rng.normal(0, 1, (B, len(init))) * width[None] * 0.5
    pars = np.clip(pars, bounds[:, 0] + 1e-2, bounds[:, 1] - 1e-2)
    pars[:, fixed] = init[fixed]
    pars[0] = init
    return pars


def eval_block(model, pars, data):
    pyhf.set_backend("numpy", "scipy")
    lp = np.array([float(model.logpdf(p, data)[0]) for p in pars])
    ed = np.array([np.asarray(model.expected_data(p)) for p in pars])
    grad = autograd(model, pars, data)
    # autograd through xlogy(0, 0) is NaN (rate exactly 0 under n = 0) although ln L is finite and
    # one-sidedly differentiable there: such rows carry forward differences of the reference
    # ln L (Richardson, h and h/2) and are flagged in grad_fd (compared to 1e-6, not 1e-12)
    grad_fd = np.zeros(len(pars), bool)
    for r in range(len(pars)):
        if np.isfinite(lp[r]) and not np.all(np.isfinite(grad[r])):
            f = lambda x: float(model.logpdf(x, data)[0])  # noqa: E731
            for j in range(pars.shape[1]):
                h = 1e-4 * max(1.0, abs(pars[r, j]))
                e = np.zeros(pars.shape[1]); e[j] = h
                d1 = (f(pars[r] + e) - lp[r]) / h
                d2 = (f(pars[r] + 0.5 * e) - lp[r]) / (0.5 * h)
                grad[r, j] = 2.0 * d2 - d1
            grad_fd[r] = True
    return dict(pars=pars, data=np.asarray(data, float), logpdf=lp, expected_data=ed,
                grad=grad, grad_fd=grad_fd)


def fit_block(model, data, mu_test, tight_tol=1e-12):
    """reference fits: default SLSQP and tolerance=1e-12 (optimize/opt_scipy.py:79)."""
    out = {}
    for tag, kw in (("default", {}), ("tight", {"tolerance": tight_tol})):
        pyhf.set_backend("numpy", pyhf.optimize.scipy_optimizer(**kw))
        x, f = pyhf.infer.mle.fit(data, model, return_fitted_val=True)
        xf, ff = pyhf.infer.mle.fixed_poi_fit(mu_test, data, model, return_fitted_val=True)
        out[f"free_x_{tag}"] = np.asarray(x)
        out[f"free_fun_{tag}"] = float(f)
        out[f"fixed_x_{tag}"] = np.asarray(xf)
        out[f"fixed_fun_{tag}"] = float(ff)
    pyhf.set_backend("numpy", "scipy")
    return out


def cls_block(model, data, mu, test_stat):
    """asymptotic CLs (observed + expected band) with the default and the tight reference optimizer"""
    out = {}
    for tag, kw in (("", {}), ("_tight", {"tolerance": 1e-12})):
        pyhf.set_backend("numpy", pyhf.optimize.scipy_optimizer(**kw))
        obs, exps = pyhf.infer.hypotest(mu, data, model, test_stat=test_stat, return_expected_set=True)
        out[f"cls_obs{tag}"] = float(obs)
        out[f"cls_exp{tag}"] = np.array([float(e) for e in exps])
    pyhf.set_backend("numpy", "scipy")
    return out


def save(name, spec, model, blocks, extra=None):
    plan = extract_plan(model)
    d = {f"plan__{k}": v for k, v in plan.arrays().items()}
    for k, v in blocks.items():
        d[k] = np.asarray(v)
    meta = dict(name=name, poi_index=model.config.poi_index, par_order=model.config.par_order,
                npars=model.config.npars, numpy=np.__version__,
                scipy=__import__("scipy").__version__)
    meta.update(extra or {})
    d["meta_json"] = np.array(json.dumps(meta))
    if spec is not None:
        d["spec_json"] = np.array(json.dumps(spec))
    path = os.path.join(HERE, f"{name}.npz")
    np.savez_compressed(path, **d)
    print("wrote", path, os.path.getsize(path), "bytes")


# ------------------------------------------------------------------ extra small specs
def spec_staterror():
    """tests/test_pdf.py:200-250 shape + a zero-uncertainty bin (staterror.py:128-133) and a
    zero-yield shapesys bin (shapesys.py:23)."""
    return {
        "channels": [
            {
                "name": "firstchannel",
                "samples": [
                    {"name": "mu", "data": [10.0, 10.0, 4.0],
                     "modifiers": [{"name": "mu", "type": "normfactor", "data": None}]},
                    {"name": "bkg1", "data": [50.0, 70.0, 30.0],
                     "modifiers": [{"name": "stat_firstchannel", "type": "staterror",
                                    "data": [12.0, 12.0, 0.0]}]},
                    {"name": "bkg2", "data": [30.0, 20.0, 10.0],
                     "modifiers": [{"name": "stat_firstchannel", "type": "staterror",
                                    "data": [5.0, 5.0, 0.0]},
                                   {"name": "shp", "type": "shapesys", "data": [3.0, 0.0, 2.0]}]},
                    {"name": "bkg3", "data": [20.0, 15.0, 5.0], "modifiers": []},
                ],
            }
        ]
    }


def spec_mixed():
    """every in-scope modifier type in one 2-channel model, shared histosys+normsys name."""
    return {
        "channels": [
            {"name": "SR", "samples": [
                {"name": "sig", "data": [5.0, 8.0, 3.0], "modifiers": [
                    {"name": "mu", "type": "normfactor", "data": None},
                    {"name": "lumi", "type": "lumi", "data": None},
                    {"name": "sys_sig", "type": "normsys", "data": {"hi": 1.12, "lo": 0.91}}]},
                {"name": "bkgA", "data": [60.0, 45.0, 22.0], "modifiers": [
                    {"name": "lumi", "type": "lumi", "data": None},
                    {"name": "jes", "type": "histosys",
                     "data": {"hi_data": [66.0, 47.0, 25.0], "lo_data": [55.0, 44.0, 20.0]}},
                    {"name": "jes", "type": "normsys", "data": {"hi": 1.05, "lo": 0.96}},
                    {"name": "statSR", "type": "staterror", "data": [4.0, 3.0, 2.5]}]},
                {"name": "bkgB", "data": [12.0, 9.0, 4.0], "modifiers": [
                    {"name": "shapeB", "type": "shapesys", "data": [2.0, 1.5, 1.0]},
                    {"name": "jer", "type": "histosys",
                     "data": {"hi_data": [13.5, 9.2, 4.4], "lo_data": [11.0, 8.5, 3.9]}},
                    {"name": "statSR", "type": "staterror", "data": [1.0, 1.2, 0.8]}]},
            ]},
            {"name": "CR", "samples": [
                {"name": "bkgA", "data": [160.0, 145.0], "modifiers": [
                    {"name": "lumi", "type": "lumi", "data": None},
                    {"name": "jes", "type": "histosys",
                     "data": {"hi_data": [170.0, 150.0], "lo_data": [150.0, 141.0]}},
                    {"name": "kA", "type": "shapefactor", "data": None}]},
                {"name": "bkgB", "data": [30.0, 28.0], "modifiers": [
                    {"name": "jer", "type": "normsys", "data": {"hi": 1.2, "lo": 0.85}}]},
            ]},
        ],
    }


MIXED_MEASUREMENT = {
    "name": "meas", "config": {"poi": "mu", "parameters": [
        {"name": "lumi", "auxdata": [1.0], "sigmas": [0.02], "bounds": [[0.5, 1.5]], "inits": [1.0]}]},
}


def mixed_model(**settings):
    spec = spec_mixed()
    spec["parameters"] = MIXED_MEASUREMENT["config"]["parameters"]
    return spec, pyhf.Model(spec, poi_name="mu", validate=False, **settings)


# ------------------------------------------------------------------ stages
def stage_small():
    pyhf.set_backend("numpy", "scipy")
    # config 1: README hello-world (README.rst:53-63)
    spec = synth.spec_hello()
    model = pyhf.Model(spec, poi_name="mu")
    data = [51.0, 48.0] + model.config.auxdata
    blocks = eval_block(model, par_batch(model, 6, 1), data)
    blocks.update(fit_block(model, data, 1.0))
    cb = cls_block(model, data, 1.0, "qtilde")
    blocks.update({k + "_qtilde" if not k.endswith("_tight") else k.replace("_tight", "_qtilde_tight"): v
                   for k, v in cb.items()})
    blocks["qmu_tilde"] = float(pyhf.infer.test_statistics.qmu_tilde(
        1.0, data, model, model.config.suggested_init(), model.config.suggested_bounds(),
        model.config.suggested_fixed()))
    assert abs(blocks["cls_obs_qtilde"] - 0.05251497) < 1e-7  # README.rst:63
    assert abs(blocks["qmu_tilde"] - 3.938244920380498) < 1e-6  # tests/test_infer.py:635-637
    save("hello", spec, model, blocks, dict(mu_test=1.0, poi_name="mu"))

    # reference validation models with their stored CLs (tests/test_validation.py)
    for key, (src, spec, exp) in load_validation_fixtures().items():
        if key not in VALIDATION_CFG:
            continue
        mu, ts, tol = VALIDATION_CFG[key]
        settings = {"normsys": {"interpcode": "code1"}}  # tests/test_validation.py:774-778
        model = pyhf.Model(spec, modifier_settings=settings, poi_name="mu")
        data = source_data(src, model)
        blocks = eval_block(model, par_batch(model, 6, 2), data)
        blocks.update(fit_block(model, data, mu))
        blocks.update(cls_block(model, data, mu, ts))
        obs = blocks["cls_obs"]
        blocks["known_obs"] = float(exp["obs"])
        blocks["known_exp"] = np.array(exp["exp"], float)
        assert abs(blocks["cls_obs"] - exp["obs"]) / exp["obs"] < max(tol, 2e-6), (key, obs, exp)
        save(f"val_{key}", spec, model, blocks, dict(mu_test=mu, test_stat=ts, tolerance=tol, poi_name="mu",
                  settings={"modifier_settings": settings}))

    # alternative interpolation codes (modifiers/histosys.py:107, normsys.py:75)
    for tag, settings in (
        ("code4p_code4", {}),
        ("code0_code1", {"modifier_settings": {"histosys": {"interpcode": "code0"},
                                               "normsys": {"interpcode": "code1"}}}),
    ):
        spec, model = mixed_model(**settings)
        np.random.seed(3)
        data = np.asarray(model.make_pdf(pyhf.tensorlib.astensor(model.config.suggested_init())).sample((1,))[0])
        pars = par_batch(model, 8, 4)
        # make sure both interpolation branches run: |alpha| > 1 for some rows
        for nm in ("jes", "jer", "sys_sig"):
            sl = model.config.par_map[nm]["slice"]
            pars[1, sl] = 1.7
            pars[2, sl] = -2.3
            pars[3, sl] = 1.0
            pars[4, sl] = -1.0
        blocks = eval_block(model, pars, data)
        blocks.update(fit_block(model, data, 1.0))
        save(f"mixed_{tag}", spec, model, blocks,
             dict(mu_test=1.0, settings=settings, poi_name="mu"))

    spec = spec_staterror()
    model = pyhf.Model(spec, poi_name="mu")
    data = [85.0, 120.0, 47.0] + model.config.auxdata
    blocks = eval_block(model, par_batch(model, 6, 5), data)
    blocks.update(fit_block(model, data, 1.0))
    save("staterror_zero", spec, model, blocks, dict(mu_test=1.0, poi_name="mu"))


def stage_toys():
    """hello-world toys under np.random.seed(0) (tests/test_infer.py:548-637)."""
    pyhf.set_backend("numpy", "scipy")
    spec = synth.spec_hello()
    model = pyhf.Model(spec, poi_name="mu")
    data = [51.0, 48.0] + model.config.auxdata
    init = model.config.suggested_init()
    bounds = model.config.suggested_bounds()
    fixed = model.config.suggested_fixed()
    qt = pyhf.infer.test_statistics.qmu_tilde
    blocks = {}
    # test_emperical_distribution
    np.random.seed(0)
    samples = np.asarray(model.make_pdf(pyhf.tensorlib.astensor(init)).sample((10,)))
    q = np.array([float(qt(1.0, s, model, init, bounds, fixed)) for s in samples])
    known = [0.0, 0.13298492825293806, 0.0, 0.7718560148925349, 1.814884694401428, 0.0, 0.0,
             0.0, 0.0, 0.06586643485326249]
    assert np.allclose(q, known, rtol=1e-7, atol=1e-9), q
    blocks["emp_samples"], blocks["emp_q"] = samples, q
    # test_toy_calculator: s+b sample drawn first, then b-only (calculators.py:797-817)
    np.random.seed(0)
    sig_pars = pyhf.infer.mle.fixed_poi_fit(1.0, data, model, init, bounds, fixed)
    sig = np.asarray(model.make_pdf(pyhf.tensorlib.astensor(sig_pars)).sample((10,)))
    bkg_pars = pyhf.infer.mle.fixed_poi_fit(0.0, data, model, init, bounds, fixed)
    bkg = np.asarray(model.make_pdf(pyhf.tensorlib.astensor(bkg_pars)).sample((10,)))
    qs = np.array([float(qt(1.0, s, model, init, bounds, fixed)) for s in sig])
    qb = np.array([float(qt(1.0, s, model, init, bounds, fixed)) for s in bkg])
    known_s = [0.0, 0.017350013494649374, 0.0, 0.2338008822475217, 0.020328779776718875,
               0.8911134903562186, 0.04408274703718007, 0.0, 0.03977591672014569, 0.0]
    known_b = [5.642956861215396, 0.37581364290284114, 4.875367689039649, 3.4299006094989295,
               1.0161021805475343, 0.03345317321810626, 0.21984803001140563, 1.274869119189077,
               9.368264062021098, 3.0716486684082156]
    assert np.allclose(qs, known_s, rtol=1e-7, atol=1e-9), qs
    assert np.allclose(qb, known_b, rtol=1e-7, atol=1e-9), qb
    blocks.update(sig_pars=np.asarray(sig_pars), bkg_pars=np.asarray(bkg_pars), sig_samples=sig,
                  bkg_samples=bkg, sig_q=qs, bkg_q=qb, data=np.asarray(data, float))
    # a longer injected-toy run: 200 + 200 toys, tight reference q~ for CLs parity
    np.random.seed(1)
    sig2 = np.asarray(model.make_pdf(pyhf.tensorlib.astensor(sig_pars)).sample((200,)))
    bkg2 = np.asarray(model.make_pdf(pyhf.tensorlib.astensor(bkg_pars)).sample((200,)))
    pyhf.set_backend("numpy", pyhf.optimize.scipy_optimizer(tolerance=1e-12))
    blocks["sig2_samples"], blocks["bkg2_samples"] = sig2, bkg2
    blocks["sig2_q_tight"] = np.array([float(qt(1.0, s, model, init, bounds, fixed)) for s in sig2])
    blocks["bkg2_q_tight"] = np.array([float(qt(1.0, s, model, init, bounds, fixed)) for s in bkg2])
    blocks["qobs_tight"] = float(qt(1.0, data, model, init, bounds, fixed))
    pyhf.set_backend("numpy", "scipy")
    save("hello_toys", spec, model, blocks, dict(mu_test=1.0, poi_name="mu"))


def stage_c3(ntoys=3):
    pyhf.set_backend("numpy", "scipy")
    model = pyhf.Model(synth.spec_c3(), poi_name="mu", validate=False)
    init = model.config.suggested_init()
    np.random.seed(0)
    toys = np.asarray(model.make_pdf(pyhf.tensorlib.astensor(init)).sample((ntoys,)))
    data = toys[0]
    blocks = eval_block(model, par_batch(model, 6, 6), data)
    blocks["toys"] = toys
    bounds = model.config.suggested_bounds()
    fixed = model.config.suggested_fixed()
    for tag, kw in (("default", {}), ("tight", {"tolerance": 1e-12})):
        pyhf.set_backend("numpy", pyhf.optimize.scipy_optimizer(**kw))
        res = []
        for t in toys:
            t0 = time.time()
            xf, ff = pyhf.infer.mle.fixed_poi_fit(1.0, t, model, init, bounds, fixed, return_fitted_val=True)
            xu, fu = pyhf.infer.mle.fit(t, model, init, bounds, fixed, return_fitted_val=True)
            res.append((np.asarray(xf), float(ff), np.asarray(xu), float(fu)))
            print(tag, "toy", len(res), time.time() - t0, "s", float(ff) - float(fu), flush=True)
        blocks[f"toy_fixed_x_{tag}"] = np.array([r[0] for r in res])
        blocks[f"toy_fixed_fun_{tag}"] = np.array([r[1] for r in res])
        blocks[f"toy_free_x_{tag}"] = np.array([r[2] for r in res])
        blocks[f"toy_free_fun_{tag}"] = np.array([r[3] for r in res])
    pyhf.set_backend("numpy", "scipy")
    save("c3", None, model, blocks, dict(mu_test=1.0, generator="spec_c3", poi_name="mu"))


def stage_c4():
    pyhf.set_backend("numpy", "scipy")
    model = pyhf.Model(synth.spec_c4(), poi_name="mu", validate=False)
    init = model.config.suggested_init()
    np.random.seed(0)
    data = np.asarray(model.make_pdf(pyhf.tensorlib.astensor(init)).sample((1,))[0])
    plan = extract_plan(model)
    is_alpha = np.zeros(plan.n_pars, bool)
    is_alpha[plan.hs_par] = True
    is_alpha[plan.sf_par[plan.sf_kind != 0]] = True
    is_gamma = np.zeros(plan.n_pars, bool)
    is_gamma[plan.bf_par[plan.bf_par >= 0]] = True
    pars = synth.param_batch(init, plan.lo, plan.hi, is_alpha, is_gamma, plan.poi_index, 65536, seed=0)[:8]
    pars[0] = init
    blocks = eval_block(model, pars, data)
    assert abs(blocks["logpdf"][0] - (-4514.22633127)) < 1e-6
    # C4 plan tables are large (6.4 MB histosys): the fixture keeps inputs/outputs only
    d = {k: np.asarray(v) for k, v in blocks.items()}
    d["meta_json"] = np.array(json.dumps(dict(name="c4", generator="spec_c4", poi_name="mu",
                                              poi_index=model.config.poi_index, npars=model.config.npars)))
    path = os.path.join(HERE, "c4.npz")
    np.savez_compressed(path, **d)
    print("wrote", path, os.path.getsize(path))


def stage_scan():
    """config 2: 41-point asymptotic q~ CLs scan + upper limit on the 2-bin x 2-channel model
    (infer/intervals/upper_limits.py:148-213), default and tight reference optimizer"""
    import numpy as np
    key = "2bin_2channel"
    src, spec, exp = load_validation_fixtures()[key]
    model = pyhf.Model(spec, modifier_settings={"normsys": {"interpcode": "code1"}}, poi_name="mu")
    data = source_data(src, model)
    scan = np.linspace(0, 5, 41)
    d = {"scan": scan, "data": np.asarray(data, float)}
    for tag, kw in (("", {}), ("_tight", {"tolerance": 1e-12})):
        pyhf.set_backend("numpy", pyhf.optimize.scipy_optimizer(**kw))
        obs, exps, (sc, res) = pyhf.infer.intervals.upper_limits.upper_limit(
            data, model, scan, return_results=True)
        d[f"obs_limit{tag}"] = float(obs)
        d[f"exp_limits{tag}"] = np.array([float(e) for e in exps])
        d[f"cls_obs{tag}"] = np.array([float(r[0]) for r in res])
        d[f"cls_exp{tag}"] = np.array([[float(e) for e in r[1]] for r in res])
    pyhf.set_backend("numpy", "scipy")
    d["spec_json"] = np.array(json.dumps(spec))
    d["meta_json"] = np.array(json.dumps(dict(name="scan_2bin_2channel", poi_name="mu",
                                              settings={"modifier_settings": {"normsys": {"interpcode": "code1"}}})))
    np.savez_compressed(os.path.join(HERE, "scan_2bin_2channel.npz"), **d)
    print("obs UL", d["obs_limit"], d["obs_limit_tight"], "exp", d["exp_limits"], d["exp_limits_tight"])



# ------------------------------------------------------------------ edge cases (round 2)
def spec_clipping():
    """the model of the reference's tests/test_pdf.py:1098-1165 (test_pdf_clipping): hi/lo histosys
    templates that drive the expected rates negative at alpha = -0.6"""
    hs = lambda n, hi, lo: {"data": {"hi_data": hi, "lo_data": lo}, "name": n, "type": "histosys"}  # noqa: E731
    return {
        "channels": [
            {"name": "ch1", "samples": [{"name": "signal", "data": [100.0, 100.0], "modifiers": [
                {"data": None, "name": "mu_sig", "type": "normfactor"},
                hs("np_1", [125.0, 75.0], [175.0, 10.0]), hs("np_2", [125.0, 75.0], [175.0, 10.0])]}]},
            {"name": "ch2", "samples": [{"name": "signal", "data": [15000.0], "modifiers": [
                {"data": None, "name": "mu_sig", "type": "normfactor"},
                hs("np_1", [15500.0], [1200.0]), hs("np_2", [15500.0], [1200.0])]}]},
        ]
    }


def spec_empty_signal_bin():
    """a signal-only bin with zero observed events: at mu = 0 (the background-only / Asimov fits of
    every asymptotic CLs) its expected rate is exactly 0 with n = 0 -- xlogy(0, 0) = 0
    (tensor/numpy_backend.py:435-436), and the analytic gradient must stay finite there"""
    return {
        "channels": [
            {"name": "SR", "samples": [
                {"name": "sig", "data": [4.0, 6.0, 3.0], "modifiers": [
                    {"name": "mu", "type": "normfactor", "data": None}]},
                {"name": "bkg", "data": [0.0, 55.0, 30.0], "modifiers": [
                    {"name": "bnorm", "type": "normsys", "data": {"hi": 1.1, "lo": 0.9}},
                    {"name": "bshape", "type": "shapesys", "data": [0.0, 5.0, 4.0]}]},
            ]}
        ]
    }


def stage_edge():
    pyhf.set_backend("numpy", "scipy")
    # ---- clipping: pdf.py:696-699 (by sample), :708-709 (by bin)
    spec = spec_clipping()
    data = [100.0, 100.0, 10.0, 0.0, 0.0]
    for tag, kw in (("clip_sample0", {"clip_sample_data": 0.0}), ("clip_bin0", {"clip_bin_data": 0.0}),
                    ("clip_both", {"clip_sample_data": 20.0, "clip_bin_data": 150.0})):
        model = pyhf.Model(spec, poi_name="mu_sig", validate=False, **kw)
        pars = par_batch(model, 8, 7)
        names = model.config.par_names
        pars[1] = [(-0.6 if "np" in n else 1.0) for n in names]   # the reference test's point
        pars[2] = [(-1.4 if "np" in n else 0.7) for n in names]   # extrapolation branch, clipped too
        pars[3] = [(0.3 if "np" in n else 1.2) for n in names]    # nothing clipped
        blocks = eval_block(model, pars, data)
        if tag != "clip_both":
            assert np.allclose(blocks["expected_data"][1][:3], [1.830496e02, 0.0, 0.0], atol=1e-3)  # test_pdf.py:1187-1198
        blocks.update(fit_block(model, data, 1.0))
        save(tag, spec, model, blocks, dict(mu_test=1.0, poi_name="mu_sig", settings=kw))
    # the all-modifier model with clips that bite on some cells only
    for tag, kw in (("mixed_clip", {"clip_sample_data": 9.0, "clip_bin_data": 60.0}),):
        spec, model = mixed_model(**kw)
        np.random.seed(3)
        data = np.asarray(model.make_pdf(pyhf.tensorlib.astensor(model.config.suggested_init())).sample((1,))[0])
        pars = par_batch(model, 8, 4)
        for nm in ("jes", "jer", "sys_sig"):
            sl = model.config.par_map[nm]["slice"]
            pars[1, sl] = 1.7
            pars[2, sl] = -2.3
        pars[3, model.config.poi_index] = 0.2     # signal cells below the per-sample clip
        blocks = eval_block(model, pars, data)
        save(tag, spec, model, blocks, dict(mu_test=1.0, settings=kw, poi_name="mu"))
    # ---- signal-only bin with zero observed events
    spec = spec_empty_signal_bin()
    model = pyhf.Model(spec, poi_name="mu", validate=False)
    data = [0.0, 58.0, 37.0] + model.config.auxdata
    pars = par_batch(model, 8, 8)
    pars[1, model.config.poi_index] = 0.0        # lambda = 0 with n = 0: value finite, gradient finite
    blocks = eval_block(model, pars, data)
    assert np.all(np.isfinite(blocks["logpdf"]))
    blocks.update(fit_block(model, data, 1.0))
    for mu0 in (0.0,):   # the Asimov-generating fit of AsymptoticCalculator (calculators.py:45-97)
        for tagt, kwt in (("default", {}), ("tight", {"tolerance": 1e-12})):
            pyhf.set_backend("numpy", pyhf.optimize.scipy_optimizer(**kwt))
            x0, f0 = pyhf.infer.mle.fixed_poi_fit(mu0, data, model, return_fitted_val=True)
            blocks[f"fixed0_x_{tagt}"], blocks[f"fixed0_fun_{tagt}"] = np.asarray(x0), float(f0)
    pyhf.set_backend("numpy", "scipy")
    cb = cls_block(model, data, 1.0, "qtilde")
    blocks.update(cb)
    # data with n > 0 in the bin whose rate is 0 at mu = 0: ln L = -inf (tests/test_tensor.py:304-314)
    d2 = [3.0, 58.0, 37.0] + model.config.auxdata
    p0 = np.array(model.config.suggested_init(), float)
    p0[model.config.poi_index] = 0.0
    blocks["neg_inf_data"] = np.asarray(d2, float)
    blocks["neg_inf_pars"] = p0
    blocks["neg_inf_logpdf"] = float(model.logpdf(p0, d2)[0])
    assert blocks["neg_inf_logpdf"] == -np.inf
    save("empty_signal_bin", spec, model, blocks, dict(mu_test=1.0, poi_name="mu", test_stat="qtilde"))



def stage_teststats():
    """every test statistic of infer/test_statistics.py (qmu :63-153, qmu_tilde :156-254, tmu
    :257-341, tmu_tilde :344-436, q0 :439-523) on the hello-world model and on the all-modifier
    model: observed data + 7 toys each, tight reference optimizer.  The unbounded-POI statistics
    (qmu, tmu) use POI bounds [-10, 10] as the reference's docstrings do (:96-97)."""
    import warnings

    ts = pyhf.infer.test_statistics
    out = {}
    specs = {}
    for key in ("hello", "mixed"):
        if key == "hello":
            spec = synth.spec_hello()
            model = pyhf.Model(spec, poi_name="mu")
            data = np.array([51.0, 48.0] + model.config.auxdata)
        else:
            spec, model = mixed_model()
            np.random.seed(3)
            data = np.asarray(model.make_pdf(pyhf.tensorlib.astensor(model.config.suggested_init())).sample((1,))[0])
        specs[key] = spec
        init = model.config.suggested_init()
        fixed = model.config.suggested_fixed()
        b_tilde = model.config.suggested_bounds()
        b_free = [list(b) for b in b_tilde]
        b_free[model.config.poi_index] = [-10.0, 10.0]
        np.random.seed(11)
        gen = np.array(init, float)
        gen[model.config.poi_index] = 0.5
        toys = np.asarray(model.make_pdf(pyhf.tensorlib.astensor(gen)).sample((7,)))
        rows = np.vstack([data[None], toys])
        pyhf.set_backend("numpy", pyhf.optimize.scipy_optimizer(tolerance=1e-12))
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            for name, fn, mu, bounds in (("qmu", ts.qmu, 1.0, b_free), ("qmu_tilde", ts.qmu_tilde, 1.0, b_tilde),
                                         ("tmu", ts.tmu, 1.0, b_free), ("tmu_tilde", ts.tmu_tilde, 1.0, b_tilde),
                                         ("q0", ts.q0, 0.0, b_tilde)):
                vals, muhats = [], []
                for r in rows:
                    v, (fixed_pars, free_pars) = fn(mu, r, model, init, bounds, fixed, return_fitted_pars=True)
                    vals.append(float(v))
                    muhats.append(float(free_pars[model.config.poi_index]))
                out[f"{key}__{name}"] = np.array(vals)
                out[f"{key}__{name}_muhat"] = np.array(muhats)
                out[f"{key}__{name}_lo"] = np.array([b[0] for b in bounds])
                out[f"{key}__{name}_hi"] = np.array([b[1] for b in bounds])
        pyhf.set_backend("numpy", "scipy")
        out[f"{key}__rows"] = rows
        plan = extract_plan(model)
        for k, v in plan.arrays().items():
            out[f"{key}__plan__{k}"] = v
    out["meta_json"] = np.array(json.dumps(dict(name="teststats", models=list(specs))))
    path = os.path.join(HERE, "teststats.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path))
    for k in out:
        if k.count("__") == 1 and not k.endswith(("_lo", "_hi", "rows", "muhat")) and k != "meta_json":
            print(k, np.round(out[k], 4))



if __name__ == "__main__":
    stages = sys.argv[1:] or ["small", "toys", "c4", "c3"]
    for st in stages:
        t0 = time.time()
        globals()[f"stage_{st}"]()
        print(f"stage {st}: {time.time() - t0:.1f}s")
