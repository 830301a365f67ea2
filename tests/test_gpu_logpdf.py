"""K1/K2 parity (through the C-ABI) against the reference goldens and the numpy oracle.

Tolerance: logpdf rtol 1e-12 in fp64 (BASELINE.json north_star); gradients 1e-10 relative to the
largest gradient component (autograd oracle through the unmodified reference graph)."""
import numpy as np
import pytest
import torch

from conftest import SMALL

pytestmark = pytest.mark.gpu

RTOL = 1e-12


def _dev_plan(g, cuda):
    from pyhf_b200.batched import DevicePlan

    return DevicePlan(g.host_plan(), cuda)


@pytest.mark.parametrize("name", SMALL + ["c3"])
def test_logpdf_matches_reference(name, golden, cuda):
    g = golden(name)
    plan = _dev_plan(g, cuda)
    val = plan.logpdf(g["pars"], g["data"]).cpu().numpy()
    ref = g["logpdf"]
    assert np.allclose(val, ref, rtol=RTOL, atol=0), (name, val, ref, np.abs(val - ref) / np.abs(ref))
    # single point, data given per point
    B = len(ref)
    val2 = plan.logpdf(g["pars"], np.tile(g["data"], (B, 1))).cpu().numpy()
    assert np.allclose(val2, ref, rtol=RTOL, atol=0), (name, val2, ref)
    one = plan.logpdf(g["pars"][0], g["data"]).cpu().numpy()
    assert one.shape == (1,) and np.allclose(one, ref[:1], rtol=RTOL)


@pytest.mark.parametrize("name", SMALL + ["c3"])
def test_logpdf_grad_matches_autograd(name, golden, cuda):
    g = golden(name)
    plan = _dev_plan(g, cuda)
    val, grad = plan.logpdf_grad(g["pars"], g["data"])
    val, grad = val.cpu().numpy(), grad.cpu().numpy()
    assert np.allclose(val, g["logpdf"], rtol=RTOL, atol=0), (name, val, g["logpdf"])
    ref = g["grad"]
    scale = np.max(np.abs(ref), axis=1, keepdims=True) + 1e-300
    err = np.abs(grad - ref) / scale
    hp = g.host_plan()
    # rows with ln L = -inf (rates clipped to 0 under n > 0, pdf.py:696-709) have no gradient
    finite = np.isfinite(g["logpdf"])
    if hp.hs_code == 0 or (hp.sf_kind == 2).any():
        # code0/code1 are piecewise with a kink at alpha == 0 (code0.py:70, code1.py:84): row 0 is
        # the suggested init (alpha == 0 exactly) where autograd returns abs'(0) = 0 and the kernel
        # the one-sided derivative of the branch the reference selects; compare off the kink only
        finite[0] = False
    # rows where autograd through xlogy(0, 0) is NaN carry finite differences of the reference
    fd = g["grad_fd"] if "grad_fd" in g else np.zeros(len(finite), bool)
    assert (finite & ~fd).sum() >= 3
    assert err[finite & ~fd].max() < 1e-10, (name, err, np.unravel_index(err.argmax(), err.shape))
    assert not (finite & fd).any() or err[finite & fd].max() < 1e-6
    assert np.isfinite(grad[finite]).all()


@pytest.mark.parametrize("name", ["clip_sample0", "clip_bin0", "clip_both", "mixed_clip", "empty_signal_bin"])
@pytest.mark.parametrize("B", [300, 2500])
def test_clipped_and_empty_bin_models_on_every_point_tile(name, B, golden, cuda):
    """clip_sample_data / clip_bin_data (pdf.py:696-699, :708-709) and a signal-only empty bin through
    the larger point tiles of hf_main_kernel (PT = 8 for B >= 512, 16 for B >= 2368; clipped models
    never take the DMMA kernel): the golden rows tiled to B points, per-point data"""
    g = golden(name)
    plan = _dev_plan(g, cuda)
    reps = -(-B // len(g["logpdf"]))
    pars = np.tile(g["pars"], (reps, 1))[:B]
    ref = np.tile(g["logpdf"], reps)[:B]
    gref = np.tile(g["grad"], (reps, 1))[:B]
    val, grad = plan.logpdf_grad(pars, np.tile(g["data"], (B, 1)))
    val, grad = val.cpu().numpy(), grad.cpu().numpy()
    assert np.allclose(val, ref, rtol=RTOL, atol=0)
    assert np.allclose(plan.logpdf(pars, g["data"]).cpu().numpy(), ref, rtol=RTOL, atol=0)
    ed = plan.expected_data(pars[:16]).cpu().numpy()
    assert np.allclose(ed, np.tile(g["expected_data"], (2, 1))[:16], rtol=1e-12, atol=1e-300)
    fin = np.isfinite(ref)
    hp = g.host_plan()
    fd = np.tile(g["grad_fd"], reps)[:B] if "grad_fd" in g else np.zeros(B, bool)
    scale = np.max(np.abs(gref), axis=1, keepdims=True) + 1e-300
    err = np.abs(grad - gref) / scale
    assert err[fin & ~fd].max() < 1e-10
    assert np.isfinite(grad[fin]).all()


@pytest.mark.parametrize("name", SMALL + ["c3"])
def test_expected_data_matches_reference(name, golden, cuda):
    g = golden(name)
    plan = _dev_plan(g, cuda)
    ed = plan.expected_data(g["pars"]).cpu().numpy()
    ref = g["expected_data"]
    assert np.allclose(ed, ref, rtol=1e-12, atol=1e-300), (name, np.abs(ed - ref).max())


@pytest.mark.parametrize("name", SMALL)
def test_against_oracle_random_batches(name, golden, cuda):
    """odd batch sizes (ragged point tiles), per-point data, parameters on their bounds"""
    from oracle import hf_oracle as O

    g = golden(name)
    hp = g.host_plan()
    op = g.oracle_plan()
    plan = _dev_plan(g, cuda)
    rng = np.random.default_rng(11)
    for B in (1, 3, 17, 130):
        width = np.minimum(0.05 * (hp.hi - hp.lo), 0.6)
        pars = hp.init[None] + rng.normal(0, 1, (B, hp.n_pars)) * width[None]
        pars = np.clip(pars, hp.lo + 0.05, hp.hi)
        pars[:, hp.fixed.astype(bool)] = hp.init[hp.fixed.astype(bool)]
        if hp.poi_index >= 0 and B > 1:
            pars[0, hp.poi_index] = hp.lo[hp.poi_index]  # POI on its bound (mu = 0 in q~ fits)
        lam = O.expected_data(op, pars)
        data = np.where(np.arange(lam.shape[1])[None] < hp.n_bins, rng.poisson(np.maximum(lam, 0)), lam)
        data = data.astype(np.float64)
        v_ref, g_ref = O.logpdf_grad(op, pars, data)
        v, gr = plan.logpdf_grad(pars, data)
        v, gr = v.cpu().numpy(), gr.cpu().numpy()
        fin = np.isfinite(v_ref)
        assert np.array_equal(fin, np.isfinite(v)), (name, B)
        # random points can land on |ln L| << the magnitude of its summands (n ln(lam), lam,
        # lgamma(n+1)); there the honest bound is a few ulp of that magnitude, not of the sum
        from scipy.special import gammaln

        nb = hp.n_bins
        mag = (np.abs(data[:, :nb] * np.log(np.maximum(lam[:, :nb], 1e-300))) + lam[:, :nb]
               + gammaln(data[:, :nb] + 1.0)).sum(axis=1) + np.abs(data[:, nb:]).sum(axis=1) * 8 + 1.0
        tol = np.maximum(RTOL * np.abs(v_ref), 32 * np.finfo(float).eps * mag)
        assert (np.abs(v - v_ref)[fin] <= tol[fin]).all(), (name, B, (np.abs(v - v_ref) / tol)[fin].max())
        scale = np.max(np.abs(g_ref[fin]), axis=1, keepdims=True) + 1e-300
        assert (np.abs(gr[fin] - g_ref[fin]) / scale).max() < 1e-10, (name, B)


def test_poisson_edge_cases(golden, cuda):
    """n = 0 with lambda = 0 contributes 0; n > 0 with lambda = 0 gives -inf
    (reference tests/test_tensor.py:272-321)"""
    g = golden("empty_signal_bin")
    plan = _dev_plan(g, cuda)
    hp = g.host_plan()
    # bin 0 holds signal only: at mu = 0 its rate is exactly 0
    p0 = g["neg_inf_pars"]
    assert p0[hp.poi_index] == 0.0
    lam = plan.expected_data(p0).cpu().numpy()
    assert lam[0] == 0.0
    # n > 0 on lambda = 0  ->  -inf, as the reference returns on the same inputs
    assert float(g["neg_inf_logpdf"]) == -np.inf
    v = plan.logpdf(p0, g["neg_inf_data"]).cpu().numpy()
    assert v[0] == -np.inf
    v, _ = plan.logpdf_grad(p0, g["neg_inf_data"])
    assert float(v[0]) == -np.inf
    # n = 0 on lambda = 0  ->  xlogy(0, 0) = 0: finite value equal to the reference's, finite gradient
    d0 = g["neg_inf_data"].copy()
    d0[0] = 0.0
    assert np.array_equal(d0, g["data"])
    v, gr = plan.logpdf_grad(p0, d0)
    row = int(np.nonzero(g["pars"][:, hp.poi_index] == 0.0)[0][0])
    p_row = g["pars"][row]
    v_row, g_row = plan.logpdf_grad(p_row, d0)
    assert np.isfinite(float(v[0])) and np.isfinite(gr.cpu().numpy()).all()
    assert np.allclose(float(v_row[0]), g["logpdf"][row], rtol=RTOL, atol=0)
    assert np.isfinite(g_row.cpu().numpy()).all()
    # the Poisson constraint term has the same two cases (constraints.py:219-262): aux = 0 on rate 0
    gh = golden("hello")
    ph = _dev_plan(gh, cuda)
    hh = gh.host_plan()
    x = hh.init.copy()
    x[hh.p_par[0]] = 0.0
    d = gh["data"].copy()
    assert ph.logpdf(x, d).cpu().numpy()[0] == -np.inf     # aux > 0 on rate 0
    d[hh.n_bins + hh.p_aux[0]] = 0.0
    vv, gg = ph.logpdf_grad(x, d)
    # (the bin the zeroed gamma multiplies still has lambda > 0 from the signal sample)
    assert np.isfinite(float(vv[0])) and np.isfinite(gg.cpu().numpy()).all()


@pytest.mark.needs_pyhf
def test_c4_goldens_and_large_batch_properties(golden, cuda):
    """C4 (1000 bins / 300 NPs): golden points, then B = 4099 against the oracle on a subsample
    and size-independent properties (row permutation equivariance, batch == singles)."""
    from oracle import hf_oracle as O

    g = golden("c4")
    plan = _dev_plan(g, cuda)
    val, grad = plan.logpdf_grad(g["pars"], g["data"])
    val, grad = val.cpu().numpy(), grad.cpu().numpy()
    assert np.allclose(val, g["logpdf"], rtol=RTOL, atol=0), (val, g["logpdf"])
    scale = np.max(np.abs(g["grad"]), axis=1, keepdims=True)
    assert (np.abs(grad - g["grad"]) / scale).max() < 1e-10
    hp = g.host_plan()
    op = g.oracle_plan()
    from pyhf_b200 import synth

    is_alpha = np.zeros(hp.n_pars, bool)
    is_alpha[hp.hs_par] = True
    is_alpha[hp.sf_par[hp.sf_kind != 0]] = True
    is_gamma = np.zeros(hp.n_pars, bool)
    is_gamma[hp.bf_par[hp.bf_par >= 0]] = True
    B = 4099
    pars = synth.param_batch(hp.init, hp.lo, hp.hi, is_alpha, is_gamma, hp.poi_index, B, seed=3)
    v, gr = plan.logpdf_grad(pars, g["data"])
    sub = np.arange(0, B, 257)
    v_ref, g_ref = O.logpdf_grad(op, pars[sub], g["data"])
    fin = np.isfinite(v_ref)
    assert np.array_equal(fin, np.isfinite(v.cpu().numpy()[sub]))
    assert np.allclose(v.cpu().numpy()[sub][fin], v_ref[fin], rtol=RTOL, atol=0)
    sc = np.max(np.abs(g_ref[fin]), axis=1, keepdims=True)
    assert (np.abs(gr.cpu().numpy()[sub][fin] - g_ref[fin]) / sc).max() < 1e-10
    perm = torch.randperm(B, device=v.device)
    v2, gr2 = plan.logpdf_grad(torch.as_tensor(pars, device=v.device)[perm], g["data"])
    ok = torch.isfinite(v[perm])          # far-out random points may have lambda < 0 -> nan
    assert torch.equal(ok, torch.isfinite(v2)) and ok.float().mean() > 0.9
    rel = ((v2 - v[perm]).abs() / v[perm].abs())[ok].max().item()
    assert rel < 1e-12, rel
    relg = ((gr2 - gr[perm]).abs().max(dim=1).values / gr[perm].abs().max(dim=1).values)[ok].max().item()
    assert relg < 1e-10, relg
    v1 = plan.logpdf(pars[:5], g["data"])
    assert ((v1 - v[:5]).abs() / v[:5].abs()).max().item() < 1e-12


def test_host_buffer_entry_point(golden, cuda):
    g = golden("mixed_code4p_code4")
    plan = _dev_plan(g, cuda)
    pars = torch.as_tensor(g["pars"]).pin_memory()
    data = torch.as_tensor(g["data"]).reshape(1, -1).pin_memory()
    out = torch.empty(pars.shape[0], dtype=torch.float64).pin_memory()
    grad = torch.empty_like(pars).pin_memory()
    plan.logpdf_grad_host(pars, data, out, grad)
    assert np.allclose(out.numpy(), g["logpdf"], rtol=RTOL, atol=0)
    scale = np.max(np.abs(g["grad"]), axis=1, keepdims=True)
    assert (np.abs(grad.numpy() - g["grad"]) / scale).max() < 1e-10


@pytest.mark.parametrize("B", [28417, 40001, 70001])
def test_host_buffer_entry_point_pipelined(B, golden, cuda):
    """above two 3-wave chunks the host entry pipelines H2D / kernels / D2H over a ramped chunk
    schedule (1, 3, 8, 8, ... waves up, 2 and 1 waves down; ragged here, 70001 points have a steady
    chunk in the middle): every point must come back in its own slot, equal to the device-resident call"""
    g = golden("mixed_code4p_code4")
    hp = g.host_plan()
    plan = _dev_plan(g, cuda)
    rng = np.random.default_rng(11)
    pars = hp.init + rng.normal(0, 0.3, (B, hp.n_pars)) * (1 - hp.fixed)
    pars = torch.as_tensor(np.clip(pars, hp.lo + 1e-6, hp.hi)).pin_memory()
    data = torch.as_tensor(g["data"]).reshape(1, -1).pin_memory()
    out = torch.full((B,), np.nan, dtype=torch.float64).pin_memory()
    grad = torch.full((B, hp.n_pars), np.nan, dtype=torch.float64).pin_memory()
    plan.logpdf_grad_host(pars, data, out, grad)
    v, gr = plan.logpdf_grad(pars.to(cuda), g["data"])
    ok = torch.isfinite(v).cpu()
    assert ok.float().mean() > 0.9 and torch.equal(ok, torch.isfinite(out))
    np.testing.assert_allclose(out[ok].numpy(), v.cpu()[ok].numpy(), rtol=1e-13, atol=0)
    np.testing.assert_allclose(grad[ok].numpy(), gr.cpu()[ok].numpy(), rtol=1e-12, atol=1e-12)


@pytest.mark.parametrize("name", ["mixed_code4p_code4", "mixed_code0_code1", "val_2bin_histosys",
                                  "val_2bin_2channel_coupledhistosys"])
@pytest.mark.parametrize("B", [256, 1000])
def test_tensor_core_path_on_small_models(name, B, golden, cuda, monkeypatch):
    """B >= 256 routes histosys models through hf_main_mma_kernel (DMMA + bulk-copy staging):
    padded histosys rows (R < 4), a single K chunk, several per-bin factors per cell, code0,
    ragged last point tile, per-point data; checked against the vector-FMA kernel and the oracle"""
    from oracle import hf_oracle as O

    g = golden(name)
    hp, op = g.host_plan(), g.oracle_plan()
    assert hp.n_hs > 0
    plan = _dev_plan(g, cuda)
    rng = np.random.default_rng(B)
    width = np.minimum(0.05 * (hp.hi - hp.lo), 0.6)
    pars = hp.init[None] + rng.normal(0, 1, (B, hp.n_pars)) * width[None]
    pars = np.clip(pars, hp.lo + 0.05, hp.hi)
    pars[:, hp.fixed.astype(bool)] = hp.init[hp.fixed.astype(bool)]
    pars[0, hp.poi_index] = hp.lo[hp.poi_index]
    lam = O.expected_data(op, pars)
    data = np.where(np.arange(lam.shape[1])[None] < hp.n_bins, rng.poisson(np.maximum(lam, 0)), lam).astype(float)
    monkeypatch.setenv("HF_MMA", "1")
    v1, g1 = plan.logpdf_grad(pars, data)
    v1s = plan.logpdf(pars, data[0])
    monkeypatch.setenv("HF_MMA", "0")
    v0, g0 = plan.logpdf_grad(pars, data)
    v0s = plan.logpdf(pars, data[0])
    v_ref, g_ref = O.logpdf_grad(op, pars, data)
    for v, gr in ((v1, g1), (v0, g0)):
        v, gr = v.cpu().numpy(), gr.cpu().numpy()
        assert np.allclose(v, v_ref, rtol=1e-11, atol=1e-10), np.abs(v - v_ref).max()
        sc = np.max(np.abs(g_ref), axis=1, keepdims=True) + 1e-300
        assert (np.abs(gr - g_ref) / sc).max() < 1e-10
    assert torch.allclose(v1s, v0s, rtol=1e-12, atol=1e-10)


def _spec_two_channels(nb, staterror, shared_shapefactor, shapesys):
    """histosys on every background + per-bin factors in the three forms the DMMA epilogue
    distinguishes: one parameter per bin (interleaved path, plain gradient store), one parameter on a
    bin of EACH channel (shapefactor shared by name: atomic adds), two per-bin kinds on one bin (general path)"""
    rng = np.random.default_rng(5)
    channels = []
    for c in range(2):
        samples = [{"name": "signal", "data": rng.uniform(1, 10, nb).round(3).tolist(),
                    "modifiers": [{"name": "mu", "type": "normfactor", "data": None}]}]
        for s in range(2):
            nom = rng.uniform(20, 200, nb).round(3)
            mods = [{"name": f"hs{k}", "type": "histosys",
                     "data": {"hi_data": (nom * (1 + rng.uniform(0, 0.08, nb))).round(3).tolist(),
                              "lo_data": (nom * (1 - rng.uniform(0, 0.08, nb))).round(3).tolist()}} for k in range(3)]
            mods.append({"name": "ns0", "type": "normsys", "data": {"hi": 1.05 + 0.01 * s, "lo": 0.94}})
            if staterror:
                mods.append({"name": f"stat_ch{c}", "type": "staterror", "data": (0.3 * np.sqrt(nom)).round(3).tolist()})
            if shared_shapefactor and s == 0:
                mods.append({"name": "sf_shared", "type": "shapefactor", "data": None})
            if shapesys and s == 1:
                mods.append({"name": f"shape_ch{c}", "type": "shapesys", "data": (0.1 * nom).round(3).tolist()})
            samples.append({"name": f"bkg{s}", "data": nom.tolist(), "modifiers": mods})
        channels.append({"name": f"ch{c}", "samples": samples})
    return {"channels": channels}


@pytest.mark.parametrize("staterror,shared_shapefactor,shapesys",
                         [(True, False, False), (False, True, False), (True, False, True), (True, True, True), (False, False, True)])
def test_dmma_epilogue_per_bin_factor_forms(staterror, shared_shapefactor, shapesys, cuda):
    """every way a per-bin parameter can reach the gradient in hf_main_mma_kernel (plain store for a
    parameter of one bin, RED when a parameter acts on bins of two channels, the general path when a bin
    carries two per-bin kinds) against the numpy oracle, B = 512 points, 40 bins per channel (tiles
    that mix the forms)"""
    from oracle import hf_oracle as O
    from pyhf_b200.batched import DevicePlan
    from pyhf_b200.compile import compile_spec

    hp = compile_spec(_spec_two_channels(40, staterror, shared_shapefactor, shapesys), poi_name="mu")
    op = O.Plan(hp.arrays())
    plan = DevicePlan(hp, cuda)
    rng = np.random.default_rng(17)
    free = ~hp.fixed.astype(bool)
    lam0 = O.expected_data(op, hp.init[None])[0]
    data = np.concatenate([rng.poisson(lam0[:hp.n_bins]), hp.auxdata]).astype(float)
    data[3] = 0.0  # an empty bin: w = -1 exactly
    x = hp.init[None] + rng.normal(0, 0.3, (512, hp.n_pars)) * free[None]
    x = np.clip(x, hp.lo + 1e-3, hp.hi - 1e-3)
    v_ref, g_ref = O.logpdf_grad(op, x, data)
    v, g = plan.logpdf_grad(x, data)
    v, g = v.cpu().numpy(), g.cpu().numpy()
    assert np.isfinite(v_ref).all()
    np.testing.assert_allclose(v, v_ref, rtol=1e-12, atol=1e-11)
    sc = np.max(np.abs(g_ref), axis=1, keepdims=True)
    assert (np.abs(g - g_ref) / sc).max() < 1e-10
    np.testing.assert_allclose(plan.logpdf(x, data).cpu().numpy(), v_ref, rtol=1e-12, atol=1e-11)
