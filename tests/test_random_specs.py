"""CPU, needs the staged reference: randomly structured HistFactory specs (samples missing from
channels, modifiers shared across samples / channels / types, zero-yield and zero-uncertainty
bins, lumi with a parameter configuration) -- the pyhf-free plan compiler must reproduce the
reference model builder's tables, and the numpy oracle evaluated on those tables must reproduce
``Model.logpdf`` / ``Model.expected_data`` of the unmodified reference (rtol 1e-12)."""
import numpy as np
import pytest

from oracle import hf_oracle as O
from pyhf_b200.compile import compile_spec
from test_compile import _assert_same

pytestmark = pytest.mark.needs_pyhf


def random_spec(rng):
    n_ch = int(rng.integers(1, 4))
    sample_names = ["sig"] + [f"bkg{i}" for i in range(int(rng.integers(1, 4)))]
    shared_ns = [f"ns{i}" for i in range(int(rng.integers(0, 3)))]
    shared_hs = [f"hs{i}" for i in range(int(rng.integers(0, 3)))]
    use_lumi = bool(rng.integers(0, 2))
    channels = []
    for c in range(n_ch):
        nb = int(rng.integers(1, 5))
        samples = []
        for s in sample_names:
            if s != "sig" and len(sample_names) > 2 and rng.random() < 0.25:
                continue  # this background is absent from the channel
            nom = np.round(rng.uniform(5, 80, nb), 3)
            if s != "sig" and rng.random() < 0.2:
                nom[int(rng.integers(0, nb))] = 0.0  # a zero-yield bin
            mods = []
            if s == "sig":
                mods.append({"name": "mu", "type": "normfactor", "data": None})
            for n in shared_ns:
                if rng.random() < 0.6:
                    mods.append({"name": n, "type": "normsys",
                                 "data": {"hi": float(np.round(rng.uniform(1.02, 1.2), 4)),
                                          "lo": float(np.round(rng.uniform(0.8, 0.98), 4))}})
            for n in shared_hs:
                if rng.random() < 0.6:
                    mods.append({"name": n, "type": "histosys",
                                 "data": {"hi_data": np.round(nom * rng.uniform(1.0, 1.15, nb), 3).tolist(),
                                          "lo_data": np.round(nom * rng.uniform(0.85, 1.0, nb), 3).tolist()}})
            if s != "sig" and rng.random() < 0.4:
                unc = np.round(0.1 * nom * rng.uniform(0.5, 1.5, nb), 3)
                if rng.random() < 0.3:
                    unc[int(rng.integers(0, nb))] = 0.0
                mods.append({"name": f"shape_{s}_c{c}", "type": "shapesys", "data": unc.tolist()})
            if s != "sig" and rng.random() < 0.5:
                mods.append({"name": f"staterror_c{c}", "type": "staterror",
                             "data": np.round(0.3 * np.sqrt(nom), 3).tolist()})
            if s != "sig" and rng.random() < 0.2:
                mods.append({"name": f"free_c{c}", "type": "shapefactor", "data": None})
            if use_lumi and rng.random() < 0.7:
                mods.append({"name": "lumi", "type": "lumi", "data": None})
            if s != "sig" and rng.random() < 0.2:
                mods.append({"name": "kfac", "type": "normfactor", "data": None})
            samples.append({"name": s, "data": nom.tolist(), "modifiers": mods})
        if not any(smp["name"] == "sig" for smp in samples):
            continue
        channels.append({"name": f"ch{c}", "samples": samples})
    pars = [{"name": "mu", "bounds": [[0.0, 8.0]], "inits": [1.0]}]
    if use_lumi:
        pars.append({"name": "lumi", "auxdata": [1.0], "sigmas": [0.03], "bounds": [[0.7, 1.3]], "inits": [1.0]})
    return {"channels": channels, "parameters": pars}


@pytest.mark.parametrize("seed", range(24))
def test_random_spec_tables_and_logpdf(seed):
    import pyhf

    from pyhf_b200.plan import extract_plan

    pyhf.set_backend("numpy")
    rng = np.random.default_rng(1000 + seed)
    spec = random_spec(rng)
    settings = None if seed % 3 else {"normsys": {"interpcode": "code1"}, "histosys": {"interpcode": "code0"}}
    model = pyhf.Model(spec, poi_name="mu", validate=False,
                       **({"modifier_settings": settings} if settings else {}))
    ref = extract_plan(model)
    mine = compile_spec(spec, poi_name="mu", modifier_settings=settings)
    _assert_same(mine.arrays(), ref.arrays())
    plan = O.Plan(mine.arrays())
    lo, hi = mine.lo, mine.hi
    data = np.concatenate([rng.poisson(np.maximum(O.expected_data(plan, mine.init[None])[0][:mine.n_bins], 0.1)),
                           mine.auxdata]).astype(float)
    for _ in range(4):
        x = mine.init + rng.normal(0, 0.4, mine.n_pars) * (1 - mine.fixed)
        x = np.clip(x, lo + 1e-3, hi - 1e-3)
        x[mine.fixed.astype(bool)] = mine.init[mine.fixed.astype(bool)]
        want = float(model.logpdf(x.tolist(), data.tolist())[0])
        got = float(O.logpdf(plan, x[None], data[None])[0])
        exp_ref = np.asarray(model.expected_data(x.tolist()))
        np.testing.assert_allclose(O.expected_data(plan, x[None])[0], exp_ref, rtol=1e-12, atol=1e-12)
        if np.isfinite(want):
            assert abs(got - want) <= 1e-12 * max(1.0, abs(want)), (got, want)
        else:
            assert not np.isfinite(got) or got == want
