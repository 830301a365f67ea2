"""The pyhf-free plan compiler (pyhf_b200/compile.py) must produce, table for table, what
``extract_plan`` read off the live reference model when the golden fixtures were made
(tests/golden/make_golden.py stores those tables as ``plan__*``)."""
import numpy as np
import pytest

from conftest import golden_names
from pyhf_b200.compile import compile_spec
from pyhf_b200.plan import PlanError

WITH_PLAN = [n for n in golden_names() if n not in ("c4", "c4_fits")]


def _settings(g):
    s = g.meta.get("settings") or {}
    return s.get("modifier_settings", s if "normsys" in s or "histosys" in s else None)


def _assert_same(a, b):
    assert set(a) == set(b)
    for k in sorted(a):
        x, y = np.asarray(a[k]), np.asarray(b[k])
        assert x.shape == y.shape, (k, x.shape, y.shape)
        if x.dtype.kind == "f":
            # code4 coefficients come from a 6x6 matrix product in the reference: same formula,
            # summation order may differ in the last bits
            np.testing.assert_allclose(x, y, rtol=1e-13, atol=1e-15, err_msg=k)
        else:
            np.testing.assert_array_equal(x, y, err_msg=k)


@pytest.mark.parametrize("name", WITH_PLAN)
def test_compiler_matches_golden_plan(golden, name):
    g = golden(name)
    s = g.meta.get("settings") or {}
    clips = {k: s[k] for k in ("clip_sample_data", "clip_bin_data") if k in s}
    plan = compile_spec(g.spec(), poi_name=g.meta.get("poi_name", "mu"), modifier_settings=_settings(g), **clips)
    _assert_same(plan.arrays(), g.plan_arrays())


@pytest.mark.needs_pyhf
@pytest.mark.parametrize("gen", ["spec_c3", "spec_c4", "spec_hello"])
def test_compiler_matches_live_extraction(gen):
    import pyhf

    from pyhf_b200 import synth
    from pyhf_b200.plan import extract_plan

    spec = getattr(synth, gen)()
    ref = extract_plan(pyhf.Model(spec, poi_name="mu", validate=False))
    _assert_same(compile_spec(spec, poi_name="mu").arrays(), ref.arrays())


def test_compiler_rejects_out_of_scope():
    spec = {"channels": [{"name": "c", "samples": [
        {"name": "s", "data": [1.0], "modifiers": [{"name": "mu", "type": "normfactor", "data": None}]}]}]}
    with pytest.raises(PlanError):
        compile_spec(spec, modifier_settings={"normsys": {"interpcode": "code2"}})
    spec["channels"][0]["samples"][0]["modifiers"].append({"name": "x", "type": "custom", "data": None})
    with pytest.raises(PlanError):
        compile_spec(spec)


@pytest.mark.needs_pyhf
@pytest.mark.parametrize("case", ["synthetic", "reference_example"])
def test_compiler_matches_live_extraction_on_signal_group_models(case):
    """the group model of SURVEY.md section 8 row f4 (background once, K signal variants with a
    POI each; pyhf_b200/patchset.py) and every per-patch model: the tables compiled from the spec
    equal what the reference's own model builder produces for it"""
    import json
    import os

    import pyhf

    from pyhf_b200 import patchset as PS
    from pyhf_b200.plan import extract_plan

    c = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "patchset.json")))[case]
    patched = [PS.apply_patch(c["workspace"], p) for p in c["patches"]]
    specs, pois = zip(*[PS._model_spec(w, 0) for w in patched])
    group = PS.build_signal_group(list(specs), pois[0])
    ref = extract_plan(pyhf.Model(group.spec, poi_name=group.poi_names[-1]))
    mine = compile_spec(group.spec, poi_name=group.poi_names[-1])
    _assert_same(mine.arrays(), ref.arrays())
    # parameter bookkeeping the mask table is built from
    model = pyhf.Model(group.spec, poi_name=group.poi_names[0])
    for name, (start, n) in mine.par_slices.items():
        sl = model.config.par_slice(name)
        assert (sl.start, sl.stop - sl.start) == (start, n), name
    for spec, poi in zip(specs, pois):
        _assert_same(compile_spec(spec, poi_name=poi).arrays(),
                     extract_plan(pyhf.Model(spec, poi_name=poi)).arrays())


def test_compiler_enforces_the_reference_builder_checks():
    """pdf.py:116-135 (a shapesys name may not be reused: is_shared = False),
    parameters/utils.py:26-62 (all requirement properties of a shared name must agree; a modifier
    type that does not use an attribute rejects a user value for it)"""
    def chan(name, mods_b):
        return {"name": name, "samples": [
            {"name": "s", "data": [5.0, 6.0], "modifiers": [{"name": "mu", "type": "normfactor", "data": None}]},
            {"name": "b", "data": [50.0, 60.0], "modifiers": mods_b}]}

    shp = {"name": "shp", "type": "shapesys", "data": [3.0, 4.0]}
    with pytest.raises(PlanError, match="other paramsets exist"):
        compile_spec({"channels": [chan("c1", [shp]), chan("c2", [dict(shp)])]})
    ns = {"name": "sys", "type": "normsys", "data": {"hi": 1.1, "lo": 0.9}}
    spec = {"channels": [chan("c1", [ns])],
            "parameters": [{"name": "sys", "sigmas": [2.0]}]}
    with pytest.raises(PlanError, match="does not use the sigmas"):
        compile_spec(spec)
    # same name on a 2-bin shapefactor and a scalar normsys: n_parameters disagree
    sf = {"name": "sys", "type": "shapefactor", "data": None}
    with pytest.raises(PlanError, match="Multiple values"):
        compile_spec({"channels": [chan("c1", [ns, sf])]})
    # histosys + normsys sharing a name is fine (identical requirements)
    hs = {"name": "sys", "type": "histosys", "data": {"hi_data": [55.0, 66.0], "lo_data": [45.0, 54.0]}}
    assert compile_spec({"channels": [chan("c1", [ns, hs])]}).n_pars == 2


@pytest.mark.needs_pyhf
def test_reference_builder_raises_on_the_same_specs():
    """the negative cases above are errors in the unmodified reference too"""
    import pyhf

    def chan(name, mods_b):
        return {"name": name, "samples": [
            {"name": "s", "data": [5.0, 6.0], "modifiers": [{"name": "mu", "type": "normfactor", "data": None}]},
            {"name": "b", "data": [50.0, 60.0], "modifiers": mods_b}]}

    shp = {"name": "shp", "type": "shapesys", "data": [3.0, 4.0]}
    ns = {"name": "sys", "type": "normsys", "data": {"hi": 1.1, "lo": 0.9}}
    sf = {"name": "sys", "type": "shapefactor", "data": None}
    with pytest.raises(pyhf.exceptions.InvalidModel):
        pyhf.Model({"channels": [chan("c1", [shp]), chan("c2", [dict(shp)])]}, poi_name="mu", validate=False)
    with pytest.raises(pyhf.exceptions.InvalidModel):
        pyhf.Model({"channels": [chan("c1", [ns])], "parameters": [{"name": "sys", "sigmas": [2.0]}]},
                   poi_name="mu", validate=False)
    with pytest.raises(pyhf.exceptions.InvalidNameReuse):
        pyhf.Model({"channels": [chan("c1", [ns, sf])]}, poi_name="mu", validate=False)
