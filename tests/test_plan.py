"""CPU: plan extraction from live reference models equals the committed plan tables; the
synthetic generators are reproducible."""
import numpy as np
import pytest

from conftest import SMALL

pytestmark = pytest.mark.needs_pyhf


@pytest.mark.parametrize("name", SMALL)
def test_extract_plan_matches_golden_tables(name, golden):
    from pyhf_b200.plan import extract_plan

    g = golden(name)
    model = g.model()
    fresh = extract_plan(model).arrays()
    stored = g.plan_arrays()
    assert set(fresh) == set(stored)
    for k in fresh:
        assert np.array_equal(np.asarray(fresh[k]), np.asarray(stored[k])), k


def test_plan_independent_of_active_backend(golden):
    import pyhf

    from pyhf_b200.backend import TorchCarrierBackend
    from pyhf_b200.plan import extract_plan

    g = golden("mixed_code4p_code4")
    model = g.model()
    ref = extract_plan(model).arrays()
    pyhf.set_backend(TorchCarrierBackend())
    try:
        model2 = g.model()  # built while a torch backend is active
        for m in (model, model2):
            got = extract_plan(m).arrays()
            for k in ref:
                assert np.array_equal(np.asarray(ref[k]), np.asarray(got[k])), k
    finally:
        pyhf.set_backend("numpy", "scipy")


def test_synthetic_generators():
    import pyhf

    from pyhf_b200 import synth
    from pyhf_b200.plan import extract_plan

    m3 = pyhf.Model(synth.spec_c3(), poi_name="mu", validate=False)
    assert (m3.config.npars, m3.config.nmaindata, m3.config.nauxdata) == (120, 100, 119)
    p3 = extract_plan(m3)
    assert (p3.n_hs, p3.n_sf, p3.n_bf_slots, p3.n_gauss, p3.n_pois) == (0, 77, 1, 19, 100)
    hello = pyhf.Model(synth.spec_hello(), poi_name="mu")
    ref = pyhf.simplemodels.uncorrelated_background([12, 11], [50, 52], [3, 7])
    assert hello.spec == ref.spec


def test_unsupported_models_raise():
    import pyhf

    from pyhf_b200.plan import PlanError, extract_plan
    from pyhf_b200 import synth

    spec = synth.spec_hello()
    with pytest.raises(PlanError):
        extract_plan(pyhf.Model(spec, poi_name="mu", batch_size=2))
    spec2 = {"channels": [{"name": "c", "samples": [
        {"name": "s", "data": [5.0, 6.0], "modifiers": [{"name": "mu", "type": "normfactor", "data": None}]},
        {"name": "b", "data": [50.0, 60.0], "modifiers": [
            {"name": "h", "type": "histosys", "data": {"hi_data": [55.0, 66.0], "lo_data": [45.0, 54.0]}}]}]}]}
    m = pyhf.Model(spec2, poi_name="mu", modifier_settings={"histosys": {"interpcode": "code2"}})
    with pytest.raises(PlanError):
        extract_plan(m)
