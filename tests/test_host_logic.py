"""CPU checks of host-side arithmetic that needs no device: percentile interpolation (numpy's
"linear" method, reference infer/calculators.py:697-699, 970-972), the scan interpolation of
infer/intervals/upper_limits.py:16-18, shard bookkeeping."""
import numpy as np
import pytest
import torch


@pytest.mark.parametrize("n", [1, 2, 7, 200, 1001])
def test_percentiles_match_numpy_linear(n):
    from pyhf_b200.batched import _NORMAL_PERCENTILES, _percentiles_sorted

    rng = np.random.default_rng(n)
    x = np.sort(rng.exponential(size=n))
    got = _percentiles_sorted(torch.as_tensor(x), _NORMAL_PERCENTILES).numpy()
    np.testing.assert_allclose(got, np.percentile(x, _NORMAL_PERCENTILES), rtol=1e-14, atol=0)
    qs = [0.0, 100.0, 33.3]
    np.testing.assert_allclose(_percentiles_sorted(torch.as_tensor(x), qs).numpy(), np.percentile(x, qs), rtol=1e-14)


def test_upper_limit_interpolation_matches_reference_formula():
    from pyhf_b200.batched import upper_limit_from_scan

    mus = np.linspace(0, 5, 41)
    cls = np.exp(-1.3 * mus)
    want = np.interp(0.05, cls[::-1], mus[::-1])  # upper_limits.py:16-18 on a falling CLs curve
    assert upper_limit_from_scan(mus, cls) == pytest.approx(want, rel=1e-15)
    band = np.stack([cls * f for f in (0.5, 0.8, 1.0, 1.2, 1.5)], axis=1)
    got = upper_limit_from_scan(torch.as_tensor(mus), torch.as_tensor(band))
    assert got == [pytest.approx(np.interp(0.05, band[::-1, i], mus[::-1])) for i in range(5)]


def test_shard_ranges_tile_the_toys():
    from pyhf_b200.dist import shard_range

    for n in (0, 1, 7, 100000, 100003):
        for world in (1, 2, 3, 8):
            parts = [shard_range(n, r, world) for r in range(world)]
            assert parts[0][0] == 0 and sum(c for _, c in parts) == n
            assert all(parts[r][0] + parts[r][1] == parts[r + 1][0] for r in range(world - 1))
            assert max(c for _, c in parts) - min(c for _, c in parts) <= 1


@pytest.mark.needs_pyhf
def test_toy_batches_are_tagged_where_they_are_created():
    """Simultaneous.sample (probability.py:263-274) under the torch carrier: every row of the toy
    matrix is recognised as row i of ONE tagged batch (through concatenate / einsum / gather of
    _TensorViewer.stitch, tensor/common.py:40-52); rows of a user's own matrix are not"""
    import pyhf

    from pyhf_b200 import synth
    from pyhf_b200.backend import TorchCarrierBackend
    from pyhf_b200.optimizer import B200Optimizer

    tb = TorchCarrierBackend()
    pyhf.set_backend(tb)
    try:
        for spec in (synth.spec_hello(), synth.spec_c3()):
            model = pyhf.Model(spec, poi_name="mu", validate=False)
            D = model.config.nmaindata + model.config.nauxdata
            toys = model.make_pdf(tb.astensor(model.config.suggested_init())).sample((7,))
            got = [B200Optimizer._toy_batch_of(r, D) for r in toys]
            assert all(x is not None for x in got)
            assert [x[2] for x in got] == list(range(7))
            assert len({x[0] for x in got}) == 1            # one batch key
            assert torch.equal(got[3][1][3], toys[3])
            user = torch.as_tensor(np.random.rand(5, D))
            assert B200Optimizer._toy_batch_of(user[2], D) is None
    finally:
        pyhf.set_backend("numpy", "scipy")
