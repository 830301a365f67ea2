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
