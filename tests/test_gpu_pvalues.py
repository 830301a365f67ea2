"""EmpiricalDistribution / ToyCalculator p-value arithmetic on the
device, O(N log N), against the reference's own arithmetic on the SAME stored toy test statistics
(tests/golden/hello_toys.npz, fields written by tests/golden/add_toy_pvalues.py)."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def test_pvalues_and_expected_band_match_reference(golden, cuda):
    from pyhf_b200 import batched

    g = golden("hello_toys")
    q_sig = torch.as_tensor(g["sig2_q_tight"], device=cuda)
    q_bkg = torch.as_tensor(g["bkg2_q_tight"], device=cuda)
    qobs = torch.as_tensor(g["qobs_tight"], device=cuda).reshape(1)
    n_sb = int(batched.pvalue_counts(q_sig, qobs)[0])
    n_b = int(batched.pvalue_counts(q_bkg, qobs)[0])
    clsb, clb = n_sb / q_sig.numel(), n_b / q_bkg.numel()
    np.testing.assert_allclose([clsb, clb, clsb / clb], g["toys2_pvalues"], rtol=1e-15)
    band = batched.empirical_expected_pvalues(q_sig, q_bkg)
    got = np.stack([b.cpu().numpy() for b in band])
    np.testing.assert_allclose(got, g["toys2_expected_pvalues"], rtol=1e-13, atol=1e-15)
    ev = batched.empirical_expected_value(q_bkg, [-2, -1, 0, 1, 2]).cpu().numpy()
    np.testing.assert_allclose(ev, g["toys2_expected_values_b"], rtol=1e-8)  # normal_cdf differs in the last bits


@pytest.mark.parametrize("n,m", [(1, 1), (1000, 7), (5000, 5000), (70001, 333)])
def test_sorted_and_swept_counts_agree(n, m, cuda):
    """both counting kernels, ties and out-of-range values included (integer results: exact)"""
    from pyhf_b200 import batched

    gen = torch.Generator(device="cpu").manual_seed(n + m)
    samples = torch.randint(0, 50, (n,), generator=gen).double().to(cuda) * 0.25  # many ties
    values = torch.cat([torch.randint(-2, 52, (max(m - 2, 0),), generator=gen).double() * 0.25,
                        torch.tensor([-1e300, 1e300])])[:m].to(cuda)
    want = (samples[None, :] >= values[:, None]).sum(dim=1)
    swept = batched.pvalue_counts(samples, values)
    srt = batched.pvalue_counts(torch.sort(samples).values, values, assume_sorted=True)
    assert torch.equal(swept, want) and torch.equal(srt, want)


def test_toy_hypotest_expected_band_is_consistent(golden, cuda):
    """end to end on the hello-world model: the band from toy_hypotest equals the band recomputed
    with numpy from the toy test statistics it returns (reference formula, calculators.py:955-973)"""
    from pyhf_b200 import batched
    from pyhf_b200.batched import DevicePlan

    g = golden("hello")
    plan = DevicePlan(g.host_plan(), cuda)
    res = batched.toy_hypotest(plan, 1.0, g["data"], ntoys=400, seed=3, return_expected_set=True)
    qs, qb = res["q_sig"].cpu().numpy(), res["q_bkg"].cpu().numpy()
    clsb = np.array([(qs >= t).mean() for t in qb])
    clb = np.array([(qb >= t).mean() for t in qb])
    pct = [2.27501319, 15.86552539, 50.0, 84.13447461, 97.72498681]
    for got, arr in zip(res["expected"], (clsb, clb, clsb / clb)):
        np.testing.assert_allclose(got.cpu().numpy(), np.percentile(arr, pct), rtol=1e-13, atol=1e-15)
