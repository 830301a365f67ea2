"""every signal patch of a patchset on one background workspace,
all fits of a group in two lock-step batches over a mask table (hf_fit_batch_masks), against the
reference's per-patch ``hypotest`` results (tests/golden/patchset.json: unmodified pyhf, numpy
backend, SLSQP at tolerance 1e-12; generator tests/golden/make_patchset_golden.py).
Tolerance: CLs within 1e-6 (BASELINE north star)."""
import json
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

HERE = os.path.dirname(os.path.abspath(__file__))
CASES = json.load(open(os.path.join(HERE, "golden", "patchset.json")))


def _check(out, c, rows=None):
    obs, exp = np.array(c["cls_obs_tight"]), np.array(c["cls_exp_tight"])
    sl = slice(None) if rows is None else rows
    np.testing.assert_allclose(out["CLs_obs"].cpu().numpy(), obs[sl], rtol=0, atol=1e-6)
    np.testing.assert_allclose(out["CLs_exp"].cpu().numpy(), exp[sl], rtol=0, atol=1e-6)
    # the default-tolerance reference is itself only converged to a few 1e-6
    np.testing.assert_allclose(out["CLs_obs"].cpu().numpy(), np.array(c["cls_obs"])[sl], rtol=0, atol=2e-5)


def test_six_signal_points_in_one_group(cuda):
    from pyhf_b200.patchset import patchset_hypotest

    c = CASES["synthetic"]
    K, M = len(c["patches"]), len(c["mus"])
    out = patchset_hypotest(c["workspace"], {"patches": c["patches"]}, c["mus"], device=cuda)
    assert out["mode"] == ["group"] * K
    assert out["n_fits"] == 2 * K * (M + 1) + 1 and out["n_failed"] == 0   # 49 instead of 5 K M = 90
    assert tuple(out["CLs_obs"].shape) == (K, M) and tuple(out["CLs_exp"].shape) == (K, M, 5)
    _check(out, c)


def test_group_size_only_changes_the_batching(cuda):
    from pyhf_b200.patchset import patchset_hypotest

    c = CASES["synthetic"]
    out = patchset_hypotest(c["workspace"], c["patches"], c["mus"], group_size=4, device=cuda)
    assert out["mode"] == ["group"] * 6 and out["n_fits"] == (2 * 4 * 4 + 1) + (2 * 2 * 4 + 1)
    _check(out, c)
    one = patchset_hypotest(c["workspace"], [c["patches"][4]["patch"]], 1.0, device=cuda)
    assert abs(float(one["CLs_obs"][0, 0]) - c["cls_obs_tight"][4][1]) < 1e-6


def test_reference_example_patchset(cuda):
    """tests/test_patchset/example_bkgonly.json + example_patchset.json of the reference: the
    patch replaces the normsys values inside the sample that carries the POI; lumi is fixed."""
    from pyhf_b200.patchset import patchset_hypotest

    c = CASES["reference_example"]
    out = patchset_hypotest(c["workspace"], c["patches"], c["mus"], device=cuda)
    assert out["mode"] == ["group"] and out["n_failed"] == 0
    _check(out, c)


def test_ungroupable_patches_run_one_at_a_time_on_the_same_path(cuda):
    from pyhf_b200.patchset import patchset_hypotest

    c, s = CASES["synthetic_single"], CASES["synthetic"]
    assert c["mus"] == s["mus"][1:]
    patches = s["patches"][:3] + c["patches"]
    out = patchset_hypotest(c["workspace"], patches, c["mus"], device=cuda)
    # the staterror patch cannot be grouped; the patch that edits a background is alone in its bucket
    assert out["mode"] == ["group", "group", "group", "single", "group"]
    assert out["n_failed"] == 0
    np.testing.assert_allclose(out["CLs_obs"][3:].cpu().numpy(), np.array(c["cls_obs_tight"]), rtol=0, atol=1e-6)
    np.testing.assert_allclose(out["CLs_exp"][3:].cpu().numpy(), np.array(c["cls_exp_tight"]), rtol=0, atol=1e-6)
    np.testing.assert_allclose(out["CLs_obs"][:3].cpu().numpy(), np.array(s["cls_obs_tight"])[:3, 1:], rtol=0, atol=1e-6)


def test_argument_errors(cuda):
    from pyhf_b200.patchset import PatchError, patchset_hypotest

    c = CASES["synthetic"]
    with pytest.raises(ValueError):
        patchset_hypotest(c["workspace"], c["patches"], 1.0, test_stat="q0", device=cuda)
    with pytest.raises(ValueError):
        patchset_hypotest(c["workspace"], [], 1.0, device=cuda)
    with pytest.raises(PatchError):
        patchset_hypotest(c["workspace"], [[{"op": "remove", "path": "/channels/7"}]], 1.0, device=cuda)
