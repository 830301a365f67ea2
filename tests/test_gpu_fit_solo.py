"""K3, per-fit form (hf_fit_solo.cu: one CTA per fit, no host in the loop) against the lock-step
driver (hf_fit.cu) on the same fits: both implement the damped projected Newton iteration of
oracle/fit_proto.py, so they must end in the same minima; and the per-fit form does not depend on
how a batch is composed (a fit alone == the same fit inside a batch, bit for bit)."""
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def _fit_both(plan, data, **kw):
    os.environ.pop("HF_FIT_SOLO", None)
    solo = plan.fit(data, **kw)
    os.environ["HF_FIT_SOLO"] = "0"
    try:
        lock = plan.fit(data, **kw)
    finally:
        os.environ.pop("HF_FIT_SOLO", None)
    return solo, lock


@pytest.mark.parametrize("name", ["hello", "mixed_code4p_code4", "mixed_code0_code1", "val_2bin_2channel_couplednorm",
                                  "val_1bin_lumi", "val_2bin_histosys", "staterror_zero", "empty_signal_bin", "c3"])
def test_per_fit_kernel_matches_lockstep(name, golden, cuda):
    from oracle import hf_oracle as O
    from pyhf_b200.batched import DevicePlan

    g = golden(name)
    hp, op = g.host_plan(), g.oracle_plan()
    plan = DevicePlan(hp, cuda)
    n = 48
    toys = plan.sample_toys(hp.init, n, seed=21)
    toys[0] = torch.as_tensor(np.asarray(g["data"], float), device=cuda)
    fixed = hp.fixed.copy()
    fixed[hp.poi_index] = 1
    for kw in ({}, {"fixed": fixed}):
        solo, lock = _fit_both(plan, toys, **kw)
        assert (solo.status == 0).all() and (lock.status == 0).all(), (solo.status.tolist(), lock.status.tolist())
        fs, fl = solo.fun.cpu().numpy(), lock.fun.cpu().numpy()
        # twice_nll of the returned point is what the oracle computes there
        f_or = O.twice_nll(op, solo.x.cpu().numpy(), toys.cpu().numpy())
        assert np.abs(fs - f_or).max() < 1e-9 * max(1.0, np.abs(f_or).max())
        # same minima (a handful of toys of a multimodal likelihood may pick another basin; never a worse one
        # by more than the other path does)
        same = np.abs(fs - fl) < 1e-6
        assert same.mean() >= 0.9, (name, fs - fl)
        assert np.median(np.abs(fs - fl)) < 1e-7
        if kw:
            assert (solo.x[:, hp.poi_index] == float(hp.init[hp.poi_index])).all()


def test_per_fit_kernel_is_batch_independent(golden, cuda):
    from pyhf_b200.batched import DevicePlan

    g = golden("c3")
    hp = g.host_plan()
    plan = DevicePlan(hp, cuda)
    toys = plan.sample_toys(hp.init, 300, seed=4)
    full = plan.fit(toys)
    part = plan.fit(toys[100:140])
    assert torch.equal(full.x[100:140], part.x)
    assert torch.equal(full.fun[100:140], part.fun)
    assert torch.equal(full.niter[100:140], part.niter)


def test_per_fit_kernel_c4(golden, cuda):
    """C4 (P = 300, 100 histosys on DMMA, 200 x 200 reduced system): first-order optimality of the minima
    with the oracle gradient, and agreement with the lock-step driver"""
    pytest.importorskip("pyhf")
    from conftest import synthetic_host_plan
    from oracle import hf_oracle as O
    from pyhf_b200.batched import DevicePlan

    hp = synthetic_host_plan("spec_c4")
    op = O.Plan(hp.arrays())
    plan = DevicePlan(hp, cuda)
    toys = plan.sample_toys(hp.init, 6, seed=5)
    solo, lock = _fit_both(plan, toys)
    assert (solo.status == 0).all() and (lock.status == 0).all()
    x = solo.x.cpu().numpy()
    v, gr = O.logpdf_grad(op, x, toys.cpu().numpy())
    gr = -2.0 * gr
    assert np.abs(-2.0 * v - solo.fun.cpu().numpy()).max() < 1e-9 * np.abs(v).max()
    free = ~hp.fixed.astype(bool)
    at_lo = (x <= hp.lo + 1e-12) & (gr > 0)
    at_hi = (x >= hp.hi - 1e-12) & (gr < 0)
    pg = np.where(free[None] & ~at_lo & ~at_hi, gr, 0.0)
    assert np.abs(pg).max() < 5e-4, np.abs(pg).max(axis=1)
    d = solo.fun.cpu().numpy() - lock.fun.cpu().numpy()
    assert (np.abs(d) < 1e-6).sum() >= 4, d   # the synthetic C4 likelihood is multimodal
