"""CPU: the numpy prototype of the K3 minimiser (oracle/fit_proto.py, which hf_fit.cu transcribes)
reproduces the reference's scipy fits."""
import numpy as np
import pytest

from conftest import CLIPPED, FIT_GOLDENS
from oracle import fit_proto as F


@pytest.mark.parametrize("name", [n for n in FIT_GOLDENS if n not in CLIPPED])
def test_proto_matches_reference_fits(name, golden):
    g = golden(name)
    pl = g.oracle_plan()
    x, f, st, _, _ = F.fit_batch(pl, g["data"], pl.init, pl.lo, pl.hi, pl.fixed)
    assert st[0] == 0
    assert f[0] <= float(g["free_fun_default"]) + 1e-6
    assert abs(f[0] - float(g["free_fun_tight"])) < 1e-6
    fx = pl.fixed.astype(bool).copy()
    fx[pl.poi_index] = True
    init = pl.init.copy()
    init[pl.poi_index] = g.meta["mu_test"]
    x2, f2, st2, _, _ = F.fit_batch(pl, g["data"], init, pl.lo, pl.hi, fx)
    assert st2[0] == 0 and abs(f2[0] - float(g["fixed_fun_tight"])) < 1e-6


def test_proto_c3_toys(golden):
    g = golden("c3")
    pl = g.oracle_plan()
    x, f, st, ni, info = F.fit_batch(pl, g["toys"], pl.init, pl.lo, pl.hi, pl.fixed)
    assert (st == 0).all()
    assert np.abs(f - g["toy_free_fun_tight"]).max() < 1e-6
    assert (f <= g["toy_free_fun_default"] + 1e-6).all()
