"""small end-to-end run for compute-sanitizer: DMMA path, DFMA path, mixed fits, Hessian, scan"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch
from pyhf_b200 import batched, synth
from pyhf_b200.compile import compile_spec

for gen, B in (("spec_c4", 288), ("spec_c3", 40), ("spec_hello", 5)):
    hp = compile_spec(getattr(synth, gen)(), poi_name="mu")
    plan = batched.DevicePlan(hp)
    rng = np.random.default_rng(0)
    pars = np.clip(hp.init + 0.05 * rng.standard_normal((B, hp.n_pars)), hp.lo, hp.hi)
    toys = plan.sample_toys(hp.init, 6, seed=1)
    v, g = plan.logpdf_grad(pars, toys[0])
    assert torch.isfinite(v).all()
    if gen == "spec_c4":
        # per-fit kernel, 256-thread instance: DMMA sweep / Schur / blocked Cholesky / channel blocks
        out = batched.qmu_tilde_batch(plan, 1.0, toys[:int(os.environ.get("HF_SAN_C4_TOYS", "2"))])
        assert int((out["free"].status != 0).sum()) == 0
    else:
        out = batched.qmu_tilde_batch(plan, 1.0, toys)
        assert int((out["free"].status != 0).sum()) == 0
        H = plan.hessian(out["free"].x[:2], toys[:2])
        sc = batched.asymptotic_scan(plan, [0.5, 1.0], toys[0])
        assert torch.isfinite(sc["CLs_obs"]).all() and torch.isfinite(H).all()
    print(gen, "ok", float(v[0]))
