"""quick GPU probe: fp64 pipe peak, K1/K2 timing on C4, K3 timing on C3 (not the bench contract)"""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "baseline", "_ref"))
import numpy as np
import torch

import pyhf
from pyhf_b200 import batched, synth
from pyhf_b200.plan import extract_plan


def timed(fn, n=3):
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    best = 1e30
    for _ in range(n):
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    return best


print("fp64 peak TFLOP/s:", batched.fp64_peak_tflops())
which = sys.argv[1:] or ["c4", "c3"]
if "c4" in which:
    t0 = time.time()
    model = pyhf.Model(synth.spec_c4(), poi_name="mu", validate=False)
    hp = extract_plan(model)
    plan = batched.DevicePlan(hp)
    print("c4 model+plan build s:", time.time() - t0)
    is_alpha = np.zeros(hp.n_pars, bool); is_alpha[hp.hs_par] = True; is_alpha[hp.sf_par[hp.sf_kind != 0]] = True
    is_gamma = np.zeros(hp.n_pars, bool); is_gamma[hp.bf_par[hp.bf_par >= 0]] = True
    np.random.seed(0)
    data = np.asarray(model.make_pdf(pyhf.tensorlib.astensor(hp.init)).sample((1,))[0])
    for B in (4096, 16384, 65536):
        pars = torch.as_tensor(synth.param_batch(hp.init, hp.lo, hp.hi, is_alpha, is_gamma, hp.poi_index, B, 0), device="cuda")
        d = torch.as_tensor(data, device="cuda")
        plan.logpdf_grad(pars, d)
        ms = timed(lambda: plan.logpdf_grad(pars, d))
        ms1 = timed(lambda: plan.logpdf(pars, d))
        print(f"c4 B={B}: logpdf+grad {ms:.2f} ms -> {B/ms*1e3:.3e} evals/s ({B/ms*1e3*3.5e6/1e12:.2f} TFLOP/s alg) | logpdf {ms1:.2f} ms -> {B/ms1*1e3:.3e}/s")
if "c3" in which:
    model = pyhf.Model(synth.spec_c3(), poi_name="mu", validate=False)
    hp = extract_plan(model)
    plan = batched.DevicePlan(hp)
    for n in (1000, 20000):
        toys = plan.sample_toys(hp.init, n, seed=1)
        torch.cuda.synchronize(); t0 = time.time()
        out = batched.qmu_tilde_batch(plan, 1.0, toys)
        torch.cuda.synchronize(); dt = time.time() - t0
        st = torch.cat([out["fixed"].status, out["free"].status])
        print(f"c3 {n} toys: {dt:.2f}s -> {2*n/dt:.1f} fits/s; failed {(st != 0).sum().item()} ; mean iters {out['free'].niter.float().mean().item():.1f}; q mean {out['q'].mean().item():.4f}")
