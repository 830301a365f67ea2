"""C3 (no histosys: vector-FMA kernel) logpdf+grad throughput at a large batch -- this is the
evaluation every finite-difference Hessian of a C3 toy fit goes through"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch
from pyhf_b200 import batched, synth
from pyhf_b200.compile import compile_spec

B = int(sys.argv[1]) if len(sys.argv) > 1 else 262144
hp = compile_spec(synth.spec_c3(), poi_name="mu")
plan = batched.DevicePlan(hp)
rng = np.random.default_rng(0)
pars = torch.as_tensor(np.clip(hp.init + 0.3 * rng.standard_normal((B, hp.n_pars)), hp.lo + 1e-3, hp.hi), device="cuda")
toys = plan.sample_toys(hp.init, B, seed=1)
for rows in (toys[:1], toys):
    plan.logpdf_grad(pars, rows)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5):
        plan.logpdf_grad(pars, rows)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 5
    print(f"C3 B={B} data_rows={rows.shape[0]}: {ms:.3f} ms -> {B / ms * 1e3:.3e} evals/s")
