"""end-to-end time of hf_logpdf_grad_host on C4 for the chunk schedule in HF_E2E_SCHED (tuning aid):
HF_E2E_SCHED=1,4,8 python tools/e2e_sched.py"""
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

B = 65536
model = pyhf.Model(synth.spec_c4(), poi_name="mu", validate=False)
hp = extract_plan(model)
plan = batched.DevicePlan(hp)
is_alpha = np.zeros(hp.n_pars, bool); is_alpha[hp.hs_par] = True; is_alpha[hp.sf_par[hp.sf_kind != 0]] = True
is_gamma = np.zeros(hp.n_pars, bool); is_gamma[hp.bf_par[hp.bf_par >= 0]] = True
np.random.seed(0)
data = np.asarray(model.make_pdf(pyhf.tensorlib.astensor(hp.init)).sample((1,))[0])
pars_h = torch.as_tensor(synth.param_batch(hp.init, hp.lo, hp.hi, is_alpha, is_gamma, hp.poi_index, B, 0)).pin_memory()
data_h = torch.as_tensor(data).reshape(1, -1).pin_memory()
out_h = torch.empty(B, dtype=torch.float64).pin_memory()
grad_h = torch.empty(B, hp.n_pars, dtype=torch.float64).pin_memory()
for _ in range(3):
    plan.logpdf_grad_host(pars_h, data_h, out_h, grad_h)
torch.cuda.synchronize()
best = 1e9
for rep in range(3):
    t0 = time.perf_counter()
    for _ in range(10):
        plan.logpdf_grad_host(pars_h, data_h, out_h, grad_h)
    best = min(best, (time.perf_counter() - t0) / 10)
print(f"HF_E2E_SCHED={os.environ.get('HF_E2E_SCHED', '(default)')}: {best * 1e3:.3f} ms per call -> {B / best:.4e} evals/s", flush=True)
