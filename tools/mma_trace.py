"""phase time stamps of CTA 0 of hf_main_mma_kernel on C4 (HF_MMA_TRACE=1; tuning aid):
HF_MMA_TRACE=1 python tools/mma_trace.py [B]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "baseline", "_ref"))
import numpy as np
import torch

import pyhf
from pyhf_b200 import batched, synth
from pyhf_b200.plan import extract_plan

B = int(sys.argv[1]) if len(sys.argv) > 1 else 65536
model = pyhf.Model(synth.spec_c4(), poi_name="mu", validate=False)
hp = extract_plan(model)
plan = batched.DevicePlan(hp)
is_alpha = np.zeros(hp.n_pars, bool); is_alpha[hp.hs_par] = True; is_alpha[hp.sf_par[hp.sf_kind != 0]] = True
is_gamma = np.zeros(hp.n_pars, bool); is_gamma[hp.bf_par[hp.bf_par >= 0]] = True
np.random.seed(0)
data = np.asarray(model.make_pdf(pyhf.tensorlib.astensor(hp.init)).sample((1,))[0])
pars = torch.as_tensor(synth.param_batch(hp.init, hp.lo, hp.hi, is_alpha, is_gamma, hp.poi_index, B, 0), device="cuda")
d = torch.as_tensor(data, device="cuda")
for _ in range(2):
    print("---- launch", file=sys.stderr, flush=True)
    plan.logpdf_grad(pars, d)
    torch.cuda.synchronize()
