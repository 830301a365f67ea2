import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "baseline", "_ref"))
import numpy as np, torch, pyhf
from pyhf_b200 import batched, synth
from pyhf_b200.plan import extract_plan
model = pyhf.Model(synth.spec_c3(), poi_name="mu", validate=False)
hp = extract_plan(model); plan = batched.DevicePlan(hp)
np.random.seed(0)
data = np.asarray(model.make_pdf(pyhf.tensorlib.astensor(hp.init)).sample((1,))[0])
def T():
    torch.cuda.synchronize(); return time.perf_counter()
for n in (256, 12500, 12500, 100000):
    t0 = T()
    obs = batched.qmu_tilde_batch(plan, 1.0, torch.as_tensor(data, device="cuda").reshape(1, -1)); t1 = T()
    fp = plan.fixed.clone(); fp[hp.poi_index] = 1
    i_s = plan.init.clone(); i_b = plan.init.clone(); i_b[hp.poi_index] = 0.0
    gen = plan.fit(torch.as_tensor(data, device="cuda").reshape(1, -1), torch.stack([i_s, i_b]), None, None, fp); t2 = T()
    toys = plan.sample_toys(gen.x[0], n, seed=1); t3 = T()
    q = batched.qmu_tilde_batch(plan, 1.0, toys); t4 = T()
    init_f = plan.init.clone(); init_f[hp.poi_index] = 1.0
    r1 = plan.fit(toys, init_f, None, None, fp); t5 = T()
    r2 = plan.fit(toys); t6 = T()
    print(f"n={n}: obs q {t1-t0:.3f}s, gen fits {t2-t1:.3f}s, sample {t3-t2:.4f}s, qmu_tilde_batch {t4-t3:.3f}s (fixed fit {t5-t4:.3f}s, free fit {t6-t5:.3f}s)")
