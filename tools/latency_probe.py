"""latency of ONE asymptotic hypotest (5 fits) through the unchanged pyhf call, reference vs plugin"""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "baseline", "_ref"))
import numpy as np, torch, pyhf, pyhf_b200
from pyhf_b200 import synth

def run(model, data, n):
    pyhf.infer.hypotest(1.0, data, model, test_stat="qtilde")
    if torch.cuda.is_available(): torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(n):
        r = pyhf.infer.hypotest(1.0, data, model, test_stat="qtilde")
    if torch.cuda.is_available(): torch.cuda.synchronize()
    return (time.perf_counter() - t0) / n, float(r)

for name, spec, nref in (("hello", synth.spec_hello(), 20), ("c3", synth.spec_c3(), 0)):
    pyhf.set_backend("numpy", "scipy")
    model = pyhf.Model(spec, poi_name="mu", validate=False)
    np.random.seed(0)
    data = list(np.asarray(model.make_pdf(pyhf.tensorlib.astensor(model.config.suggested_init())).sample((1,))[0]))
    if nref:
        t, c = run(model, data, nref)
        print(f"{name}: reference {1e3*t:.1f} ms per hypotest, CLs {c:.6f}")
    pyhf.set_backend(pyhf_b200.b200_backend(), pyhf_b200.b200_optimizer())
    model = pyhf.Model(spec, poi_name="mu", validate=False)
    t, c = run(model, data, 20)
    print(f"{name}: b200 plugin {1e3*t:.1f} ms per hypotest, CLs {c:.6f}")
