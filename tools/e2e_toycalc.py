"""unchanged pyhf ToyCalculator under the b200 plugin: wall time incl. the Python toy loop"""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "baseline", "_ref"))
import numpy as np, torch, pyhf, pyhf_b200
from pyhf_b200 import synth
cases = [("hello", synth.spec_hello(), [51, 48], 2000), ("c3", synth.spec_c3(), None, 2000)]
for name, spec, obs, ntoys in cases:
    pyhf.set_backend("numpy", "scipy")
    model = pyhf.Model(spec, poi_name="mu", validate=False)
    if obs is None:
        np.random.seed(0)
        data = list(np.asarray(model.make_pdf(pyhf.tensorlib.astensor(model.config.suggested_init())).sample((1,))[0]))
    else:
        data = obs + model.config.auxdata
    if name == "hello":
        np.random.seed(0); t0 = time.time()
        ref = pyhf.infer.hypotest(1.0, data, model, test_stat="qtilde", calctype="toybased", ntoys=200, track_progress=False)
        print(f"{name}: reference numpy+scipy 200 toys: {time.time()-t0:.2f}s CLs={float(ref):.4f}")
    pyhf.set_backend(pyhf_b200.b200_backend(seed=1), pyhf_b200.b200_optimizer())
    model = pyhf.Model(spec, poi_name="mu", validate=False)
    pyhf.infer.hypotest(1.0, data, model, test_stat="qtilde", calctype="toybased", ntoys=50, track_progress=False)
    torch.cuda.synchronize(); t0 = time.time()
    cls = pyhf.infer.hypotest(1.0, data, model, test_stat="qtilde", calctype="toybased", ntoys=ntoys, track_progress=False)
    torch.cuda.synchronize(); dt = time.time() - t0
    print(f"{name}: b200 plugin, unchanged ToyCalculator, {ntoys} toys: {dt:.2f}s ({4*ntoys/dt:.0f} fits/s incl. python loop) CLs={float(cls):.4f} stats={pyhf.optimizer._stats}")
