"""cProfile of the unchanged pyhf.infer.hypotest (asymptotic, hello-world) under the plugin"""
import os, sys, time, cProfile, pstats
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "baseline", "_ref"))
import numpy as np, torch, pyhf, pyhf_b200
from pyhf_b200 import synth
pyhf.set_backend(pyhf_b200.b200_backend(), pyhf_b200.b200_optimizer())
model = pyhf.Model(synth.spec_hello(), poi_name="mu", validate=False)
data = [51.0, 48.0] + model.config.auxdata
for _ in range(5):
    pyhf.infer.hypotest(1.0, data, model, test_stat="qtilde")
torch.cuda.synchronize()
pr = cProfile.Profile(); pr.enable()
for _ in range(50):
    pyhf.infer.hypotest(1.0, data, model, test_stat="qtilde")
torch.cuda.synchronize(); pr.disable()
st = pstats.Stats(pr); st.sort_stats("cumulative").print_stats(45)
