"""C4 (P = 300) q~ fits: sweep of the minimiser's Hessian-refresh ratio (HF_FIT_REFRESH, read per
fit launch): a refresh costs n_global + 1 = 201 gradients on C4 against 21 on C3, so the optimum
differs.  One process, plan built once without pyhf.

    python tools/c4_refresh_sweep.py [ntoys] [ratio[:exact_below] ...]   (SWEEP_MODEL=c3 for the C3 model)

``exact_below`` (HF_FIT_EXACT_BELOW): the exact Hessian replaces the Fisher matrix (Hessian on the
Asimov data of the current point, positive semi-definite) once the estimated distance to the
minimum drops below it; 0 = Fisher scoring all the way.
"""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

from pyhf_b200 import batched, synth  # noqa: E402
from pyhf_b200.compile import compile_spec  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
ratios = sys.argv[2:] or ["0.05", "0.2", "0.4", "0.6", "0.8"]
hp = compile_spec(synth.spec_c3() if os.environ.get("SWEEP_MODEL") == "c3" else synth.spec_c4(), poi_name="mu")
plan = batched.DevicePlan(hp)
toys = plan.sample_toys(hp.init, n, seed=5)
batched.qmu_tilde_batch(plan, 1.0, toys[:64])  # warm-up (workspace allocation)
ref = None
for r in ratios:
    f = (r.split(":") + [""])[:2]
    os.environ["HF_FIT_REFRESH"] = f[0]
    os.environ["HF_FIT_EXACT_BELOW"] = f[1] or "10"
    torch.cuda.synchronize()
    t0 = time.time()
    out = batched.qmu_tilde_batch(plan, 1.0, toys)
    torch.cuda.synchronize()
    dt = time.time() - t0
    st = torch.cat([out["fixed"].status, out["free"].status])
    fun = torch.cat([out["fixed"].fun, out["free"].fun])
    if ref is None:
        ref = fun
    print(f"HF_FIT_REFRESH:EXACT_BELOW={r}: {dt:.2f}s -> {2 * n / dt:.1f} fits/s; failed {(st != 0).sum().item()}; "
          f"iters {out['free'].niter.float().mean().item():.1f} (max {out['free'].niter.max().item()}); "
          f"max |2NLL - first setting| {(fun - ref).abs().max().item():.2e}; worse than first by > 1e-6: "
          f"{((fun - ref) > 1e-6).sum().item()}, better by > 1e-6: {((ref - fun) > 1e-6).sum().item()}", flush=True)
