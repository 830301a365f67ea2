#!/usr/bin/env python
"""Reference fits of C4 toys (BASELINE.json configs 4/5) by the UNMODIFIED reference:
``pyhf.infer.mle.fit`` / ``fixed_poi_fit`` on numpy + scipy SLSQP (optimize/opt_scipy.py:96-105
under infer/mle.py:69-205), default tolerance and ``tolerance=1e-12``.

One C4 fit costs ~(P+1)=301 forward-difference evaluations of 25 ms per SLSQP iteration, i.e.
tens of CPU-minutes; run one toy per process in the background:

    for i in 0 1 2 3; do OMP_NUM_THREADS=1 nice python tests/golden/make_c4_fit_golden.py toy $i & done
    python tests/golden/make_c4_fit_golden.py merge      # -> tests/golden/c4_fits.npz

Toys: ``model.make_pdf(suggested_init).sample((4,))`` under ``np.random.seed(0)``; toy 0 is the
observed data of configs 4/5.  Parts are written after every fit, so a
partial run still merges (missing entries are NaN).
"""
import json
import os
import sys
import time

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "baseline", "_ref"))
PARTS = os.path.join(HERE, "c4_fit_parts")
NTOYS = 4
MU_TEST = 1.0


def c4_model_and_toys():
    import pyhf

    from pyhf_b200 import synth

    pyhf.set_backend("numpy", "scipy")
    model = pyhf.Model(synth.spec_c4(), poi_name="mu", validate=False)
    init = model.config.suggested_init()
    np.random.seed(0)
    toys = np.asarray(model.make_pdf(pyhf.tensorlib.astensor(init)).sample((NTOYS,)))
    return model, toys


def run_toy(i):
    import pyhf

    os.makedirs(PARTS, exist_ok=True)
    path = os.path.join(PARTS, f"toy{i}.npz")
    model, toys = c4_model_and_toys()
    t = toys[i]
    init = model.config.suggested_init()
    bounds = model.config.suggested_bounds()
    fixed = model.config.suggested_fixed()
    out = {"toy": t}
    if os.path.exists(path):
        out.update(dict(np.load(path)))
    for tag, kw in (("default", {}), ("tight", {"tolerance": 1e-12})):
        for which in ("fixed", "free"):
            if f"{which}_x_{tag}" in out:
                continue
            opt = pyhf.optimize.scipy_optimizer(**kw)
            pyhf.set_backend("numpy", opt)
            t0 = time.time()
            if which == "fixed":
                x, f, res = pyhf.infer.mle.fixed_poi_fit(MU_TEST, t, model, init, bounds, fixed,
                                                         return_fitted_val=True, return_result_obj=True)
            else:
                x, f, res = pyhf.infer.mle.fit(t, model, init, bounds, fixed,
                                               return_fitted_val=True, return_result_obj=True)
            dt = time.time() - t0
            out[f"{which}_x_{tag}"] = np.asarray(x)
            out[f"{which}_fun_{tag}"] = np.float64(f)
            out[f"{which}_nit_{tag}"] = np.int64(res.nit)
            out[f"{which}_nfev_{tag}"] = np.int64(res.nfev)
            out[f"{which}_seconds_{tag}"] = np.float64(dt)
            print(f"toy {i} {which} {tag}: 2nll={float(f):.9f} nit={res.nit} nfev={res.nfev} {dt:.0f}s",
                  flush=True)
            np.savez(path + ".tmp.npz", **out)
            os.replace(path + ".tmp.npz", path)


def merge():
    keys = [f"{w}_{q}_{t}" for w in ("fixed", "free") for q in ("x", "fun", "nit", "nfev", "seconds")
            for t in ("default", "tight")]
    parts = []
    for i in range(NTOYS):
        p = os.path.join(PARTS, f"toy{i}.npz")
        parts.append(dict(np.load(p)) if os.path.exists(p) else {})
    P = next(len(v["toy"]) for v in parts if v)  # noqa: F841  (D)
    npar = next(len(v[k]) for v in parts for k in v if k.endswith("_x_default") or k.endswith("_x_tight"))
    d = {}
    toys = np.full((NTOYS, P), np.nan)
    for i, v in enumerate(parts):
        if v:
            toys[i] = v["toy"]
    d["toys"] = toys
    for k in keys:
        if "_x_" in k:
            a = np.full((NTOYS, npar), np.nan)
        else:
            a = np.full((NTOYS,), np.nan)
        for i, v in enumerate(parts):
            if k in v:
                a[i] = v[k]
        d[f"toy_{k}"] = a
    import scipy

    d["meta_json"] = np.array(json.dumps(dict(
        name="c4_fits", generator="spec_c4", poi_name="mu", mu_test=MU_TEST, numpy=np.__version__,
        scipy=scipy.__version__,
        note="reference SLSQP fits of 4 C4 toys; toy 0 = observed data of configs 4/5")))
    path = os.path.join(HERE, "c4_fits.npz")
    np.savez_compressed(path, **d)
    for t in ("default", "tight"):
        print(t, "fixed", d[f"toy_fixed_fun_{t}"], "free", d[f"toy_free_fun_{t}"])
    print("wrote", path, os.path.getsize(path))


if __name__ == "__main__":
    if sys.argv[1] == "toy":
        run_toy(int(sys.argv[2]))
    else:
        merge()
