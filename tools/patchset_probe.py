"""Timing of SURVEY.md §8 row f4: K signal points on one background workspace, grouped
(pyhf_b200.patchset.patchset_hypotest: two lock-step batches per group over a mask table) against
one patch at a time (batched.asymptotic_scan per patched model -- the same CUDA kernels, five
fit kinds per patch in two launches), and the agreement of the CLs values between the two.

    python tools/patchset_probe.py [--model c3|c4] [--k 64] [--group 32] [--mus 1.0]
"""
import argparse
import json
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pyhf_b200 import batched, synth  # noqa: E402
from pyhf_b200.batched import DevicePlan  # noqa: E402
from pyhf_b200.compile import compile_spec  # noqa: E402
from pyhf_b200.patchset import _model_spec, _observed, apply_patch, patchset_hypotest  # noqa: E402


def workspace(model, K, seed=7):
    spec = synth.spec_c3() if model == "c3" else synth.spec_c4()
    rng = np.random.default_rng(seed)
    obs, patches = [], [[] for _ in range(K)]
    for ci, ch in enumerate(spec["channels"]):
        ch["samples"] = [s for s in ch["samples"] if s["name"] != "signal"]
        nb = len(ch["samples"][0]["data"])
        tot = np.sum([s["data"] for s in ch["samples"]], axis=0)
        obs.append({"name": ch["name"], "data": [float(v) for v in rng.poisson(tot)]})
        x = np.arange(nb)
        for k in range(K):
            centre, width = (k + 0.5) * nb / K, 2.0 + 0.05 * nb
            sig = (30.0 + 2.0 * k) * np.exp(-0.5 * ((x - centre) / width) ** 2) + 0.5
            patches[k].append({"op": "add", "path": f"/channels/{ci}/samples/0", "value": {
                "name": "signal", "data": [round(float(v), 4) for v in sig], "modifiers": [
                    {"name": "mu", "type": "normfactor", "data": None},
                    {"name": "sig_theory", "type": "normsys", "data": {"hi": 1.06, "lo": 0.95}}]}})
    ws = {"channels": spec["channels"], "observations": obs, "version": "1.0.0",
          "measurements": [{"name": "m", "config": {"poi": "mu", "parameters": []}}]}
    return ws, patches


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--model", default="c3")
    ap.add_argument("--k", type=int, default=64)
    ap.add_argument("--group", type=int, default=32)
    ap.add_argument("--mus", type=float, nargs="+", default=[1.0])
    ap.add_argument("--reps", type=int, default=3)
    ap.add_argument("--no-single", action="store_true")
    a = ap.parse_args()
    dev = torch.device("cuda", 0)
    ws, patches = workspace(a.model, a.k)
    out = {"model": a.model, "K": a.k, "group_size": a.group, "mus": a.mus}

    def timed(fn):
        best, res = None, None
        for _ in range(a.reps):
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            res = fn()
            torch.cuda.synchronize()
            dt = time.perf_counter() - t0
            best = dt if best is None else min(best, dt)
        return best, res

    tg, g = timed(lambda: patchset_hypotest(ws, patches, a.mus, group_size=a.group, device=dev))
    out.update(group_s=tg, group_modes=sorted(set(g["mode"])), group_fits=g["n_fits"],
               group_failed=g["n_failed"], group_ms_per_signal_point=1e3 * tg / a.k)
    # host-side share: patch application + group model + plan compilation, no fits
    t0 = time.perf_counter()
    patched = [apply_patch(ws, p) for p in patches]
    out["apply_patches_s"] = time.perf_counter() - t0
    if not a.no_single:
        def singles():
            cls = []
            for w in patched:
                spec, poi = _model_spec(w, 0)
                host = compile_spec(spec, poi_name=poi)
                plan = DevicePlan(host, dev)
                data = np.concatenate([_observed(w, spec), host.auxdata])
                r = batched.asymptotic_scan(plan, a.mus, data)
                cls.append(r["CLs_obs"])
            return torch.stack(cls)
        ts, s = timed(singles)
        d = (s - g["CLs_obs"]).abs().max().item()
        out.update(single_s=ts, single_fits=a.k * (2 * len(a.mus) + 3), single_ms_per_signal_point=1e3 * ts / a.k,
                   max_abs_dCLs_group_vs_single=d, speedup=ts / tg)
    out["CLs_obs_first"] = [float(v) for v in g["CLs_obs"][:4, 0]]
    print(json.dumps(out))


if __name__ == "__main__":
    main()
