#!/usr/bin/env python
"""Golden fixture of SURVEY.md §8 row f4 (many signal patches on one background workspace).

    python tests/golden/make_patchset_golden.py      (needs baseline/_ref, see oracle/make_ref.py)

Writes tests/golden/patchset.json with three cases:

* ``synthetic``: a 2-channel background workspace (histosys, normsys, lumi, staterror, shapesys on
  the backgrounds) and 6 signal patches (``add`` of a signal sample per channel: POI normfactor,
  a signal-only normsys shared by all patches, the background's JES histosys and lumi; one patch
  carries a normsys of its own, one adds no control-region sample);
* ``synthetic_single``: two patches on the same workspace that must run one at a time (signal
  with a staterror; patch that also edits a background systematic);
* ``reference_example``: the reference's own tests/test_patchset/example_bkgonly.json +
  example_patchset.json (the patch *replaces* values inside the sample that carries the POI).

For every patch the UNMODIFIED reference is run the way workspace.py:392-442 does -- patch the
model spec, build ``pyhf.Model``, data = observations in model channel order + auxdata -- and
``pyhf.infer.hypotest(mu, data, model, test_stat="qtilde", return_expected_set=True)`` with the
numpy backend and scipy SLSQP at tolerance 1e-12 (and at the default tolerance).  The reference
delegates patch application to the third-party ``jsonpatch`` (absent here); the patches are
applied by pyhf_b200.patchset.apply_patch, and the fixture stores the patched sample lists'
digest so that the test can pin apply_patch itself.
"""
import hashlib
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "baseline", "_ref"))

import pyhf  # noqa: E402  (the reference)

from pyhf_b200.patchset import apply_patch  # noqa: E402


def synthetic_case():
    rng = np.random.default_rng(20261020)
    nb = {"SR": 4, "CR": 3}

    def r(a):
        return [round(float(v), 3) for v in np.atleast_1d(a)]

    channels = []
    obs = []
    for c in ("SR", "CR"):
        n = nb[c]
        tt = rng.uniform(40, 120, n)
        wj = rng.uniform(20, 60, n)
        samples = [
            {"name": "ttbar", "data": r(tt), "modifiers": [
                {"name": "xsec_tt", "type": "normsys", "data": {"hi": 1.08, "lo": 0.93}},
                {"name": "jes", "type": "histosys",
                 "data": {"hi_data": r(tt * (1 + rng.uniform(0.01, 0.06, n))),
                          "lo_data": r(tt * (1 - rng.uniform(0.01, 0.06, n)))}},
                {"name": f"staterror_{c}", "type": "staterror", "data": r(0.3 * np.sqrt(tt))},
                {"name": "lumi", "type": "lumi", "data": None}]},
            {"name": "wjets", "data": r(wj), "modifiers": [
                {"name": "jes", "type": "histosys",
                 "data": {"hi_data": r(wj * (1 + rng.uniform(0.01, 0.05, n))),
                          "lo_data": r(wj * (1 - rng.uniform(0.01, 0.05, n)))}},
                {"name": f"staterror_{c}", "type": "staterror", "data": r(0.4 * np.sqrt(wj))},
                {"name": f"wj_shape_{c}", "type": "shapesys", "data": r(0.12 * wj)},
                {"name": "lumi", "type": "lumi", "data": None}]},
        ]
        channels.append({"name": c, "samples": samples})
        obs.append({"name": c, "data": [float(v) for v in rng.poisson(tt + wj)]})
    ws = {
        "channels": channels, "observations": obs, "version": "1.0.0",
        "measurements": [{"name": "meas", "config": {"poi": "mu", "parameters": [
            {"name": "lumi", "auxdata": [1.0], "sigmas": [0.02], "bounds": [[0.5, 1.5]], "inits": [1.0]},
            {"name": "mu", "bounds": [[0.0, 10.0]], "inits": [1.0]}]}}],
    }
    patches = []
    for k in range(6):
        ops = []
        for ci, c in enumerate(("SR", "CR")):
            if c == "CR" and k == 3:
                continue  # this signal point does not reach the control region
            n = nb[c]
            sig = rng.uniform(2, 9, n) * (1.0 if c == "SR" else 0.15) * (1 + 0.4 * k)
            mods = [
                {"name": "mu", "type": "normfactor", "data": None},
                {"name": "sig_xsec", "type": "normsys", "data": {"hi": 1.0 + 0.02 * (k + 1), "lo": 1.0 - 0.015 * (k + 1)}},
                {"name": "jes", "type": "histosys",
                 "data": {"hi_data": r(sig * (1 + rng.uniform(0.01, 0.05, n))),
                          "lo_data": r(sig * (1 - rng.uniform(0.01, 0.05, n)))}},
                {"name": "lumi", "type": "lumi", "data": None}]
            if k == 5:
                mods.append({"name": "sig_isr_p5", "type": "normsys", "data": {"hi": 1.1, "lo": 0.9}})
            ops.append({"op": "add", "path": f"/channels/{ci}/samples/" + ("0" if k % 2 else "-"),
                        "value": {"name": "signal", "data": r(sig), "modifiers": mods}})
        patches.append({"metadata": {"name": f"sig_{k}", "values": [100 * k]}, "patch": ops})
    return ws, patches, [0.5, 1.0, 2.0]


def single_case():
    """two patches on the synthetic workspace that cannot join a group: the signal carries the
    channel's staterror (its constraint width then depends on the patch); the patch also edits a
    background systematic"""
    ws, patches, _ = synthetic_case()
    a = json.loads(json.dumps(patches[1]))
    a["metadata"] = {"name": "sig_staterror", "values": [1000]}
    a["patch"][0]["value"]["modifiers"].append(
        {"name": "staterror_SR", "type": "staterror", "data": [0.9, 1.1, 0.8, 1.0]})
    b = json.loads(json.dumps(patches[2]))
    b["metadata"] = {"name": "sig_edits_bkg", "values": [1001]}
    # patches[2] appends the signal ("-"), so the backgrounds keep their positions
    b["patch"].append({"op": "replace", "path": "/channels/0/samples/0/modifiers/0/data/hi", "value": 1.12})
    return ws, [a, b], [1.0, 2.0]


def reference_example():
    d = "/root/reference/tests/test_patchset"
    ws = json.load(open(os.path.join(d, "example_bkgonly.json")))
    ps = json.load(open(os.path.join(d, "example_patchset.json")))
    return ws, ps["patches"], [1.0]


def run_reference(ws, patches, mus):
    out = {}
    for tag, kw in (("", {}), ("_tight", {"tolerance": 1e-12})):
        pyhf.set_backend("numpy", pyhf.optimize.scipy_optimizer(**kw))
        obs_all, exp_all = [], []
        for p in patches:
            w = apply_patch(ws, p)
            m = w["measurements"][0]
            spec = {"channels": w["channels"], "parameters": m["config"]["parameters"]}
            model = pyhf.Model(spec, poi_name=m["config"]["poi"])       # workspace.py:435-442
            observations = {o["name"]: o["data"] for o in w["observations"]}
            data = sum((list(observations[c]) for c in model.config.channels), []) + list(model.config.auxdata)
            o_row, e_row = [], []
            for mu in mus:
                o, e = pyhf.infer.hypotest(mu, data, model, test_stat="qtilde", return_expected_set=True)
                o_row.append(float(o))
                e_row.append([float(v) for v in e])
            obs_all.append(o_row)
            exp_all.append(e_row)
        out["cls_obs" + tag] = obs_all
        out["cls_exp" + tag] = exp_all
    return out


def digest(ws, patches):
    h = hashlib.sha256()
    for p in patches:
        h.update(json.dumps(apply_patch(ws, p)["channels"], sort_keys=True).encode())
    return h.hexdigest()


def main():
    cases = {}
    for name, fn in (("synthetic", synthetic_case), ("synthetic_single", single_case),
                     ("reference_example", reference_example)):
        ws, patches, mus = fn()
        ref = run_reference(ws, patches, mus)
        cases[name] = dict(workspace=ws, patches=patches, mus=mus, patched_channels_sha256=digest(ws, patches), **ref)
        print(name, "CLs_obs (tight):", np.round(np.array(ref["cls_obs_tight"]), 6).tolist())
        print(name, "max |default - tight|:", float(np.abs(np.array(ref["cls_obs"]) - np.array(ref["cls_obs_tight"])).max()))
    json.dump(cases, open(os.path.join(HERE, "patchset.json"), "w"), indent=1)


if __name__ == "__main__":
    main()
