"""CPU: host logic of SURVEY.md §8 row f4 (pyhf_b200/patchset.py).

* apply_patch against the known answers of RFC 6902 Appendix A (the reference delegates to the
  third-party jsonpatch, which restates the same RFC) and against the digest stored with the
  golden fixture;
* the signal-group construction: structure, the reasons for refusing a group, and -- through the
  numpy oracle -- that a fit row of patch k sees the per-patch likelihood up to a constant;
* the row / mask bookkeeping of the grouped hypotest, driven on the CPU by an oracle-backed
  stand-in for DevicePlan (test infrastructure only) against the reference's CLs
  (tests/golden/patchset.json, made by tests/golden/make_patchset_golden.py).
"""
import copy
import hashlib
import json
import os

import numpy as np
import pytest
import torch

from oracle import hf_oracle as O
from pyhf_b200 import patchset as PS
from pyhf_b200.compile import compile_spec

HERE = os.path.dirname(os.path.abspath(__file__))
CASES = json.load(open(os.path.join(HERE, "golden", "patchset.json")))


# ------------------------------------------------------------------------------ RFC 6902
RFC = [  # (document, patch, result) of RFC 6902 Appendix A.1 - A.8, A.10, A.11, A.14, A.16
    ({"foo": "bar"}, [{"op": "add", "path": "/baz", "value": "qux"}], {"baz": "qux", "foo": "bar"}),
    ({"foo": ["bar", "baz"]}, [{"op": "add", "path": "/foo/1", "value": "qux"}], {"foo": ["bar", "qux", "baz"]}),
    ({"baz": "qux", "foo": "bar"}, [{"op": "remove", "path": "/baz"}], {"foo": "bar"}),
    ({"foo": ["bar", "qux", "baz"]}, [{"op": "remove", "path": "/foo/1"}], {"foo": ["bar", "baz"]}),
    ({"baz": "qux", "foo": "bar"}, [{"op": "replace", "path": "/baz", "value": "boo"}], {"baz": "boo", "foo": "bar"}),
    ({"foo": {"bar": "baz", "waldo": "fred"}, "qux": {"corge": "grault"}},
     [{"op": "move", "from": "/foo/waldo", "path": "/qux/thud"}],
     {"foo": {"bar": "baz"}, "qux": {"corge": "grault", "thud": "fred"}}),
    ({"foo": ["all", "grass", "cows", "eat"]}, [{"op": "move", "from": "/foo/1", "path": "/foo/3"}],
     {"foo": ["all", "cows", "eat", "grass"]}),
    ({"baz": "qux", "foo": ["a", 2, "c"]},
     [{"op": "test", "path": "/baz", "value": "qux"}, {"op": "test", "path": "/foo/1", "value": 2}],
     {"baz": "qux", "foo": ["a", 2, "c"]}),
    ({"foo": "bar"}, [{"op": "add", "path": "/child", "value": {"grandchild": {}}}],
     {"foo": "bar", "child": {"grandchild": {}}}),
    ({"foo": "bar"}, [{"op": "add", "path": "/baz", "value": "qux", "xyz": 123}], {"foo": "bar", "baz": "qux"}),
    ({"/": 9, "~1": 10}, [{"op": "test", "path": "/~01", "value": 10}], {"/": 9, "~1": 10}),
    ({"foo": ["bar"]}, [{"op": "add", "path": "/foo/-", "value": ["abc", "def"]}], {"foo": ["bar", ["abc", "def"]]}),
]


@pytest.mark.parametrize("doc,patch,want", RFC)
def test_apply_patch_rfc6902_known_answers(doc, patch, want):
    before = copy.deepcopy(doc)
    assert PS.apply_patch(doc, patch) == want
    assert doc == before  # the input is never modified (jsonpatch apply(in_place=False))


@pytest.mark.parametrize("doc,patch", [
    ({"baz": "qux"}, [{"op": "test", "path": "/baz", "value": "bar"}]),        # A.9
    ({"foo": "bar"}, [{"op": "add", "path": "/baz/bat", "value": "qux"}]),     # A.12
    ({"/": 9, "~1": 10}, [{"op": "test", "path": "/~01", "value": "10"}]),     # A.15
    ({"foo": [1]}, [{"op": "remove", "path": "/foo/1"}]),
    ({"foo": [1]}, [{"op": "add", "path": "/foo/3", "value": 0}]),
    ({"foo": [1]}, [{"op": "replace", "path": "/bar", "value": 0}]),
    ({"foo": [1]}, [{"op": "frobnicate", "path": "/foo"}]),
    ({"foo": [1]}, [{"op": "add", "path": "/foo/01", "value": 0}]),
])
def test_apply_patch_errors(doc, patch):
    with pytest.raises(PS.PatchError):
        PS.apply_patch(doc, patch)


def test_apply_patch_copies_only_along_the_operation_paths():
    """the patched workspace shares every untouched subtree with the background workspace (a
    signal patch costs the size of the patch), and the background is never modified"""
    c = CASES["synthetic"]
    ws = c["workspace"]
    before = copy.deepcopy(ws)
    out = PS.apply_patch(ws, c["patches"][1])          # inserts the signal at samples/0
    assert ws == before
    assert out["observations"] is ws["observations"] and out["measurements"] is ws["measurements"]
    for ch_out, ch_in in zip(out["channels"], ws["channels"]):
        assert ch_out is not ch_in and ch_out["samples"] is not ch_in["samples"]
        assert [s["name"] for s in ch_out["samples"]] == ["signal", "ttbar", "wjets"]
        assert ch_out["samples"][1] is ch_in["samples"][0] and ch_out["samples"][2] is ch_in["samples"][1]
    # all patches of a patchset therefore share one background signature computation
    sigs = {PS.split_signal(PS._model_spec(PS.apply_patch(ws, p), 0)[0], "mu")[0] for p in c["patches"]}
    assert len(sigs) == 1
    # ... and an equal but separately built background lands in the same bucket
    other = PS.apply_patch(copy.deepcopy(ws), c["patches"][0])
    assert PS.split_signal(PS._model_spec(other, 0)[0], "mu")[0] in sigs


@pytest.mark.parametrize("case", sorted(CASES))
def test_patched_workspaces_match_the_golden_digest(case):
    c = CASES[case]
    h = hashlib.sha256()
    for p in c["patches"]:
        h.update(json.dumps(PS.apply_patch(c["workspace"], p)["channels"], sort_keys=True).encode())
    assert h.hexdigest() == c["patched_channels_sha256"]


# ------------------------------------------------------------------------------ grouping
def _specs(case):
    c = CASES[case]
    patched = [PS.apply_patch(c["workspace"], p) for p in c["patches"]]
    specs, pois = zip(*[PS._model_spec(w, 0) for w in patched])
    return c, patched, list(specs), pois[0]


def test_group_structure_and_masks():
    c, patched, specs, poi = _specs("synthetic")
    g = PS.build_signal_group(specs, poi)
    K = len(specs)
    assert g.poi_names == [f"mu__p{k}" for k in range(K)]
    names = {ch["name"]: [s["name"] for s in ch["samples"]] for ch in g.spec["channels"]}
    assert names["SR"] == ["ttbar", "wjets"] + [f"signal__p{k}" for k in range(K)]
    assert names["CR"] == ["ttbar", "wjets"] + [f"signal__p{k}" for k in range(K) if k != 3]
    assert g.users["jes"] == "bkg" and g.users["lumi"] == "bkg"
    assert g.users["sig_xsec"] == set(range(K)) and g.users["sig_isr_p5"] == {5}
    host = compile_spec(g.spec, poi_name=None)
    masks, poi_idx, init0 = PS.group_masks(g, host)
    assert masks.shape == (2 * K + 1, host.n_pars) and np.all(init0[poi_idx] == 0.0)
    sx, isr = host.par_slices["sig_xsec"][0], host.par_slices["sig_isr_p5"][0]
    assert np.all(masks[0, poi_idx] == 1) and masks[0, sx] == 1 and masks[0, isr] == 1
    for k in range(K):
        fx, fr = masks[1 + 2 * k], masks[2 + 2 * k]
        assert fx[poi_idx[k]] == 1 and fr[poi_idx[k]] == 0
        others = np.delete(poi_idx, k)
        assert np.all(fx[others] == 1) and np.all(fr[others] == 1)
        assert fx[sx] == 0 and fr[sx] == 0
        assert fr[isr] == (0 if k == 5 else 1)
        # nothing of the background is fixed that the model does not fix itself
        bkg = [s for n, (s, m) in host.par_slices.items() if g.users.get(n) == "bkg" for s in range(s, s + m)]
        assert np.array_equal(fr[bkg], host.fixed[bkg])
    # POI configuration of the measurement is copied to every slot
    assert np.all(host.lo[poi_idx] == 0.0) and np.all(host.hi[poi_idx] == 10.0)


def test_groups_are_refused_for_the_documented_reasons():
    c, patched, specs, poi = _specs("synthetic")
    bad = copy.deepcopy(specs[1])
    sig = next(s for s in bad["channels"][0]["samples"] if s["name"] == "signal")
    sig["modifiers"].append({"name": "staterror_SR", "type": "staterror", "data": [0.1] * 4})
    with pytest.raises(PS.NotBatchable, match="staterror"):
        PS.build_signal_group([specs[0], bad], poi)
    bad = copy.deepcopy(specs[1])
    bad["channels"][0]["samples"] = [s for s in bad["channels"][0]["samples"] if s["name"] != "wjets"]
    with pytest.raises(PS.NotBatchable, match="background"):
        PS.build_signal_group([specs[0], bad], poi)
    bad = copy.deepcopy(specs[1])
    bad["parameters"] = bad["parameters"] + [{"name": "xsec_tt", "fixed": True}]
    with pytest.raises(PS.NotBatchable, match="background|parameter"):
        PS.build_signal_group([specs[0], bad], poi)
    with pytest.raises(PS.NotBatchable, match="POI"):
        PS.build_signal_group([PS._model_spec(c["workspace"], 0)[0]], poi)


def test_ungroupable_golden_case_is_recognised():
    c, patched, specs, poi = _specs("synthetic_single")
    with pytest.raises(PS.NotBatchable, match="staterror"):
        PS.split_signal(specs[0], poi)
    sig_b = PS.split_signal(specs[1], poi)[0]
    assert sig_b != PS.split_signal(_specs("synthetic")[2][2], poi)[0]


def test_group_likelihood_is_the_patch_likelihood_up_to_a_constant():
    """2NLL of the group model with every other POI at zero minus 2NLL of the per-patch model is
    the same number at every parameter point (it cancels in every test statistic)."""
    c, patched, specs, poi = _specs("synthetic")
    g = PS.build_signal_group(specs, poi)
    gh = compile_spec(g.spec, poi_name=None)
    gp = O.Plan(gh.arrays())
    masks, poi_idx, init0 = PS.group_masks(g, gh)
    rng = np.random.default_rng(3)
    main = PS._observed(patched[0], specs[0])
    for k in (0, 3, 5):
        hk = compile_spec(specs[k], poi_name=poi)
        pk = O.Plan(hk.arrays())
        diffs = []
        for _ in range(5):
            x = hk.init + rng.normal(0, 0.15, hk.n_pars) * (1 - hk.fixed)
            x = np.clip(x, hk.lo + 1e-3, hk.hi)
            xg = init0.copy()
            for name, (s, n) in hk.par_slices.items():
                gs = gh.par_slices[g.poi_names[k] if name == poi else name][0]
                xg[gs:gs + n] = x[s:s + n]
            f_k = O.twice_nll(pk, x[None], np.concatenate([main, hk.auxdata])[None])[0]
            f_g = O.twice_nll(gp, xg[None], np.concatenate([main, gh.auxdata])[None])[0]
            diffs.append(f_g - f_k)
        assert np.ptp(diffs) < 1e-9 * max(1.0, abs(diffs[0])), diffs


# ------------------------------------------------------------------------------ row bookkeeping
class OraclePlan:
    """CPU stand-in for batched.DevicePlan in this test only: the same ``fit`` / ``expected_data``
    surface, every fit done by the oracle's tight scipy arbiter (oracle/hf_oracle.py::fit)."""

    launch_count = 0

    def __init__(self, host, device=None):
        self.host, self.device = host, torch.device("cpu")
        self.op = O.Plan(host.arrays())
        self.P = host.n_pars

    def expected_data(self, pars):
        return torch.as_tensor(O.expected_data(self.op, np.atleast_2d(pars.numpy())))

    def fit(self, data, init, masks=None, mask_id=None, data_map=None, **kw):
        from pyhf_b200.batched import FitResult

        data, init = np.atleast_2d(data.numpy()), init.numpy()
        masks, ids, rows = np.asarray(masks), mask_id.numpy(), data_map.numpy()
        xs, fs = [], []
        for i in range(len(ids)):
            x, f = O.fit(self.op, data[rows[i]], init=init[i], fixed=masks[ids[i]].astype(bool))
            xs.append(x)
            fs.append(f)
        n = len(ids)
        return FitResult(torch.as_tensor(np.array(xs)), torch.as_tensor(np.array(fs)),
                         torch.zeros(n, dtype=torch.int32), torch.zeros(n, dtype=torch.int32))


@pytest.mark.parametrize("case", ["reference_example", "synthetic"])
def test_grouped_hypotest_bookkeeping_against_reference_cls(case, monkeypatch):
    c, patched, specs, poi = _specs(case)
    monkeypatch.setattr(PS, "DevicePlan", OraclePlan)
    g = PS.build_signal_group(specs, poi)
    mus = torch.as_tensor(np.asarray(c["mus"], float))
    out = PS._group_hypotest(g, PS._observed(patched[0], specs[0]), mus, "qtilde", "normal", None, None, {})
    K, M = len(specs), len(c["mus"])
    assert out["n_fits"] == 2 * K * (M + 1) + 1 and out["n_failed"] == 0
    np.testing.assert_allclose(out["CLs_obs"].numpy(), np.array(c["cls_obs_tight"]), rtol=0, atol=1e-6)
    np.testing.assert_allclose(out["CLs_exp"].numpy(), np.array(c["cls_exp_tight"]), rtol=0, atol=1e-6)


# ------------------------------------------------------------------------------ properties
def _paths(doc, prefix=""):
    """every JSON pointer into doc (containers and leaves), RFC 6901 escaped"""
    out = [prefix]
    if isinstance(doc, dict):
        for k, v in doc.items():
            out += _paths(v, prefix + "/" + k.replace("~", "~0").replace("/", "~1"))
    elif isinstance(doc, list):
        for i, v in enumerate(doc):
            out += _paths(v, f"{prefix}/{i}")
    return out


def test_apply_patch_properties_on_random_documents():
    """size-independent properties: add then remove (and move there and back) restores the
    document, replace by the present value changes nothing, copy duplicates, the input is never
    modified -- on random nested documents with keys that need pointer escaping"""
    from hypothesis import given, settings
    from hypothesis import strategies as st

    keys = st.sampled_from(["a", "b", "c/d", "e~f", "", "0", "-"])
    leaves = st.one_of(st.integers(-5, 5), st.floats(allow_nan=False, allow_infinity=False), st.booleans(),
                       st.none(), st.text(max_size=3))
    docs = st.recursive(leaves, lambda ch: st.one_of(st.lists(ch, max_size=4), st.dictionaries(keys, ch, max_size=4)),
                        max_leaves=25)
    tops = st.one_of(st.lists(docs, max_size=4), st.dictionaries(keys, docs, min_size=1, max_size=4))

    @settings(max_examples=150, deadline=None)
    @given(tops, docs, st.data())
    def check(doc, value, data):
        before = copy.deepcopy(doc)
        paths = [p for p in _paths(doc) if p]
        if not paths:
            return
        path = data.draw(st.sampled_from(paths))
        present = PS._walk(doc, PS._tokens(path))
        # replace by the present value: nothing changes; the input is untouched
        assert PS.apply_patch(doc, [{"op": "replace", "path": path, "value": present}]) == before
        # test succeeds on the present value, fails on a different one
        PS.apply_patch(doc, [{"op": "test", "path": path, "value": present}])
        with pytest.raises(PS.PatchError):
            PS.apply_patch(doc, [{"op": "test", "path": path, "value": ["certainly", "different", 1]}])
        # remove then add back at the same pointer restores the document (list order included)
        assert PS.apply_patch(doc, [{"op": "remove", "path": path},
                                    {"op": "add", "path": path, "value": present}]) == before
        # add next to it (lists: insert before; dicts: a new member), then remove again
        parent_path, _, last = path.rpartition("/")
        parent = PS._walk(doc, PS._tokens(parent_path))
        new = parent_path + ("/" + last if isinstance(parent, list) else "/new~0member")
        if isinstance(parent, list) or "new~member" not in parent:
            out = PS.apply_patch(doc, [{"op": "add", "path": new, "value": value}])
            assert PS._walk(out, PS._tokens(new)) == value
            assert PS.apply_patch(out, [{"op": "remove", "path": new}]) == before
            # copy duplicates the value, move there and back is the identity
            dup = PS.apply_patch(doc, [{"op": "copy", "from": path, "path": new}])
            assert PS._walk(dup, PS._tokens(new)) == present
            if isinstance(parent, dict):
                there = PS.apply_patch(doc, [{"op": "move", "from": path, "path": new}])
                assert PS.apply_patch(there, [{"op": "move", "from": new, "path": path}]) == before
        assert doc == before

    check()


def test_upper_limits_per_patch_against_the_reference(monkeypatch):
    """patchset_upper_limits = the reference's upper_limit(data, model, scan) for every patch
    (tests/golden/add_patchset_limits.py), here through the oracle-backed stand-in for DevicePlan"""
    c = CASES["synthetic"]
    lim = c["limits"]
    monkeypatch.setattr(PS, "DevicePlan", OraclePlan)
    patches = [c["patches"][k] for k in lim["patches"]]
    out = PS.patchset_upper_limits(c["workspace"], patches, lim["scan"], device="cpu")
    assert out["mode"] == ["group", "group"] and out["n_failed"] == 0
    np.testing.assert_allclose(out["obs_limit"].numpy(), lim["obs_limit_tight"], rtol=0, atol=2e-6)
    np.testing.assert_allclose(out["exp_limits"].numpy(), lim["exp_limits_tight"], rtol=0, atol=2e-6)
    assert out["CLs_obs"][:, 0].tolist() == [1.0, 1.0]   # mu = 0: the degenerate point, no NaN
