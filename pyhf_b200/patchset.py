"""Many signal hypotheses on one background model.

Reference workflow: a background-only workspace plus a patchset of JSON patches, one per signal
point; for every point the reference builds a fresh ``Model`` (``workspace.py:392-442`` applies
``jsonpatch.JsonPatch(patch).apply(modelspec)``, ``patchset.py:300`` applies it to the workspace)
and runs one asymptotic ``hypotest`` = five sequential fits (``infer/calculators.py:345-456``).
``jsonpatch`` (``pyproject.toml``: ``jsonpatch>=1.15``) is a third-party dependency that is not in
this image; :func:`apply_patch` restates RFC 6902 (add / remove / replace / move / copy / test over
RFC 6901 pointers), which is all the reference uses of it.

B200 form: the K signal points of a group become ONE model.  The samples that carry the POI
normfactor are the *variants* of a patch; everything else must be the same background in every
patch.  The group model holds the background once and the variants of every patch side by side,
patch k's POI renamed ``<poi>__p<k>``.  A fit for patch k fixes every other POI at zero, which
switches the other variants off exactly (their rate is multiplied by 0), so its likelihood is the
one of the reference's per-patch model up to a constant that cancels in every test statistic.
All fits of all patches of a group -- 2 K (M + 1) + 1 instead of 5 K M for M POI values -- run as
two lock-step batches over a table of fixed masks (``hf_fit_batch_masks``): the tables of the
background (the expensive part: histosys / normsys) live once in HBM, the background-only fit and
the Asimov data set are shared by all patches.

Patches that cannot be grouped (a variant carries a staterror / shapesys / shapefactor, whose
constraint width or parameter layout would depend on the patch; a patch edits the background or
the measurement) run one patch at a time through the same CUDA path
(:func:`batched.asymptotic_scan` on the patched model) -- there is no CPU route.
"""
from __future__ import annotations

import copy

import numpy as np
import torch

from . import batched
from .batched import DevicePlan
from .compile import compile_spec
from .plan import PlanError

__all__ = ["apply_patch", "NotBatchable", "SignalGroup", "split_signal", "build_signal_group",
           "group_masks", "patchset_hypotest", "patchset_upper_limits"]


class PatchError(ValueError):
    """a JSON patch that does not apply (jsonpatch.JsonPatchException / JsonPointerException)"""


# ---------------------------------------------------------------------------------- RFC 6902
def _tokens(pointer):
    if pointer == "":
        return []
    if not pointer.startswith("/"):
        raise PatchError(f"JSON pointer must start with '/': {pointer!r}")
    return [t.replace("~1", "/").replace("~0", "~") for t in pointer[1:].split("/")]


def _index(tok, n, allow_end):
    if tok == "-" and allow_end:
        return n
    if not (tok.isdigit() and (tok == "0" or tok[0] != "0")):
        raise PatchError(f"bad array index {tok!r}")
    i = int(tok)
    if i > n or (i == n and not allow_end):
        raise PatchError(f"array index {i} out of range")
    return i


def _walk(doc, toks):
    for t in toks:
        if isinstance(doc, list):
            doc = doc[_index(t, len(doc), False)]
        elif isinstance(doc, dict):
            if t not in doc:
                raise PatchError(f"member {t!r} not found")
            doc = doc[t]
        else:
            raise PatchError(f"cannot descend into a scalar at {t!r}")
    return doc


def _shallow(x):
    return dict(x) if isinstance(x, dict) else list(x) if isinstance(x, list) else x


def _own_parent(root, toks):
    """Parent container of ``toks`` inside ``root`` (itself already a fresh shallow copy), with
    every container on the way replaced by a shallow copy of its own: an operation then edits
    only copies, and everything off its path stays shared with the input document."""
    cur = root
    for t in toks[:-1]:
        if isinstance(cur, list):
            i = _index(t, len(cur), False)
        elif isinstance(cur, dict):
            if t not in cur:
                raise PatchError(f"member {t!r} not found")
            i = t
        else:
            raise PatchError(f"cannot descend into a scalar at {t!r}")
        cur[i] = _shallow(cur[i])
        cur = cur[i]
    return cur


def _add(root, toks, value):
    if not toks:
        return value
    parent = _own_parent(root, toks)
    if isinstance(parent, list):
        parent.insert(_index(toks[-1], len(parent), True), value)
    elif isinstance(parent, dict):
        parent[toks[-1]] = value
    else:
        raise PatchError("add into a scalar")
    return root


def _remove(root, toks):
    if not toks:
        raise PatchError("cannot remove the document root")
    parent = _own_parent(root, toks)
    if isinstance(parent, list):
        return parent.pop(_index(toks[-1], len(parent), False))
    if isinstance(parent, dict):
        if toks[-1] not in parent:
            raise PatchError(f"member {toks[-1]!r} not found")
        return parent.pop(toks[-1])
    raise PatchError("remove from a scalar")


def apply_patch(doc, patch):
    """Apply an RFC 6902 patch (list of operations) and return the patched document -- what
    ``jsonpatch.JsonPatch(patch).apply(doc)`` does at workspace.py:439-440.  ``doc`` is never
    modified; the result SHARES every subtree the patch does not touch with ``doc`` (copy on
    write along the operation paths: a signal patch on a large background workspace costs the
    size of the patch, not of the workspace), so treat both as read-only afterwards."""
    if isinstance(patch, dict) and "patch" in patch:  # a patchset entry {"metadata":…, "patch":[…]}
        patch = patch["patch"]
    doc = _shallow(doc)
    for op in patch:
        kind = op.get("op")
        toks = _tokens(op["path"])
        if kind == "add":
            doc = _add(doc, toks, copy.deepcopy(op["value"]))
        elif kind == "remove":
            _remove(doc, toks)
        elif kind == "replace":
            if not toks:
                doc = copy.deepcopy(op["value"])
            else:
                _walk(doc, toks)  # the target must exist
                _remove(doc, toks)
                doc = _add(doc, toks, copy.deepcopy(op["value"]))
        elif kind == "move":
            src = _tokens(op["from"])
            if src == toks[:len(src)] and len(src) < len(toks):
                raise PatchError("cannot move a value into one of its children")
            doc = _add(doc, toks, _remove(doc, src))
        elif kind == "copy":
            doc = _add(doc, toks, copy.deepcopy(_walk(doc, _tokens(op["from"]))))
        elif kind == "test":
            if _walk(doc, toks) != op["value"]:
                raise PatchError(f"test failed at {op['path']}")
        else:
            raise PatchError(f"unknown operation {kind!r}")
    return doc


# ---------------------------------------------------------------------------------- grouping
class NotBatchable(Exception):
    """the patches cannot share one group model; the reason is the message"""


def _measurement(ws, measurement_index):
    ms = ws.get("measurements") or []
    if not ms:
        raise ValueError("No measurements have been defined.")  # workspace.py:386-387
    return ms[measurement_index]


def _model_spec(ws, measurement_index):
    m = _measurement(ws, measurement_index)
    return {"channels": ws["channels"], "parameters": m["config"].get("parameters", [])}, m["config"]["poi"]


def _observed(ws, spec):
    """main data in the model's channel order (workspace.py:460-463: sorted channel names)"""
    obs = {o["name"]: list(o["data"]) for o in ws["observations"]}
    out = []
    for c in sorted(ch["name"] for ch in spec["channels"]):
        if c not in obs:
            raise KeyError(f"Invalid channel: the workspace has no observation for {c}")
        out += obs[c]
    return np.asarray(out, float)


def _has_poi(sample, poi):
    return any(m["type"] == "normfactor" and m["name"] == poi for m in sample["modifiers"])


class SignalGroup:
    """group model of K patches: ``spec`` (the HistFactory spec of the combined model),
    ``poi_names`` [K], ``users`` (parameter-set name -> "bkg" or the set of patch indices whose
    variants carry it)."""

    def __init__(self, spec, poi_names, users):
        self.spec, self.poi_names, self.users = spec, poi_names, users


_SIG_CACHE = {}


def split_signal(spec, poi, k=0):
    """One patched model spec -> (background signature, variants): the samples without the POI
    normfactor, as a canonical JSON string together with the channel list and the parameter
    configuration (patches with equal signatures can share a group), and per channel the samples
    that carry the POI, renamed for slot ``k`` of a group.  Raises :class:`NotBatchable`."""
    import json

    bkg, variants, names = [], {}, {}
    for ch in spec["channels"]:
        variants[ch["name"]] = []
        for s in ch["samples"]:
            if not _has_poi(s, poi):
                bkg.append((ch["name"], s["name"], s))
                continue
            v = copy.deepcopy(s)
            v["name"] = f"{s['name']}__p{k}"
            for m in v["modifiers"]:
                if m["type"] in ("staterror", "shapesys", "shapefactor"):
                    raise NotBatchable(f"signal sample {s['name']!r} carries a {m['type']} "
                                       "(per-patch constraint widths / parameter layout)")
                if m["type"] == "normfactor" and m["name"] == poi:
                    m["name"] = f"{poi}__p{k}"
                else:
                    names[m["name"]] = True
            variants[ch["name"]].append(v)
    if not any(variants.values()):
        raise NotBatchable(f"no sample carries the POI normfactor {poi!r}")
    bkg.sort(key=lambda t: (t[0], t[1]))
    # canonical JSON of the background, cached by object identity: apply_patch shares the untouched
    # samples with the background workspace, so all patches of a patchset hit one cache entry
    pars = spec.get("parameters", [])
    key = (tuple(c["name"] for c in spec["channels"]), tuple((c, n, id(smp)) for c, n, smp in bkg), id(pars))
    hit = _SIG_CACHE.get(key)
    if hit is None:
        if len(_SIG_CACHE) > 64:
            _SIG_CACHE.clear()
        sig = json.dumps([list(key[0]), bkg, pars], sort_keys=True)
        _SIG_CACHE[key] = hit = (sig, [smp for _, _, smp in bkg], pars)  # keeps the ids alive
    return hit[0], variants, list(names)


def build_signal_group(specs, poi):
    """Combine the patched model specs ``specs`` (``{"channels", "parameters"}`` each, all built
    from one background workspace) into one group model, or raise :class:`NotBatchable`."""
    K = len(specs)
    if K == 0:
        raise ValueError("no patches")
    users, sig0 = {}, None
    group_channels = {c["name"]: [] for c in specs[0]["channels"]}
    for k, sp in enumerate(specs):
        try:
            sig, variants, names = split_signal(sp, poi, k)
        except NotBatchable as exc:
            raise NotBatchable(f"patch {k}: {exc}") from None
        if sig0 is None:
            sig0 = sig
        elif sig != sig0:
            raise NotBatchable(f"patch {k} edits the background samples, the channels or the "
                               "measurement's parameter configuration")
        for c, lst in variants.items():
            group_channels[c] += lst
        for n in names:
            users.setdefault(n, set()).add(k)
    channels = []
    for ch in specs[0]["channels"]:
        out = {key: val for key, val in ch.items() if key != "samples"}
        keep = [s for s in ch["samples"] if not _has_poi(s, poi)]
        for s in keep:
            for m in s["modifiers"]:
                users[m["name"]] = "bkg"
        out["samples"] = keep + group_channels[ch["name"]]
        channels.append(out)
    ref_pars = specs[0].get("parameters", [])
    poi_names = [f"{poi}__p{k}" for k in range(K)]
    pars = [p for p in ref_pars if p["name"] != poi]
    for k, n in enumerate(poi_names):
        for p in ref_pars:
            if p["name"] == poi:
                q = copy.deepcopy(p)
                q["name"] = n
                pars.append(q)
        users[n] = {k}
    return SignalGroup({"channels": channels, "parameters": pars}, poi_names, users)


def group_masks(group: SignalGroup, host):
    """mask table of a group: row 0 = background-only fit (every signal-only parameter fixed),
    row 1 + 2k = POI-fixed fit of patch k, row 2 + 2k = free fit of patch k; plus the POI index of
    every patch and the start point (all POIs at zero)."""
    K = len(group.poi_names)
    P = host.n_pars
    base = host.fixed.astype(np.uint8)
    masks = np.tile(base, (2 * K + 1, 1))
    poi_idx = np.array([host.par_slices[n][0] for n in group.poi_names], np.int64)
    for name, (start, n) in host.par_slices.items():
        u = group.users.get(name, "bkg")
        if u == "bkg":
            continue
        sl = slice(start, start + n)
        masks[0, sl] = 1
        for k in range(K):
            if k not in u:
                masks[1 + 2 * k, sl] = masks[2 + 2 * k, sl] = 1
    for k in range(K):
        masks[1 + 2 * k, poi_idx[k]] = 1
    init0 = host.init.copy()
    init0[poi_idx] = 0.0
    assert masks.shape == (2 * K + 1, P)
    return masks, poi_idx, init0


# ---------------------------------------------------------------------------------- inference
def _group_hypotest(group, data_main, mus, test_stat, calc_base_dist, device, modifier_settings, fit_kw):
    host = compile_spec(group.spec, poi_name=None, modifier_settings=modifier_settings)
    plan = DevicePlan(host, device)
    dev = plan.device
    K, M = len(group.poi_names), mus.numel()
    masks, poi_idx, init0 = group_masks(group, host)
    poi_t = torch.as_tensor(poi_idx, device=dev)
    poi_init = torch.as_tensor(host.init[poi_idx], device=dev)
    data = torch.as_tensor(np.concatenate([data_main, host.auxdata]), device=dev).reshape(1, -1)
    x0 = torch.as_tensor(init0, device=dev)
    kk = torch.arange(K, device=dev)

    def rows(with_bkg_fit):
        """[bkg-only]? | free fit of every patch | POI-fixed fit of every (patch, mu)"""
        free = x0.expand(K, -1).clone()
        free[kk, poi_t] = poi_init
        fx = x0.expand(K * M, -1).clone()
        fx[torch.arange(K * M, device=dev), poi_t.repeat_interleave(M)] = mus.repeat(K)
        init = torch.cat(([x0.reshape(1, -1)] if with_bkg_fit else []) + [free, fx])
        ids = torch.cat(([torch.zeros(1, dtype=torch.int64, device=dev)] if with_bkg_fit else [])
                        + [2 + 2 * kk, (1 + 2 * kk).repeat_interleave(M)]).to(torch.uint8)
        return init, ids

    def run(d, with_bkg_fit):
        init, ids = rows(with_bkg_fit)
        return plan.fit(d, init, masks=masks, mask_id=ids,
                        data_map=torch.zeros(ids.numel(), dtype=torch.int32, device=dev), **fit_kw)

    r1 = run(data, True)
    asimov = plan.expected_data(r1.x[0:1])            # calculators.py:45-97, asimov_mu = 0
    r2 = run(asimov.reshape(1, -1), False)
    mus2 = mus.reshape(1, M)

    def q_of(free_fun, free_mu, fixed_fun):
        q = torch.clamp(fixed_fun.reshape(K, M) - free_fun.reshape(K, 1), min=0.0)  # test_statistics.py:56
        return torch.where(free_mu.reshape(K, 1) > mus2, torch.zeros_like(q), q)      # :57-60

    muhat = r1.x[1:1 + K][kk, poi_t]
    muhat_a = r2.x[:K][kk, poi_t]
    qmu = q_of(r1.fun[1:1 + K], muhat, r1.fun[1 + K:])
    qmuA = q_of(r2.fun[:K], muhat_a, r2.fun[K:])
    qmuA = torch.where(mus2 == 0.0, torch.zeros_like(qmuA), qmuA)  # see batched.asymptotic_scan
    out = batched.asymptotic_pvalues(qmu, qmuA, test_stat, calc_base_dist)
    status = torch.cat([r1.status, r2.status])
    out.update(muhat=muhat, qmu=qmu, qmuA=qmuA, n_fits=int(status.numel()),
               n_failed=int((status != 0).sum()), launches=plan.launch_count)
    return out


def patchset_hypotest(workspace, patches, mus=1.0, test_stat="qtilde", calc_base_dist="normal",
                      measurement_index=0, group_size=32, device=None, modifier_settings=None,
                      **fit_kw):
    """Asymptotic CLs of every signal patch of a patchset on one background workspace.

    ``workspace``: the background-only workspace (dict: channels / observations / measurements);
    ``patches``: a patchset dict (``{"patches": [...]}``), or a list whose items are RFC 6902
    operation lists or patchset entries (``{"metadata":…, "patch": […]}``); ``mus``: POI value(s)
    tested for every patch.  Returns a dict of tensors: ``CLs_obs`` [K, M], ``CLs_exp`` [K, M, 5]
    (n_sigma = 2, 1, 0, -1, -2 as in the reference), ``CLsb``, ``CLb``, ``muhat`` [K], ``qmu``,
    ``qmuA`` [K, M], and ``mode`` [K] ("group" or "single"), ``n_fits``, ``n_failed``.

    Equivalent reference code (5 K M fits, one model build per patch)::

        for patch in patchset:
            model = ws.model(patches=[patch]); data = ws.data(model)     # workspace.py:392-442
            for mu in mus: pyhf.infer.hypotest(mu, data, model, test_stat=…, return_expected_set=True)
    """
    if test_stat not in ("qtilde", "q"):
        raise ValueError("patchset_hypotest supports the exclusion statistics 'qtilde' and 'q'")
    if calc_base_dist not in ("normal", "clipped_normal"):
        raise ValueError(f"unknown base distribution for asymptotics {calc_base_dist}")
    if isinstance(patches, dict):
        patches = patches["patches"]
    if not 1 <= group_size <= 127:
        raise ValueError("group_size must be in 1..127 (mask table of 2 K + 1 rows)")
    device = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
    mus_t = torch.as_tensor(np.atleast_1d(np.asarray(mus, float)), device=device)
    patched = [apply_patch(workspace, p) for p in patches]
    specs, pois = zip(*[_model_spec(w, measurement_index) for w in patched]) if patched else ((), ())
    K = len(specs)
    if K == 0:
        raise ValueError("no patches")
    results, mode = [None] * K, ["single"] * K
    todo = list(range(K))
    # patches that can share a group: same background signature and the same observations
    buckets = {}
    for i in range(K):
        try:
            sig = split_signal(specs[i], pois[i])[0]
        except NotBatchable:
            continue
        main = _observed(patched[i], specs[i])
        buckets.setdefault((pois[i], sig, main.tobytes()), []).append(i)
    for (poi, _, _), members in buckets.items():
        for g0 in range(0, len(members), group_size):
            idx = members[g0:g0 + group_size]
            try:
                group = build_signal_group([specs[i] for i in idx], poi)
                res = _group_hypotest(group, _observed(patched[idx[0]], specs[idx[0]]), mus_t, test_stat,
                                      calc_base_dist, device, modifier_settings, fit_kw)
            except (NotBatchable, PlanError):
                continue
            for j, i in enumerate(idx):
                results[i] = {k: (v[j] if torch.is_tensor(v) else v) for k, v in res.items()}
                results[i]["n_fits"] = res["n_fits"] / len(idx)
                results[i]["n_failed"] = res["n_failed"] / len(idx)
                mode[i] = "group"
                todo.remove(i)
    for i in todo:  # one patch at a time, same CUDA path
        host = compile_spec(specs[i], poi_name=pois[i], modifier_settings=modifier_settings)
        plan = DevicePlan(host, device)
        data = np.concatenate([_observed(patched[i], specs[i]), host.auxdata])
        r = batched.asymptotic_scan(plan, mus_t, data, test_stat=test_stat,
                                    calc_base_dist=calc_base_dist, **fit_kw)
        results[i] = dict(CLs_obs=r["CLs_obs"], CLs_exp=r["CLs_exp"], CLsb=r["CLsb"], CLb=r["CLb"],
                          teststat=r["teststat"], muhat=r["muhat"], qmu=r["qmu"], qmuA=r["qmuA"],
                          n_fits=r["n_fits"], n_failed=r["n_failed"])
    out = {k: torch.stack([r[k] for r in results])
           for k in ("CLs_obs", "CLs_exp", "CLsb", "CLb", "teststat", "muhat", "qmu", "qmuA")}
    out.update(mu=mus_t, mode=mode, n_fits=int(round(sum(r["n_fits"] for r in results))),
               n_failed=int(round(sum(r["n_failed"] for r in results))))
    return out


def patchset_upper_limits(workspace, patches, scan, level=0.05, **kw):
    """Observed and expected upper limits on the POI of every signal patch: the asymptotic CLs scan
    of ``pyhf.infer.intervals.upper_limits.upper_limit(data, model, scan)`` (linear_grid_scan,
    upper_limits.py:148-213: one ``hypotest`` per scan point, then ``np.interp(level, cls[::-1],
    scan[::-1])``, :16-18) for all patches at once -- ``patchset_hypotest`` with ``mus = scan``.
    Returns its dict plus ``obs_limit`` [K] and ``exp_limits`` [K, 5] (n_sigma = 2, 1, 0, -1, -2
    ordering of ``CLs_exp``, as the reference returns them)."""
    scan = np.asarray(scan, float).reshape(-1)
    out = patchset_hypotest(workspace, patches, scan, **kw)
    obs = out["CLs_obs"].detach().cpu().numpy()
    exp = out["CLs_exp"].detach().cpu().numpy()
    K = obs.shape[0]
    out["obs_limit"] = torch.as_tensor([batched.upper_limit_from_scan(scan, obs[k], level) for k in range(K)],
                                       device=out["CLs_obs"].device)
    out["exp_limits"] = torch.as_tensor([batched.upper_limit_from_scan(scan, exp[k], level) for k in range(K)],
                                        device=out["CLs_obs"].device)
    return out
