"""Plan compiler: HistFactory JSON spec -> :class:`~pyhf_b200.plan.HostPlan`, without pyhf.

``plan.extract_plan`` reads the tables off a *live* ``pyhf.Model`` (the plugin route).  This
module builds the same tables straight from the workspace-style spec (``{"channels": [...],
"parameters": [...]}``) so that the C-ABI can be driven without the reference being importable
(stand-alone benchmarks, tests on a box without pyhf, other host languages).  It restates the
ordering and book-keeping rules of the reference's model builder:

* channels, samples and (name, type) modifiers sorted            (mixins.py:50-62)
* mega-samples: zero nominal where a sample is absent             (pdf.py:63-73)
* parameter sets registered per modifier *type* in the order of ``histfactory_set``
  (histosys, lumi, normfactor, normsys, shapefactor, shapesys, staterror --
  modifiers/__init__.py:39-47), inside a type by first appearance while walking sorted
  channels x sorted samples x sorted modifiers (pdf.py:138-147); staterror sets are registered in
  ``finalize`` in modifier order (modifiers/staterror.py:81-133)
* defaults of every parameter set and user overrides from
  ``spec["parameters"]`` (parameters/utils.py:10-67)
* shapesys factors (nom/unc)^2 and fixed bins (modifiers/shapesys.py:12-34); staterror
  sigma = sqrt(sum unc^2) / sum nom over the samples carrying the modifier, sigma = 0 -> fixed
  and sigma := 1 (modifiers/staterror.py:81-133)
* code4 polynomial coefficients a = A^-1 b                      (interpolators/code4.py:60-129)

``tests/test_compile.py`` asserts table-for-table equality with ``extract_plan`` on every golden
model (code4 coefficients to 1e-14).  Out of scope (raises :class:`PlanError`): custom modifier
types, interpolation code2, user ``parameters`` entries that change array lengths.
"""
from __future__ import annotations

import math

import numpy as np

from .plan import HS_CODE0, HS_CODE4P, SF_CODE1, SF_CODE4, SF_THETA, HostPlan, PlanError

__all__ = ["compile_spec"]

TYPE_ORDER = ("histosys", "lumi", "normfactor", "normsys", "shapefactor", "shapesys", "staterror")


def _code4_coefficients(up, dn, alpha0=1.0):
    """interpolators/code4.py:60-129 for one (hi/nom, lo/nom) pair -> a1..a6"""
    a0 = alpha0
    A_inv = np.array([
        [15.0 / (16 * a0), -15.0 / (16 * a0), -7.0 / 16.0, -7.0 / 16.0, 1.0 / 16 * a0, -1.0 / 16.0 * a0],
        [3.0 / (2 * a0**2), 3.0 / (2 * a0**2), -9.0 / (16 * a0), 9.0 / (16 * a0), 1.0 / 16, 1.0 / 16],
        [-5.0 / (8 * a0**3), 5.0 / (8 * a0**3), 5.0 / (8 * a0**2), 5.0 / (8 * a0**2), -1.0 / (8 * a0), 1.0 / (8 * a0)],
        [3.0 / (-2 * a0**4), 3.0 / (-2 * a0**4), -7.0 / (-8 * a0**3), 7.0 / (-8 * a0**3), -1.0 / (8 * a0**2), -1.0 / (8 * a0**2)],
        [3.0 / (16 * a0**5), -3.0 / (16 * a0**5), -3.0 / (16 * a0**4), -3.0 / (16 * a0**4), 1.0 / (16 * a0**3), -1.0 / (16 * a0**3)],
        [1.0 / (2 * a0**6), 1.0 / (2 * a0**6), -5.0 / (16 * a0**5), 5.0 / (16 * a0**5), 1.0 / (16 * a0**4), 1.0 / (16 * a0**4)],
    ])
    up_a, dn_a = up**a0, dn**a0
    b = np.array([up_a - 1.0, dn_a - 1.0, math.log(up) * up_a, -math.log(dn) * dn_a,
                  math.log(up) ** 2 * up_a, math.log(dn) ** 2 * dn_a])
    return A_inv @ b


def compile_spec(spec, poi_name="mu", modifier_settings=None, clip_sample_data=None,
                 clip_bin_data=None) -> HostPlan:
    """HistFactory spec -> HostPlan (same tables as ``extract_plan(pyhf.Model(spec, ...))``)."""
    # pdf.py:215-222: a user dict *replaces* the defaults (code4 / code4p); a type missing from it
    # falls back to its applier's own default (histosys.py:103 code0, normsys.py:72 code1)
    if modifier_settings is None:
        modifier_settings = {"normsys": {"interpcode": "code4"}, "histosys": {"interpcode": "code4p"}}
    hs_interp = modifier_settings.get("histosys", {}).get("interpcode", "code0")
    ns_interp = modifier_settings.get("normsys", {}).get("interpcode", "code1")
    if hs_interp not in ("code0", "code4p"):
        raise PlanError(f"histosys interpolation {hs_interp} is outside the hot-path scope")
    if ns_interp not in ("code1", "code4"):
        raise PlanError(f"normsys interpolation {ns_interp} is outside the hot-path scope")

    # ---------------------------------------------------------------- structure (mixins.py:50-97)
    chan_specs = {c["name"]: c for c in spec["channels"]}
    channels = sorted(chan_specs)
    nbins = {c: len(chan_specs[c]["samples"][0]["data"]) for c in channels}
    samples = sorted({s["name"] for c in spec["channels"] for s in c["samples"]})
    mods = sorted({(m["name"], m["type"]) for c in spec["channels"] for s in c["samples"]
                   for m in s["modifiers"]})
    for _, t in mods:
        if t not in TYPE_ORDER:
            raise PlanError(f"custom modifier types are not supported on the b200 path: {t}")
    S, NSEG = len(samples), len(channels)
    NB = sum(nbins[c] for c in channels)
    off = {}
    pos = 0
    for c in channels:
        off[c] = pos
        pos += nbins[c]
    bin_seg = np.concatenate([np.full(nbins[c], i, np.int32) for i, c in enumerate(channels)])
    helper = {c: {s["name"]: (s, {f"{m['type']}/{m['name']}": m for m in s["modifiers"]})
                  for s in chan_specs[c]["samples"]} for c in channels}
    # modifier types whose parameter set may not be shared (pdf.py:116-135; shapesys.py:40): a
    # shapesys name reused on another sample or channel is an error, not one shared set
    keys_seen = set()
    for c in spec["channels"]:
        for s in c["samples"]:
            mine = set()
            for m in s["modifiers"]:
                key = f"{m['type']}/{m['name']}"
                if m["type"] == "shapesys" and (key in keys_seen or key in mine):
                    raise PlanError(f"Trying to add paramset {key} on {s['name']} sample in {c['name']} channel "
                                    "but other paramsets exist with the same name.")
                mine.add(key)
            keys_seen |= mine

    nominal = np.zeros((S, NB))
    # per (modifier key, sample): data arrays over all bins
    data = {f"{t}/{n}": {"mask": np.zeros((S, NB), bool), "a": np.zeros((S, NB)), "b": np.zeros((S, NB))}
            for n, t in mods}
    required = {t: {} for t in TYPE_ORDER}  # type -> {parname: requirement}, insertion ordered

    for c in channels:
        sl = slice(off[c], off[c] + nbins[c])
        for si, s in enumerate(samples):
            samp, moddict = helper[c].get(s, (None, None))
            nom = np.asarray(samp["data"], float) if samp else np.zeros(nbins[c])
            if len(nom) != nbins[c]:
                raise PlanError(f"expected {nbins[c]} size sample data but got {len(nom)}")
            nominal[si, sl] = nom
            for n, t in mods:
                key = f"{t}/{n}"
                m = moddict.get(key) if moddict else None
                d = data[key]
                if t == "histosys":
                    d["a"][si, sl] = np.asarray(m["data"]["lo_data"], float) if m else 0.0
                    d["b"][si, sl] = np.asarray(m["data"]["hi_data"], float) if m else 0.0
                elif t == "normsys":
                    d["a"][si, sl] = float(m["data"]["lo"]) if m else 1.0
                    d["b"][si, sl] = float(m["data"]["hi"]) if m else 1.0
                elif t in ("shapesys", "staterror"):
                    d["a"][si, sl] = np.asarray(m["data"], float) if m else 0.0
                if not m:
                    continue
                d["mask"][si, sl] = True
                if t in ("histosys", "normsys"):
                    required[t].setdefault(n, dict(kind="normal", n=1, inits=[0.0], bounds=[[-5.0, 5.0]],
                                                   fixed=[False], auxdata=[0.0]))  # no "sigmas" key: not configurable
                elif t == "normfactor":
                    required[t].setdefault(n, dict(kind="free", n=1, inits=[1.0], bounds=[[0.0, 10.0]],
                                                   fixed=[False]))
                elif t == "lumi":
                    required[t].setdefault(n, dict(kind="normal", n=1, inits=None, bounds=None,
                                                   fixed=[False], auxdata=None, sigmas=None))
                elif t == "shapefactor":
                    required[t].setdefault(n, dict(kind="free", n=nbins[c], inits=[1.0] * nbins[c],
                                                   bounds=[[0.0, 10.0]] * nbins[c], fixed=[False] * nbins[c]))
                elif t == "shapesys":
                    unc = np.asarray(m["data"], float)
                    valid = (nom > 0) & (unc > 0)
                    with np.errstate(divide="ignore", invalid="ignore"):
                        fac = np.where(valid, nom**2 / unc**2, 1.0)
                    required[t].setdefault(n, dict(kind="poisson", n=nbins[c], inits=[1.0] * nbins[c],
                                                   bounds=[[1e-10, 10.0]] * nbins[c],
                                                   fixed=(~valid).tolist(), auxdata=fac.tolist(),
                                                   factors=fac.tolist()))
    # staterror parameter sets are created in finalize, in modifier order (staterror.py:81-133)
    for n, t in mods:
        if t != "staterror":
            continue
        d = data[f"{t}/{n}"]
        carrying = d["mask"].any(axis=1)
        nomsall = nominal[carrying].sum(axis=0)
        with np.errstate(divide="ignore", invalid="ignore"):
            rel2 = np.where(nomsall[None] > 0, (d["a"] / nomsall[None]) ** 2, 0.0).sum(axis=0)
        relerr = np.sqrt(rel2)
        mask = d["mask"][np.nonzero(carrying)[0][0]]
        sig = relerr[mask]
        fixed = (sig == 0).tolist()
        sig = np.where(sig == 0, 1.0, sig)
        k = int(mask.sum())
        required[t].setdefault(n, dict(kind="normal", n=k, inits=[1.0] * k, bounds=[[1e-10, 10.0]] * k,
                                       fixed=fixed, auxdata=[1.0] * k, sigmas=sig.tolist()))

    # ---------------------------------------------------------------- parameter sets (pdf.py:157-161)
    parsets = {}
    for t in TYPE_ORDER:
        for n, req in required[t].items():
            if n in parsets:
                # parameters/utils.py:26-46: every property of the requirements must agree
                a, b = parsets[n], req
                for k in ("kind", "n", "inits", "bounds", "auxdata", "factors", "sigmas", "fixed"):
                    if a.get(k, "undefined") != b.get(k, "undefined"):
                        raise PlanError(f"Multiple values for '{k}' were found for {n}. Use unique modifier "
                                        "names when constructing the pdf.")
                continue
            parsets[n] = dict(req)
    if not parsets:
        raise PlanError("No parameters specified for the Model.")
    user = {p["name"]: p for p in spec.get("parameters", [])}
    for n, ps in parsets.items():
        u = user.get(n, {})
        for key in ("inits", "bounds", "auxdata", "sigmas"):
            if key in u:
                if key not in ps:
                    raise PlanError(f"{n} does not use the {key} attribute.")
                if ps[key] is not None and len(u[key]) != len(ps[key]):
                    raise PlanError(f"Incorrect number of values for {key} of {n}")
                ps[key] = [list(b) for b in u[key]] if key == "bounds" else list(u[key])
        if "fixed" in u:
            ps["fixed"] = [bool(u["fixed"])] * ps["n"]
        if ps["inits"] is None or ps["bounds"] is None:
            raise PlanError(f"parameter set {n} needs 'inits'/'bounds' in spec['parameters'] (lumi)")
    par_start, pos = {}, 0
    for n, ps in parsets.items():
        par_start[n] = pos
        pos += ps["n"]
    P = pos
    init = np.concatenate([np.asarray(ps["inits"], float) for ps in parsets.values()])
    bounds = np.concatenate([np.asarray(ps["bounds"], float).reshape(-1, 2) for ps in parsets.values()])
    fixed = np.concatenate([np.asarray(ps["fixed"], bool) for ps in parsets.values()])

    # constraints: aux data in par_order of the constrained sets (pdf.py:43-55)
    auxdata, g_par, g_aux, g_sigma, p_par, p_aux, p_factor = [], [], [], [], [], [], []
    for n, ps in parsets.items():
        if ps["kind"] == "free":
            continue
        if ps["auxdata"] is None:
            raise PlanError(f"parameter set {n} needs 'auxdata' in spec['parameters']")
        a0 = len(auxdata)
        auxdata += list(ps["auxdata"])
        idx = range(par_start[n], par_start[n] + ps["n"])
        if ps["kind"] == "normal":
            g_par += list(idx)
            g_aux += list(range(a0, a0 + ps["n"]))
            g_sigma += list(ps["sigmas"]) if ps.get("sigmas") else [1.0] * ps["n"]
        else:
            p_par += list(idx)
            p_aux += list(range(a0, a0 + ps["n"]))
            p_factor += list(ps["factors"])

    names_of = lambda t: [n for n, tt in mods if tt == t]  # noqa: E731

    # ---------------------------------------------------------------- histosys tables
    hs_names = names_of("histosys")
    H = len(hs_names)
    hs_code = HS_CODE0 if (H and hs_interp == "code0") else HS_CODE4P
    hs_par = np.zeros(0, np.int32)
    hs_row_sample = np.zeros(0, np.int32)
    hs_t1 = hs_t2 = np.zeros((0, 0, NB))
    if H:
        lo = np.stack([data[f"histosys/{n}"]["a"] for n in hs_names])
        hi = np.stack([data[f"histosys/{n}"]["b"] for n in hs_names])
        mask = np.stack([data[f"histosys/{n}"]["mask"] for n in hs_names])
        d_up = hi - nominal[None]
        d_dn = nominal[None] - lo
        if hs_code == HS_CODE4P:
            t1, t2 = 0.5 * (d_up + d_dn), 0.0625 * (d_up - d_dn)
        else:
            t1, t2 = d_up, d_dn
        t1 = np.where(mask, t1, 0.0)
        t2 = np.where(mask, t2, 0.0)
        rows = np.nonzero(mask.any(axis=(0, 2)))[0]
        hs_row_sample = rows.astype(np.int32)
        hs_t1 = np.ascontiguousarray(t1[:, rows, :])
        hs_t2 = np.ascontiguousarray(t2[:, rows, :])
        hs_par = np.array([par_start[n] for n in hs_names], np.int32)

    # ---------------------------------------------------------------- scalar factors
    per_q = [[] for _ in range(S * NSEG)]
    first = [off[c] for c in channels]
    for t in ("normfactor", "lumi"):
        for n in names_of(t):
            m = data[f"{t}/{n}"]["mask"]
            for si in range(S):
                for ci in range(NSEG):
                    if m[si, first[ci]]:
                        per_q[si * NSEG + ci].append((par_start[n], SF_THETA, np.zeros(8)))
    for n in names_of("normsys"):
        d = data[f"normsys/{n}"]
        for si in range(S):
            for ci in range(NSEG):
                b0 = first[ci]
                if not d["mask"][si, b0]:
                    continue
                up, dn = float(d["b"][si, b0]), float(d["a"][si, b0])
                cf = np.zeros(8)
                cf[0] = math.log(up) if up > 0 else -math.inf
                cf[1] = math.log(dn) if dn > 0 else -math.inf
                if ns_interp == "code4":
                    cf[2:8] = _code4_coefficients(up, dn)
                    per_q[si * NSEG + ci].append((par_start[n], SF_CODE4, cf))
                else:
                    per_q[si * NSEG + ci].append((par_start[n], SF_CODE1, cf))
    sf_ptr = np.zeros(S * NSEG + 1, np.int32)
    sf_par, sf_kind, sf_coef = [], [], []
    for q, lst in enumerate(per_q):
        sf_ptr[q + 1] = sf_ptr[q] + len(lst)
        for par, kind, cf in lst:
            sf_par.append(par)
            sf_kind.append(kind)
            sf_coef.append(cf)

    # ---------------------------------------------------------------- per-bin factors
    cells = [[[] for _ in range(NB)] for _ in range(S)]
    chan_of_bin = np.concatenate([np.full(nbins[c], i) for i, c in enumerate(channels)])
    for t in ("shapefactor", "shapesys", "staterror"):
        for n in names_of(t):
            m = data[f"{t}/{n}"]["mask"]
            start, npar = par_start[n], parsets[n]["n"]
            access = np.zeros(NB, np.int64)
            if t == "shapefactor":  # shapefactor.py:174-180: bin j of every channel -> parameter j
                within = np.concatenate([np.arange(nbins[c]) for c in channels])
                access = np.where(within < npar, start + within, 0)
            else:  # shapesys.py:150-171 / staterror.py:171-192: mask of the last carrying sample
                smask = m[np.nonzero(m.any(axis=1))[0][-1]]
                access[smask] = start + np.arange(int(smask.sum()))
            for si, b in zip(*np.nonzero(m)):
                cells[si][b].append(int(access[b]))
    K = max((len(c) for row in cells for c in row), default=0)
    bf_par = -np.ones((K, S, NB), np.int32)
    for si in range(S):
        for b in range(NB):
            for k, p in enumerate(cells[si][b]):
                bf_par[k, si, b] = p
    del chan_of_bin

    poi_index = par_start[poi_name] if poi_name else -1
    n_sf = len(sf_par)
    return HostPlan(
        n_pars=P, n_bins=NB, n_samples=S, n_aux=len(auxdata), n_segs=NSEG, n_hs=H,
        n_hs_rows=len(hs_row_sample), hs_code=hs_code, n_sf=n_sf, n_bf_slots=K,
        n_gauss=len(g_par), n_pois=len(p_par), poi_index=poi_index,
        has_clip_sample=clip_sample_data is not None, has_clip_bin=clip_bin_data is not None,
        clip_sample=0.0 if clip_sample_data is None else float(clip_sample_data),
        clip_bin=0.0 if clip_bin_data is None else float(clip_bin_data),
        nominal=nominal, bin_seg=bin_seg, hs_par=hs_par, hs_row_sample=hs_row_sample,
        hs_t1=hs_t1, hs_t2=hs_t2, sf_ptr=sf_ptr, sf_par=sf_par, sf_kind=sf_kind,
        sf_coef=np.array(sf_coef, float).reshape(n_sf, 8), bf_par=bf_par,
        g_par=g_par, g_aux=g_aux, g_sigma=g_sigma, p_par=p_par, p_aux=p_aux, p_factor=p_factor,
        init=init, lo=bounds[:, 0], hi=bounds[:, 1], fixed=fixed, auxdata=np.asarray(auxdata, float),
        par_slices={n: (par_start[n], ps["n"]) for n, ps in parsets.items()},
    )
