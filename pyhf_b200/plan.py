"""Plan extraction: walk a live (unbatched) ``pyhf.Model`` and flatten it into the sparse,
device-friendly tables the CUDA kernels consume.

pyhf evaluates a model through dense ``(n_modifier, n_sample, batch, n_bin)`` tensors
(reference: ``src/pyhf/pdf.py:645-713``).  The plan stores the same information once, in the
form the arithmetic actually needs:

* ``nominal[S, NB]``                                   (``pdf.py:598-600``)
* histosys  -> two tables ``T1, T2 [H, R, NB]`` over the ``R`` samples that carry any histosys,
  with ``delta = c1(alpha)*T1 + c2(alpha)*T2`` (code4p: T1=S, T2=A of
  ``interpolators/code4p.py:30-45``; code0: T1=delta_up, T2=delta_dn of ``code0.py:37-38``);
  masked cells hold exact zeros (``modifiers/histosys.py:172``).
* scalar factors (normsys code1/code4, normfactor, lumi) -> one entry per
  (modifier, sample, channel) because those factors are constant inside a channel
  (``modifiers/normsys.py:31-39``, ``normfactor.py:94-115``, ``lumi.py:92-107``); code4
  polynomial coefficients are *read from the live interpolator* (``interpolators/code4.py:127``).
* per-bin factors (shapefactor, shapesys, staterror) -> ``bf_par[K, S, NB]`` flat parameter index
  per cell taken from the live ``_access_field`` + mask (``modifiers/shapesys.py:116-171``).
* constraint tables (``constraints.py:60-80``, ``:179-199``).

Nothing here touches the GPU; the arrays are handed to ``hf_plan_create`` by ``_lib.py``.
"""
from __future__ import annotations

import weakref

import numpy as np

__all__ = ["HostPlan", "extract_plan", "PlanError"]

SF_THETA, SF_CODE4, SF_CODE1 = 0, 1, 2
HS_CODE0, HS_CODE4P = 0, 4


class PlanError(Exception):
    """The model uses a feature outside the hot-path scope (mapped to pyhf.exceptions.Unsupported
    by the plugin layer)."""


def _np(x, dtype=None):
    """numpy view of a constant held by any pyhf backend (numpy array, list, torch tensor)."""
    if hasattr(x, "detach"):
        x = x.detach().cpu().numpy()
    return np.asarray(x, dtype=dtype)


class HostPlan:
    """Flat host-side tables of one HistFactory model (all numpy, C-contiguous)."""

    _int_fields = (
        "bin_seg", "hs_par", "hs_row_sample", "sf_ptr", "sf_par", "sf_kind", "bf_par",
        "g_par", "g_aux", "p_par", "p_aux",
    )
    _dbl_fields = (
        "nominal", "hs_t1", "hs_t2", "sf_coef", "g_sigma", "p_factor",
        "init", "lo", "hi", "auxdata",
    )

    def __init__(self, **kw):
        self.__dict__.update(kw)
        for k in self._int_fields:
            setattr(self, k, np.ascontiguousarray(getattr(self, k), dtype=np.int32))
        for k in self._dbl_fields:
            setattr(self, k, np.ascontiguousarray(getattr(self, k), dtype=np.float64))
        self.fixed = np.ascontiguousarray(self.fixed, dtype=np.uint8)

    # convenient sizes
    @property
    def n_data(self):
        return self.n_bins + self.n_aux

    def arrays(self):
        """dict of every table (used by the oracle, the golden fixtures and tests)."""
        keys = self._int_fields + self._dbl_fields + ("fixed",)
        d = {k: getattr(self, k) for k in keys}
        d["meta"] = np.array(
            [self.n_pars, self.n_bins, self.n_samples, self.n_aux, self.n_segs, self.n_hs,
             self.n_hs_rows, self.hs_code, self.n_sf, self.n_bf_slots, self.n_gauss, self.n_pois,
             self.poi_index, int(self.has_clip_sample), int(self.has_clip_bin)],
            dtype=np.int64,
        )
        d["clips"] = np.array([self.clip_sample, self.clip_bin], dtype=np.float64)
        return d

    @classmethod
    def from_arrays(cls, d):
        m = [int(v) for v in d["meta"]]
        kw = {k: d[k] for k in cls._int_fields + cls._dbl_fields + ("fixed",)}
        (kw["n_pars"], kw["n_bins"], kw["n_samples"], kw["n_aux"], kw["n_segs"], kw["n_hs"],
         kw["n_hs_rows"], kw["hs_code"], kw["n_sf"], kw["n_bf_slots"], kw["n_gauss"],
         kw["n_pois"], kw["poi_index"]) = m[:13]
        kw["has_clip_sample"], kw["has_clip_bin"] = bool(m[13]), bool(m[14])
        kw["clip_sample"], kw["clip_bin"] = float(d["clips"][0]), float(d["clips"][1])
        return cls(**kw)


def _mods_of_type(config, mtype):
    return [name for name, typ in config.modifiers if typ == mtype]


def _par_start(config, name):
    return int(config.par_map[name]["slice"].start)


def extract_plan(model) -> HostPlan:
    """Build a :class:`HostPlan` from a live unbatched ``pyhf.Model``.

    Attribute map: SURVEY.md Appendix B.  Raises :class:`PlanError` for anything outside the
    hot-path scope (batched models, interpolation code2, custom modifier sets).
    """
    config = model.config
    main = model.main_model
    if getattr(model, "batch_size", None) is not None:
        raise PlanError("plans are built from unbatched models; batching is the kernel grid")
    appliers = main.modifiers_appliers
    known = {"histosys", "lumi", "normfactor", "normsys", "shapefactor", "shapesys", "staterror"}
    extra = set(appliers) - known
    if extra:
        raise PlanError(f"custom modifier types are not supported on the b200 path: {sorted(extra)}")

    samples = list(config.samples)
    channels = list(config.channels)
    S = len(samples)
    nbins = [int(config.channel_nbins[c]) for c in channels]
    NB = int(sum(nbins))
    P = int(config.npars)
    bin_seg = np.repeat(np.arange(len(channels)), nbins).astype(np.int32)
    seg_first = np.concatenate([[0], np.cumsum(nbins)[:-1]]).astype(np.int64)
    NSEG = len(channels)

    nominal = _np(main._nominal_rates, np.float64)[0, :, 0, :]
    assert nominal.shape == (S, NB), nominal.shape

    # ---------------------------------------------------------------- histosys (additive)
    hs_names = _mods_of_type(config, "histosys")
    H = len(hs_names)
    hs_code = HS_CODE4P
    hs_par = np.zeros(0, np.int32)
    hs_row_sample = np.zeros(0, np.int32)
    hs_t1 = np.zeros((0, 0, NB))
    hs_t2 = np.zeros((0, 0, NB))
    if H:
        app = appliers["histosys"]
        code = app.interpcode
        if code not in ("code0", "code4p"):
            raise PlanError(f"histosys interpolation {code} is outside the hot-path scope")
        hs_code = HS_CODE0 if code == "code0" else HS_CODE4P
        hist = _np(app._histosys_histoset, np.float64)  # (H, S, 3, NB) = [lo, nom, hi]
        mask = _np(app._histosys_mask).astype(bool)[:, :, 0, :]  # (H, S, NB)
        d_up = hist[:, :, 2] - hist[:, :, 1]  # code4p.py:30 / code0.py:37
        d_dn = hist[:, :, 1] - hist[:, :, 0]  # code4p.py:31 / code0.py:38
        if hs_code == HS_CODE4P:
            t1 = 0.5 * (d_up + d_dn)  # S, code4p.py:44
            t2 = 0.0625 * (d_up - d_dn)  # A, code4p.py:45
        else:
            t1, t2 = d_up, d_dn
        t1 = np.where(mask, t1, 0.0)
        t2 = np.where(mask, t2, 0.0)
        rows = np.nonzero(mask.any(axis=(0, 2)))[0]
        hs_row_sample = rows.astype(np.int32)
        hs_t1 = np.ascontiguousarray(t1[:, rows, :])
        hs_t2 = np.ascontiguousarray(t2[:, rows, :])
        hs_par = np.array([_par_start(config, n) for n in hs_names], np.int32)
    R = len(hs_row_sample)

    # ---------------------------------------------------------------- scalar factors
    # one list of entries per (sample, channel); CSR over q = s * NSEG + seg
    per_q = [[] for _ in range(S * NSEG)]

    def _channel_constant(arr_msb, what):
        """arr (M, S, NB) -> (M, S, NSEG), asserting constancy inside every channel."""
        out = arr_msb[:, :, seg_first]
        if not np.array_equal(np.repeat(out, nbins, axis=2), arr_msb):
            raise PlanError(f"{what} varies inside a channel; outside the hot-path scope")
        return out

    for mtype in ("normfactor", "lumi"):
        names = _mods_of_type(config, mtype)
        if not names:
            continue
        app = appliers[mtype]
        mask = _np(getattr(app, f"_{mtype}_mask")).astype(bool)[:, :, 0, :]
        mseg = _channel_constant(mask, f"{mtype} mask")
        for m, name in enumerate(names):
            # lumi.py:103 broadcasts the single lumi parameter over all lumi modifiers
            par = _par_start(config, name)
            for s in range(S):
                for c in range(NSEG):
                    if mseg[m, s, c]:
                        per_q[s * NSEG + c].append((par, SF_THETA, np.zeros(8)))

    ns_names = _mods_of_type(config, "normsys")
    if ns_names:
        app = appliers["normsys"]
        code = app.interpcode
        if code not in ("code1", "code4"):
            raise PlanError(f"normsys interpolation {code} is outside the hot-path scope")
        hist = _np(app._normsys_histoset, np.float64)  # (M, S, 3, NB) = [lo, 1, hi]
        mask = _np(app._normsys_mask).astype(bool)[:, :, 0, :]
        mseg = _channel_constant(mask, "normsys mask")
        up = _channel_constant(hist[:, :, 2] / hist[:, :, 1], "normsys hi")  # code4.py:49-51
        dn = _channel_constant(hist[:, :, 0] / hist[:, :, 1], "normsys lo")  # code4.py:52-54
        if code == "code4":
            coef = _np(app.interpolator._coefficients, np.float64)  # (6, M, S, NB)
            cseg = np.stack([_channel_constant(coef[r], "code4 coefficients") for r in range(6)])
        for m, name in enumerate(ns_names):
            par = _par_start(config, name)
            for s in range(S):
                for c in range(NSEG):
                    if not mseg[m, s, c]:
                        continue
                    cf = np.zeros(8)
                    with np.errstate(divide="ignore", invalid="ignore"):
                        cf[0] = np.log(up[m, s, c])
                        cf[1] = np.log(dn[m, s, c])
                    if code == "code4":
                        cf[2:8] = cseg[:, m, s, c]
                        per_q[s * NSEG + c].append((par, SF_CODE4, cf))
                    else:
                        per_q[s * NSEG + c].append((par, SF_CODE1, cf))

    sf_ptr = np.zeros(S * NSEG + 1, np.int32)
    sf_par, sf_kind, sf_coef = [], [], []
    for q, lst in enumerate(per_q):
        sf_ptr[q + 1] = sf_ptr[q] + len(lst)
        for par, kind, cf in lst:
            sf_par.append(par)
            sf_kind.append(kind)
            sf_coef.append(cf)
    n_sf = len(sf_par)
    sf_coef = np.array(sf_coef, np.float64).reshape(n_sf, 8)

    # ---------------------------------------------------------------- per-bin factors
    cell_lists = [[[] for _ in range(NB)] for _ in range(S)]
    for mtype in ("shapefactor", "shapesys", "staterror"):
        names = _mods_of_type(config, mtype)
        if not names:
            continue
        app = appliers[mtype]
        mask = _np(getattr(app, f"_{mtype}_mask")).astype(bool)[:, :, 0, :]  # (M, S, NB)
        access = _np(app._access_field).astype(np.int64)[:, 0, :]  # (M, NB) flat par index
        for m in range(len(names)):
            ss, bb = np.nonzero(mask[m])
            for s, b in zip(ss.tolist(), bb.tolist()):
                cell_lists[s][b].append(int(access[m, b]))
    K = max((len(c) for row in cell_lists for c in row), default=0)
    bf_par = -np.ones((K, S, NB), np.int32)
    for s in range(S):
        for b in range(NB):
            for k, p in enumerate(cell_lists[s][b]):
                bf_par[k, s, b] = p

    # ---------------------------------------------------------------- constraints
    cm = model.constraint_model
    g_par = g_aux = p_par = p_aux = np.zeros(0, np.int32)
    g_sigma = p_factor = np.zeros(0)
    cg = cm.constraints_gaussian
    if cg._access_field is not None:
        g_par = _np(cg._access_field).astype(np.int64).reshape(-1)
        g_aux = _np(cg._normal_data).astype(np.int64).reshape(-1)
        sig = _np(cg._sigmas, np.float64)
        g_sigma = sig.reshape(-1)[: len(g_par)]
    cp = cm.constraints_poisson
    if cp._access_field is not None:
        p_par = _np(cp._access_field).astype(np.int64).reshape(-1)
        p_aux = _np(cp._poisson_data).astype(np.int64).reshape(-1)
        p_factor = _np(cp._batched_factors, np.float64).reshape(-1)[: len(p_par)]

    clip_s = main.clip_sample_data
    clip_b = main.clip_bin_data

    plan = HostPlan(
        n_pars=P, n_bins=NB, n_samples=S, n_aux=int(config.nauxdata), n_segs=NSEG,
        n_hs=H, n_hs_rows=R, hs_code=hs_code, n_sf=n_sf, n_bf_slots=K,
        n_gauss=len(g_par), n_pois=len(p_par),
        poi_index=-1 if config.poi_index is None else int(config.poi_index),
        has_clip_sample=clip_s is not None, has_clip_bin=clip_b is not None,
        clip_sample=0.0 if clip_s is None else float(clip_s),
        clip_bin=0.0 if clip_b is None else float(clip_b),
        nominal=nominal, bin_seg=bin_seg,
        hs_par=hs_par, hs_row_sample=hs_row_sample, hs_t1=hs_t1, hs_t2=hs_t2,
        sf_ptr=sf_ptr, sf_par=sf_par, sf_kind=sf_kind, sf_coef=sf_coef, bf_par=bf_par,
        g_par=g_par, g_aux=g_aux, g_sigma=g_sigma, p_par=p_par, p_aux=p_aux, p_factor=p_factor,
        init=np.asarray(config.suggested_init(), np.float64),
        lo=np.array([b[0] for b in config.suggested_bounds()], np.float64),
        hi=np.array([b[1] for b in config.suggested_bounds()], np.float64),
        fixed=np.asarray(config.suggested_fixed(), bool),
        auxdata=np.asarray(config.auxdata, np.float64),
    )
    assert plan.n_aux == len(plan.auxdata)
    return plan


# ------------------------------------------------------------------------------------------
# per-model cache that does not keep the model alive (mirrors the weak-ref pattern of
# reference events.py:37-49); the tables are backend-independent so no invalidation on
# ``tensorlib_changed`` is needed for the host plan itself.
_cache: "weakref.WeakKeyDictionary" = weakref.WeakKeyDictionary()


def cached_plan(model) -> HostPlan:
    try:
        return _cache[model]
    except KeyError:
        plan = extract_plan(model)
        _cache[model] = plan
        return plan
    except TypeError:  # unhashable / not weak-referenceable model objects
        return extract_plan(model)
