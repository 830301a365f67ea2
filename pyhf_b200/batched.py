"""Explicit batched API over the C-ABI (device tensors in, device tensors out).

pyhf has no call for "evaluate B parameter points" or "fit N datasets" (its ``batch_size`` is a
build-time property of the model, reference pdf.py:598-600), so the plugin objects
(``b200_backend`` / ``b200_optimizer``) and users who want the throughput call this module:

    plan = pyhf_b200.plan_for(model)              # compile once per model / device
    val, grad = plan.logpdf_grad(pars[B, P], data)           # K2
    res = plan.fit(data[N, D], fixed=mask)                   # K3
    toys = plan.sample_toys(pars, n=100_000, seed=1)         # K4
    out = qmu_tilde_batch(plan, 1.0, toys)                   # K3 + K5

torch is only the owner of device memory and of the CUDA stream.
"""
from __future__ import annotations

import ctypes as C
import weakref
from dataclasses import dataclass

import numpy as np
import torch

from . import _lib
from .plan import HostPlan, cached_plan

__all__ = ["DevicePlan", "plan_for", "FitResult", "qmu_tilde_batch", "toy_hypotest",
           "sample_poisson", "sample_normal", "teststat", "pvalue_counts", "fp64_peak_tflops"]

F64 = torch.float64


def _stream():
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def _dev_tensor(x, device, dtype=F64):
    if isinstance(x, torch.Tensor):
        t = x.detach()
        if t.device != device or t.dtype != dtype:
            t = t.to(device=device, dtype=dtype)
        return t.contiguous()
    return torch.as_tensor(np.ascontiguousarray(np.asarray(x)), device=device).to(dtype).contiguous()


def _ptr(t):
    return C.c_void_p(t.data_ptr())


@dataclass
class FitResult:
    x: torch.Tensor        # [N, P] fitted parameters (fixed ones at their given value)
    fun: torch.Tensor      # [N]    2*NLL at x (all constant terms included, = twice_nll)
    status: torch.Tensor   # [N]    0 converged / 1 iteration limit / 2 numerical failure
    niter: torch.Tensor    # [N]


class DevicePlan:
    """A model compiled into device tables (``hf_plan``)."""

    def __init__(self, host: HostPlan, device=None):
        lib = _lib.require()
        if not torch.cuda.is_available():
            raise RuntimeError("pyhf_b200 needs a CUDA device; there is no CPU fallback")
        self.device = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
        self.host = host
        self._keep = []
        d = _lib.PlanDesc()
        for k in ("n_pars", "n_bins", "n_samples", "n_aux", "n_segs", "n_hs", "n_hs_rows",
                  "hs_code", "n_sf", "n_bf_slots", "n_gauss", "n_pois"):
            setattr(d, k, int(getattr(host, k)))
        d.has_clip_sample = int(host.has_clip_sample)
        d.has_clip_bin = int(host.has_clip_bin)
        d.clip_sample = float(host.clip_sample)
        d.clip_bin = float(host.clip_bin)
        for k in ("nominal", "bin_seg", "hs_par", "hs_row_sample", "hs_t1", "hs_t2", "sf_ptr",
                  "sf_par", "sf_kind", "sf_coef", "bf_par", "g_par", "g_aux", "g_sigma", "p_par",
                  "p_aux", "p_factor"):
            arr = np.ascontiguousarray(getattr(host, k))
            if arr.size == 0:  # keep a valid address for empty tables
                arr = np.zeros(1, dtype=arr.dtype)
            self._keep.append(arr)
            setattr(d, k, arr.ctypes.data)
        handle = C.c_void_p()
        with torch.cuda.device(self.device):
            _lib.check(lib.hf_plan_create(C.byref(d), C.byref(handle)))
        self._h = handle
        self._lib = lib
        self._finalizer = weakref.finalize(self, lib.hf_plan_destroy, handle)
        self.P = host.n_pars
        self.D = host.n_bins + host.n_aux
        dev = self.device
        self.init = torch.as_tensor(host.init, device=dev)
        self.lo = torch.as_tensor(host.lo, device=dev)
        self.hi = torch.as_tensor(host.hi, device=dev)
        self.fixed = torch.as_tensor(host.fixed.astype(np.uint8), device=dev)

    # ------------------------------------------------------------------ evaluation
    def _prep(self, pars, data):
        pars = _dev_tensor(pars, self.device)
        data = _dev_tensor(data, self.device)
        if pars.dim() == 1:
            pars = pars.reshape(1, -1)
        if data.dim() == 1:
            data = data.reshape(1, -1)
        B = pars.shape[0]
        if pars.shape[1] != self.P:
            raise ValueError(f"pars must have {self.P} columns, got {tuple(pars.shape)}")
        if data.shape[1] != self.D or data.shape[0] not in (1, B):
            raise ValueError(f"data must be [{self.D}] or [B, {self.D}], got {tuple(data.shape)}")
        return pars, data, B

    def logpdf(self, pars, data):
        """K1: ln L for every row of pars (Model.logpdf, reference pdf.py:957-996)."""
        pars, data, B = self._prep(pars, data)
        out = torch.empty(B, dtype=F64, device=self.device)
        with torch.cuda.device(self.device):
            _lib.check(self._lib.hf_logpdf(self._h, _ptr(pars), _ptr(data), data.shape[0], B,
                                           _ptr(out), _stream()))
        return out

    def logpdf_grad(self, pars, data):
        """K2: (ln L [B], d lnL/d theta [B, P])."""
        pars, data, B = self._prep(pars, data)
        out = torch.empty(B, dtype=F64, device=self.device)
        grad = torch.empty(B, self.P, dtype=F64, device=self.device)
        with torch.cuda.device(self.device):
            _lib.check(self._lib.hf_logpdf_grad(self._h, _ptr(pars), _ptr(data), data.shape[0], B,
                                                _ptr(out), _ptr(grad), _stream()))
        return out, grad

    def logpdf_grad_host(self, pars_host, data_host, out_host, grad_host):
        """End-to-end call with HOST buffers (pinned torch CPU tensors): H2D, K2, D2H, sync."""
        B = pars_host.shape[0]
        rows = 1 if data_host.dim() == 1 else data_host.shape[0]
        with torch.cuda.device(self.device):
            _lib.check(self._lib.hf_logpdf_grad_host(
                self._h, _ptr(pars_host), _ptr(data_host), rows, B, _ptr(out_host),
                _ptr(grad_host) if grad_host is not None else None, _stream()))

    def expected_data(self, pars):
        pars = _dev_tensor(pars, self.device)
        single = pars.dim() == 1
        if single:
            pars = pars.reshape(1, -1)
        out = torch.empty(pars.shape[0], self.D, dtype=F64, device=self.device)
        with torch.cuda.device(self.device):
            _lib.check(self._lib.hf_expected_data(self._h, _ptr(pars), pars.shape[0], _ptr(out), _stream()))
        return out[0] if single else out

    # ------------------------------------------------------------------ fits
    def fit(self, data, init=None, lo=None, hi=None, fixed=None, max_iter=None, tol_edm=None,
            tol_step=None, verbose=0, fixed_alt=None, use_alt=None, data_map=None, masks=None,
            mask_id=None):
        """K3: N independent bounded fits of 2*NLL.  data [N, D] or [D]; init [N, P] or [P]
        (fixed parameters keep their init value).

        Mixed batches (hf_fit_batch_mixed): ``data_map`` [N] int32 picks the data row of each fit
        (data may then have any number of rows); fits with ``use_alt`` [N] != 0 use the mask
        ``fixed_alt`` [P] instead of ``fixed``.  ``masks`` [n_masks, P] with ``mask_id`` [N]
        (hf_fit_batch_masks) is the general form: fit i uses row mask_id[i]; it replaces
        ``fixed`` / ``fixed_alt`` / ``use_alt``."""
        dev = self.device
        data = _dev_tensor(data, dev)
        if data.dim() == 1:
            data = data.reshape(1, -1)
        init = self.init if init is None else _dev_tensor(init, dev)
        if init.dim() == 1:
            init = init.reshape(1, -1)
        lo = self.lo if lo is None else _dev_tensor(lo, dev)
        hi = self.hi if hi is None else _dev_tensor(hi, dev)
        fixed = self.fixed if fixed is None else _dev_tensor(fixed, dev, torch.uint8)
        if data_map is not None:
            data_map = _dev_tensor(data_map, dev, torch.int32)
            N = data_map.numel()
        else:
            N = max(data.shape[0], init.shape[0])
        if (masks is None) != (mask_id is None):
            raise ValueError("masks and mask_id go together")
        if masks is not None:
            if fixed_alt is not None or use_alt is not None:
                raise ValueError("a mask table replaces fixed_alt / use_alt")
            masks = _dev_tensor(masks, dev, torch.uint8).reshape(-1, self.P).contiguous()
            mask_id = _dev_tensor(mask_id, dev, torch.uint8).reshape(-1)
            if not 1 <= masks.shape[0] <= 256:
                raise ValueError("the mask table holds 1..256 rows")
            if data_map is None:
                N = max(N, mask_id.numel())
            if mask_id.numel() != N or (N and int(mask_id.max()) >= masks.shape[0]):
                raise ValueError("bad mask_id: one id < n_masks per fit")
        if (fixed_alt is None) != (use_alt is None):
            raise ValueError("fixed_alt and use_alt go together")
        if fixed_alt is not None:
            fixed_alt = _dev_tensor(fixed_alt, dev, torch.uint8)
            use_alt = _dev_tensor(use_alt, dev, torch.uint8)
            if use_alt.numel() != N or fixed_alt.numel() != self.P:
                raise ValueError("bad fixed_alt / use_alt shape")
        if data.shape[1] != self.D or init.shape[1] != self.P:
            raise ValueError("bad data / init shape")
        if (data_map is None and data.shape[0] not in (1, N)) or init.shape[0] not in (1, N):
            raise ValueError("data and init must have 1 or N rows")
        opts = _lib.FitOpts()
        self._lib.hf_fit_default_opts(C.byref(opts))
        if max_iter is not None:
            opts.max_iter = int(max_iter)
        if tol_edm is not None:
            opts.tol_edm = float(tol_edm)
        if tol_step is not None:
            opts.tol_step = float(tol_step)
        opts.verbose = int(verbose)
        x = torch.empty(N, self.P, dtype=F64, device=dev)
        fun = torch.empty(N, dtype=F64, device=dev)
        status = torch.empty(N, dtype=torch.int32, device=dev)
        niter = torch.empty(N, dtype=torch.int32, device=dev)
        if masks is not None:
            with torch.cuda.device(dev):
                _lib.check(self._lib.hf_fit_batch_masks(
                    self._h, _ptr(data), data.shape[0], None if data_map is None else _ptr(data_map),
                    _ptr(init), init.shape[0], _ptr(lo), _ptr(hi), _ptr(masks), masks.shape[0],
                    _ptr(mask_id), N, C.byref(opts), _ptr(x), _ptr(fun), _ptr(status), _ptr(niter),
                    _stream()))
            return FitResult(x, fun, status, niter)
        with torch.cuda.device(dev):
            _lib.check(self._lib.hf_fit_batch_mixed(
                self._h, _ptr(data), data.shape[0], None if data_map is None else _ptr(data_map),
                _ptr(init), init.shape[0], _ptr(lo), _ptr(hi), _ptr(fixed),
                None if fixed_alt is None else _ptr(fixed_alt),
                None if use_alt is None else _ptr(use_alt), N, C.byref(opts), _ptr(x), _ptr(fun),
                _ptr(status), _ptr(niter), _stream()))
        return FitResult(x, fun, status, niter)

    def fit_multistart(self, data, starts, lo=None, hi=None, fixed=None, **fit_kw):
        """Every data row fitted from each of the K start points ``starts`` [K, P] in ONE lock-step
        batch (K * N fits sharing the N data rows); per row the lowest converged minimum wins.

        HistFactory likelihoods can be multimodal -- the reference's own validation model
        2bin_2channel_coupledhistosys (tests/test_validation.py:620-690) has two basins in the
        coupled histosys parameter, and SLSQP and a Newton method started at suggested_init end in
        different ones.  No counterpart in the reference (a fit there has one start)."""
        dev = self.device
        data = _dev_tensor(data, dev)
        if data.dim() == 1:
            data = data.reshape(1, -1)
        starts = _dev_tensor(starts, dev).reshape(-1, self.P)
        N, K = data.shape[0], starts.shape[0]
        rows = torch.arange(N, dtype=torch.int32, device=dev).repeat(K)
        res = self.fit(data, starts.repeat_interleave(N, dim=0), lo, hi, fixed, data_map=rows, **fit_kw)
        fun = torch.where(res.status == 0, res.fun, torch.full_like(res.fun, float("inf"))).reshape(K, N)
        best = torch.argmin(fun, dim=0)                       # [N] winning start per row
        pick = best.to(torch.int64) * N + torch.arange(N, device=dev)
        return FitResult(res.x[pick], res.fun[pick], res.status[pick], res.niter[pick])

    # ------------------------------------------------------------------ curvature
    def hessian(self, pars, data, lo=None, hi=None, fixed=None, method="fd"):
        """Hessian of 2*NLL at pars [B, P] (or [P]); rows/columns of fixed parameters are the
        identity's.  ``method``: "fd" = central differences of the analytic gradient (hf_hessian);
        "analytic" / "fisher" = the closed form the minimiser uses (hf_hessian_analytic: exact
        Hessian / Fisher matrix = Hessian on the Asimov data of each point)."""
        dev = self.device
        pars = _dev_tensor(pars, dev)
        single = pars.dim() == 1
        if single:
            pars = pars.reshape(1, -1)
        data = _dev_tensor(data, dev)
        if data.dim() == 1:
            data = data.reshape(1, -1)
        lo = self.lo if lo is None else _dev_tensor(lo, dev)
        hi = self.hi if hi is None else _dev_tensor(hi, dev)
        fixed = self.fixed if fixed is None else _dev_tensor(fixed, dev, torch.uint8)
        B = pars.shape[0]
        out = torch.empty(B, self.P, self.P, dtype=F64, device=dev)
        with torch.cuda.device(dev):
            if method == "fd":
                _lib.check(self._lib.hf_hessian(self._h, _ptr(pars), _ptr(data), data.shape[0], B, _ptr(lo),
                                                _ptr(hi), _ptr(fixed), _ptr(out), _stream()))
            elif method in ("analytic", "fisher"):
                _lib.check(self._lib.hf_hessian_analytic(self._h, _ptr(pars), _ptr(data), data.shape[0], B,
                                                         _ptr(fixed), 2 if method == "analytic" else 1,
                                                         _ptr(out), _stream()))
            else:
                raise ValueError(f"unknown Hessian method {method}")
        return out[0] if single else out

    def covariance(self, pars, data, lo=None, hi=None, fixed=None):
        """(covariance, uncertainties, correlations) of the parameters at a minimum: covariance =
        2 * inverse(Hessian of 2*NLL) on the parameters that are free AND not on a bound (what
        MINUIT reports with errordef = 1 on 2*NLL, reference optimize/opt_minuit.py:136-152); the
        others get zero uncertainty (optimize/mixins.py:95-103) and zero correlation rows."""
        dev = self.device
        x = _dev_tensor(pars, dev).reshape(-1)
        lo_t = self.lo if lo is None else _dev_tensor(lo, dev)
        hi_t = self.hi if hi is None else _dev_tensor(hi, dev)
        fx = (self.fixed if fixed is None else _dev_tensor(fixed, dev, torch.uint8)).bool()
        H = self.hessian(x, data, lo, hi, fixed)
        free = ~fx & (x > lo_t) & (x < hi_t)
        idx = torch.nonzero(free).reshape(-1)
        cov = torch.zeros(self.P, self.P, dtype=F64, device=dev)
        if idx.numel():
            sub = 2.0 * torch.linalg.inv(H[idx][:, idx])
            cov[idx.unsqueeze(1), idx.unsqueeze(0)] = sub
        unc = torch.sqrt(torch.clamp(torch.diagonal(cov), min=0.0))
        denom = torch.where(unc > 0, unc, torch.ones_like(unc))
        corr = cov / denom.unsqueeze(0) / denom.unsqueeze(1)
        return cov, unc, corr

    # ------------------------------------------------------------------ toys
    def sample_toys(self, pars, n, seed=0, offset=0):
        """K4: n toys ~ Model.make_pdf(pars).sample((n,)) (reference probability.py:263-274)."""
        pars = _dev_tensor(pars, self.device).reshape(-1)
        out = torch.empty(int(n), self.D, dtype=F64, device=self.device)
        with torch.cuda.device(self.device):
            _lib.check(self._lib.hf_sample_toys(self._h, _ptr(pars), int(seed), int(offset), int(n),
                                                _ptr(out), _stream()))
        return out

    @property
    def launch_count(self):
        return int(self._lib.hf_plan_launch_count(self._h))

    def fit_kernel_info(self):
        """launch shape of the most recent fit batch on this plan (hf_plan_fit_info): which form of
        K3 ran, persistent grid, CTA size, resident CTAs per SM, shared memory per CTA and the
        theoretical occupancy that follows (resident warps / 64 per SM)"""
        info = (C.c_int32 * 8)()
        _lib.check(self._lib.hf_plan_fit_info(self._h, info))
        solo = bool(info[0])
        out = {"form": "per-fit kernel (one CTA per fit, hf_fit_solo.cu)" if solo
               else "lock-step driver (hf_fit.cu: one batched K2 launch per round)"}
        if solo:
            out.update(grid=int(info[1]), threads_per_cta=int(info[2]), ctas_per_sm=int(info[3]),
                       smem_bytes_per_cta=int(info[4]), sms=int(info[5]),
                       low_rank_per_bin_block=bool(info[6]),
                       occupancy_theoretical=info[2] * info[3] / 2048.0)
        return out

    def set_timing(self, enable=True):
        """bracket every launch of the dominant kernel with CUDA events (bench.py roofline)"""
        _lib.check(self._lib.hf_plan_set_timing(self._h, int(bool(enable))))

    def main_kernel_ms(self):
        """(summed milliseconds, launches) of hf_main_kernel since the last call"""
        ms, n = C.c_double(0.0), C.c_int64(0)
        _lib.check(self._lib.hf_plan_main_kernel_ms(self._h, C.byref(ms), C.byref(n)))
        return ms.value, n.value


# ---------------------------------------------------------------------- per-model cache
_dev_cache: "weakref.WeakKeyDictionary" = weakref.WeakKeyDictionary()


def plan_for(model, device=None) -> DevicePlan:
    """DevicePlan of a live ``pyhf.Model`` (cached with a weak reference to the model)."""
    dev = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
    try:
        per_model = _dev_cache.setdefault(model, {})
    except TypeError:
        return DevicePlan(cached_plan(model), dev)
    key = str(dev)
    if key not in per_model:
        per_model[key] = DevicePlan(cached_plan(model), dev)
    return per_model[key]


# ---------------------------------------------------------------------- element samplers (backend)
def sample_poisson(backend, rate, sample_shape):
    lib = _lib.require()
    rate = _dev_tensor(rate, backend.device)
    n = int(np.prod(sample_shape)) if len(sample_shape) else 1
    m = rate.numel()
    out = torch.empty((n, m), dtype=F64, device=backend.device)
    off = backend._toy_offset
    backend._toy_offset += n
    _lib.check(lib.hf_sample_poisson(_ptr(rate.reshape(-1)), m, n, backend.seed, off, _ptr(out), _stream()))
    return out.reshape(tuple(sample_shape) + tuple(rate.shape))


def sample_normal(backend, loc, scale, sample_shape):
    lib = _lib.require()
    loc, scale = torch.broadcast_tensors(_dev_tensor(loc, backend.device), _dev_tensor(scale, backend.device))
    loc, scale = loc.contiguous(), scale.contiguous()
    n = int(np.prod(sample_shape)) if len(sample_shape) else 1
    m = loc.numel()
    out = torch.empty((n, m), dtype=F64, device=backend.device)
    off = backend._toy_offset
    backend._toy_offset += n
    _lib.check(lib.hf_sample_normal(_ptr(loc.reshape(-1)), _ptr(scale.reshape(-1)), m, n,
                                    backend.seed ^ 0x5DEECE66D, off, _ptr(out), _stream()))
    return out.reshape(tuple(sample_shape) + tuple(loc.shape))


# ---------------------------------------------------------------------- test statistics
_MODES = {"q": 0, "qtilde": 0, "q0": 1, "t": 2, "ttilde": 2}


def teststat(fun_fixed, fun_free, muhat, mu, mode="qtilde"):
    """K5: q = clip(2NLL_fixed - 2NLL_free, 0), zeroed where muhat > mu (q / q~) or muhat < 0 (q0)
    (reference infer/test_statistics.py:16-60, :507-520)."""
    lib = _lib.require()
    n = fun_fixed.numel()
    q = torch.empty_like(fun_fixed)
    _lib.check(lib.hf_teststat(_ptr(fun_fixed), _ptr(fun_free), _ptr(muhat.contiguous()), float(mu),
                               _MODES[mode], n, _ptr(q), _stream()))
    return q


def pvalue_counts(samples, values, assume_sorted=False):
    """#{samples >= value} per value (EmpiricalDistribution.pvalue, calculators.py:632-640).
    Few values: one O(N M) sweep over the unsorted samples; many values (the expected band asks for
    the p-values of every toy): sort once, then a binary search per value."""
    lib = _lib.require()
    samples = samples.contiguous()
    values = values.contiguous()
    counts = torch.empty(values.numel(), dtype=torch.int64, device=samples.device)
    if assume_sorted or values.numel() * samples.numel() > (1 << 24):
        srt = samples if assume_sorted else torch.sort(samples).values
        _lib.check(lib.hf_pvalue_counts_sorted(_ptr(srt), srt.numel(), _ptr(values), values.numel(),
                                               _ptr(counts), _stream()))
    else:
        _lib.check(lib.hf_pvalue_counts(_ptr(samples), samples.numel(), _ptr(values), values.numel(),
                                        _ptr(counts), _stream()))
    return counts


_NORMAL_PERCENTILES = (2.27501319, 15.86552539, 50.0, 84.13447461, 97.72498681)  # calculators.py:966-968


def _percentiles_sorted(srt, qs):
    """numpy.percentile(..., method="linear") on an ascending 1-D device tensor"""
    n = srt.numel()
    pos = torch.as_tensor(qs, dtype=F64, device=srt.device) / 100.0 * (n - 1)
    lo = torch.floor(pos).long().clamp_(0, n - 1)
    hi = (lo + 1).clamp_(max=n - 1)
    g = pos - lo.to(F64)
    a, b = srt[lo], srt[hi]
    # numpy's _lerp: a + (b - a) * g, evaluated from the far end for g >= 0.5
    return torch.where(g >= 0.5, b - (b - a) * (1.0 - g), a + (b - a) * g)


def empirical_expected_pvalues(q_sig, q_bkg):
    """ToyCalculator.expected_pvalues (calculators.py:924-973): the p-values of EVERY
    background-only toy under both empirical distributions, then the (-2..2 sigma) percentiles of
    each.  O(N log N) (two sorts + binary searches) instead of the reference's O(N^2).
    Returns (CLsb_exp, CLb_exp, CLs_exp), each a device tensor [5]."""
    s_sig = torch.sort(q_sig.reshape(-1)).values
    s_bkg = torch.sort(q_bkg.reshape(-1)).values
    clsb = pvalue_counts(s_sig, s_bkg, assume_sorted=True).to(F64) / s_sig.numel()
    clb = pvalue_counts(s_bkg, s_bkg, assume_sorted=True).to(F64) / s_bkg.numel()
    cls = clsb / clb
    return tuple(_percentiles_sorted(torch.sort(v).values, _NORMAL_PERCENTILES) for v in (clsb, clb, cls))


def empirical_expected_value(samples, nsigma):
    """EmpiricalDistribution.expected_value (calculators.py:642-699)"""
    q = torch.special.ndtr(torch.as_tensor(nsigma, dtype=F64, device=samples.device).reshape(-1)) * 100.0
    return _percentiles_sorted(torch.sort(samples.reshape(-1)).values, q)


def _view(r: "FitResult", sl) -> "FitResult":
    return FitResult(r.x[sl], r.fun[sl], r.status[sl], r.niter[sl])


def qmu_tilde_batch(plan: DevicePlan, mu, data, init=None, lo=None, hi=None, fixed=None,
                    mode="qtilde", warm_start=False, **fit_kw):
    """Test statistic of every row of ``data``: one fixed-POI fit and one free fit per row
    (reference _tmu_like, infer/test_statistics.py:38-60).  Both fits of all rows run as ONE
    lock-step batch (hf_fit_batch_mixed) unless ``warm_start`` asks for the free fit to start from
    the constrained one.  Returns a dict of device tensors."""
    poi = plan.host.poi_index
    if poi < 0:
        raise ValueError("the model has no POI")
    dev = plan.device
    init = plan.init if init is None else _dev_tensor(init, dev)
    fixed = plan.fixed if fixed is None else _dev_tensor(fixed, dev, torch.uint8)
    data = _dev_tensor(data, dev)
    if data.dim() == 1:
        data = data.reshape(1, -1)
    n = data.shape[0]
    init_fixed = init.clone()
    init_fixed[..., poi] = float(mu) if mode != "q0" else 0.0
    fixed_poi = fixed.clone()
    fixed_poi[poi] = 1
    if warm_start:
        rf = plan.fit(data, init_fixed, lo, hi, fixed_poi, **fit_kw)
        ru = plan.fit(data, rf.x, lo, hi, fixed, **fit_kw)
    else:
        rows = torch.arange(n, dtype=torch.int32, device=dev).repeat(2)
        use_alt = torch.zeros(2 * n, dtype=torch.uint8, device=dev)
        use_alt[:n] = 1
        i_f = init_fixed.reshape(-1, plan.P).expand(n, -1) if init_fixed.dim() == 1 or init_fixed.shape[0] == 1 else init_fixed
        i_u = init.reshape(-1, plan.P).expand(n, -1) if init.dim() == 1 or init.shape[0] == 1 else init
        both = plan.fit(data, torch.cat([i_f, i_u]), lo, hi, fixed, fixed_alt=fixed_poi, use_alt=use_alt,
                        data_map=rows, **fit_kw)
        rf, ru = _view(both, slice(0, n)), _view(both, slice(n, 2 * n))
    muhat = ru.x[:, poi].contiguous()
    q = teststat(rf.fun.contiguous(), ru.fun.contiguous(), muhat, float(mu), mode)
    return dict(q=q, muhat=muhat, fixed=rf, free=ru)


def toy_hypotest(plan: DevicePlan, mu_test, data_obs, ntoys, seed=0, toy_offset=0, init=None,
                 lo=None, hi=None, fixed=None, mode="qtilde", gather=None, profile=None,
                 return_expected_set=False, max_failed_fraction=None, **fit_kw):
    """Toy-based CLs for one signal strength: the arithmetic of ToyCalculator.distributions /
    pvalues (reference infer/calculators.py:754-915) with every fit batched on the device: the
    4 * ntoys toy fits (2 hypotheses x {POI fixed, free}) are one lock-step batch.

    ``gather`` (optional) maps a per-rank [2, n] tensor to the all-rank [2, n_total] tensor (see dist.py); toys
    [toy_offset, toy_offset + ntoys) are this rank's shard.  ``profile`` (optional dict) receives
    the wall time of each phase (adds a device synchronisation per phase).
    ``return_expected_set``: also the (-2..2 sigma) band of CLsb / CLb / CLs over the
    background-only toys (``expected``), the O(N log N) form of calculators.py:924-973.
    Toys with a failed fit are excluded from the p-values and counted in ``n_failed`` (global over
    all ranks); ``max_failed_fraction`` turns more than that fraction into an error."""
    import time

    dev = plan.device

    def mark(name, _t=[None]):
        if profile is None:
            return
        torch.cuda.synchronize(dev)
        now = time.perf_counter()
        if _t[0] is not None:
            profile[name] = profile.get(name, 0.0) + now - _t[0]
        _t[0] = now

    mark("start")
    data_obs = _dev_tensor(data_obs, dev).reshape(1, -1)
    init = plan.init if init is None else _dev_tensor(init, dev)
    fixed = plan.fixed if fixed is None else _dev_tensor(fixed, dev, torch.uint8)
    poi = plan.host.poi_index
    # observed test statistic and the generation points (conditional fits at mu_test and at the
    # background-only value, calculators.py:808-817): 4 fits on the observed data, one batch
    bkg_mu = 1.0 if mode == "q0" else 0.0
    fp = fixed.clone()
    fp[poi] = 1
    i_f = init.clone(); i_f[poi] = float(mu_test) if mode != "q0" else 0.0
    i_s = init.clone(); i_s[poi] = float(mu_test)
    i_b = init.clone(); i_b[poi] = bkg_mu
    four = plan.fit(data_obs, torch.stack([i_f, init, i_s, i_b]), lo, hi, fixed, fixed_alt=fp,
                    use_alt=torch.tensor([1, 0, 1, 1], dtype=torch.uint8, device=dev),
                    data_map=torch.zeros(4, dtype=torch.int32, device=dev), **fit_kw)
    obs_f, obs_u, gen = _view(four, slice(0, 1)), _view(four, slice(1, 2)), _view(four, slice(2, 4))
    muhat_obs = obs_u.x[:, poi].contiguous()
    qobs = teststat(obs_f.fun.contiguous(), obs_u.fun.contiguous(), muhat_obs, float(mu_test), mode)
    obs = dict(q=qobs, muhat=muhat_obs, fixed=obs_f, free=obs_u)
    mark("observed_fits")
    toys = torch.empty(2 * ntoys, plan.D, dtype=F64, device=dev)
    with torch.cuda.device(dev):
        _lib.check(plan._lib.hf_sample_toys(plan._h, _ptr(gen.x[0].contiguous()), int(seed), int(toy_offset),
                                            int(ntoys), _ptr(toys[:ntoys]), _stream()))
        _lib.check(plan._lib.hf_sample_toys(plan._h, _ptr(gen.x[1].contiguous()), int(seed) + 1, int(toy_offset),
                                            int(ntoys), _ptr(toys[ntoys:]), _stream()))
    mark("sampling")
    qq = qmu_tilde_batch(plan, mu_test, toys, init, lo, hi, fixed, mode, **fit_kw)
    mark("toy_fits")
    # toys whose fits did not converge carry no test statistic: the reference raises
    # FailedMinimization there (optimize/mixins.py:64-68); here they are marked NaN BEFORE the
    # gather, left out of the p-value counts and denominators, and reported (every rank sees the
    # same global count)
    bad = (qq["fixed"].status != 0) | (qq["free"].status != 0)
    q2 = torch.where(bad, torch.full_like(qq["q"], float("nan")), qq["q"]).reshape(2, ntoys)
    if gather is not None:
        q2 = gather(q2)          # one collective for both hypotheses
    mark("gather")
    keep = torch.isfinite(q2)
    n_failed = int((~keep).sum())
    if max_failed_fraction is not None and n_failed > max_failed_fraction * q2.numel():
        raise RuntimeError(f"{n_failed} of {q2.numel()} toys have a failed fit")
    q_sig, q_bkg = q2[0][keep[0]].contiguous(), q2[1][keep[1]].contiguous()
    n_sb = pvalue_counts(q_sig, qobs)[0].item()
    n_b = pvalue_counts(q_bkg, qobs)[0].item()
    clsb = n_sb / max(q_sig.numel(), 1)
    clb = n_b / max(q_bkg.numel(), 1)
    mark("pvalues")
    expected = None
    if return_expected_set:
        expected = empirical_expected_pvalues(q_sig, q_bkg)
        mark("expected_band")
    return dict(qobs=float(qobs[0]), CLsb=clsb, CLb=clb, CLs=(clsb / clb if clb > 0 else float("nan")),
                q_sig=q_sig, q_bkg=q_bkg, obs=obs, gen=gen, expected=expected, n_failed=n_failed,
                n_iter_mean=float(torch.cat([qq["fixed"].niter, qq["free"].niter]).double().mean()))


def asymptotic_pvalues(qmu, qmuA, test_stat="qtilde", calc_base_dist="normal"):
    """Closing arithmetic of AsymptoticCalculator on tensors of any shape: test statistic
    (calculators.py:416-445), observed p-values and the expected band (:447-523).  Plain torch
    elementwise code on a handful of numbers per hypothesis.  Returns ``teststat``, ``CLsb``,
    ``CLb``, ``CLs_obs`` (shape of qmu) and ``CLs_exp`` (one more trailing axis of 5:
    n_sigma = 2, 1, 0, -1, -2 as in the reference)."""
    sq, sqA = torch.sqrt(qmu), torch.sqrt(qmuA)
    if test_stat in ("q", "q0"):
        ts = sq - sqA
    else:
        safe = torch.where(sqA == 0, torch.ones_like(sqA), sqA)
        ts = torch.where((sq <= sqA) | (sqA == 0), sq - sqA, (qmu - qmuA) / (2.0 * safe))
    ndtr = torch.special.ndtr
    cutoff = -sqA if calc_base_dist == "clipped_normal" else torch.full_like(sqA, float("-inf"))
    nan = torch.full_like(sqA, float("nan"))

    def pvalues(t):
        clsb = torch.where(t >= cutoff, ndtr(-(t + sqA)), nan)     # sb_dist.shift = -sqrt(qmuA)
        clb = torch.where(t >= cutoff, ndtr(-t), nan)
        return clsb, clb, clsb / clb

    clsb, clb, cls = pvalues(ts)
    exp = []
    for nsig in (2.0, 1.0, 0.0, -1.0, -2.0):
        t = torch.full_like(sqA, nsig)
        t = torch.where(t > cutoff, t, cutoff)                      # b_dist.expected_value
        exp.append(pvalues(t)[2])
    return dict(teststat=ts, CLsb=clsb, CLb=clb, CLs_obs=cls, CLs_exp=torch.stack(exp, dim=-1))


def asymptotic_scan(plan: DevicePlan, mus, data, init=None, lo=None, hi=None, fixed=None,
                    test_stat="qtilde", calc_base_dist="normal", **fit_kw):
    """Asymptotic CLs for a whole scan of POI values in TWO batched fit launches.

    The reference runs AsymptoticCalculator.teststatistic (infer/calculators.py:345-456) once per
    scan point: 5 fits each, 205 for the 41-point scan of upper_limits.py:148-213 -- of which only
    2 * M + 3 are distinct (the free fit on the data, the Asimov-generating constrained fit and
    the free fit on the Asimov data do not depend on the scan point).  Here:

      batch 1 (M + 2 fits on the data):   free | POI fixed at asimov_mu | POI fixed at each mu
      Asimov data = expected_data(constrained fit at asimov_mu)         (calculators.py:45-97)
      batch 2 (M + 1 fits on the Asimov data):  free | POI fixed at each mu

    followed by the arithmetic of teststatistic / pvalues / expected_pvalues
    (calculators.py:416-445, :447-523) on M-element tensors.  Returns a dict of device tensors:
    ``CLs_obs`` [M], ``CLs_exp`` [M, 5] (n_sigma = 2, 1, 0, -1, -2 as in the reference), ``CLsb``,
    ``CLb``, ``qmu``, ``qmuA``, ``teststat``, ``muhat``, ``n_fits``.
    """
    if test_stat not in ("qtilde", "q", "q0"):
        raise ValueError(f"unknown test statistic {test_stat}")
    if calc_base_dist not in ("normal", "clipped_normal"):
        raise ValueError(f"unknown base distribution for asymptotics {calc_base_dist}")
    poi = plan.host.poi_index
    if poi < 0:
        raise ValueError("the model has no POI")
    dev = plan.device
    mus = _dev_tensor(mus, dev).reshape(-1)
    M = mus.numel()
    data = _dev_tensor(data, dev).reshape(1, -1)
    init = plan.init if init is None else _dev_tensor(init, dev).reshape(-1)
    fixed = plan.fixed if fixed is None else _dev_tensor(fixed, dev, torch.uint8)
    fp = fixed.clone()
    fp[poi] = 1
    asimov_mu = 1.0 if test_stat == "q0" else 0.0

    def batch(rows, with_asimov_fit):
        k = 2 if with_asimov_fit else 1
        x0 = init.expand(M + k, -1).clone()
        x0[k:, poi] = mus
        if with_asimov_fit:
            x0[1, poi] = asimov_mu
        use_alt = torch.ones(M + k, dtype=torch.uint8, device=dev)
        use_alt[0] = 0
        return plan.fit(rows, x0, lo, hi, fixed, fixed_alt=fp, use_alt=use_alt,
                        data_map=torch.zeros(M + k, dtype=torch.int32, device=dev), **fit_kw)

    r1 = batch(data, True)
    asimov = plan.expected_data(r1.x[1:2])
    r2 = batch(asimov.reshape(1, -1), False)

    def q_of(free_fun, free_mu, fixed_fun):
        q = torch.clamp(fixed_fun - free_fun, min=0.0)           # test_statistics.py:56
        if test_stat == "q0":
            return torch.where(free_mu < 0.0, torch.zeros_like(q), q)   # :507-520
        return torch.where(free_mu > mus, torch.zeros_like(q), q)       # :57-60

    qmu = q_of(r1.fun[0], r1.x[0, poi], r1.fun[2:])
    qmuA = q_of(r2.fun[0], r2.x[0, poi], r2.fun[1:])
    # at mu == asimov_mu the Asimov data ARE the constrained-fit expectation, so both fits share
    # their minimum and qmu_A = 0 identically; the two numerical minima differ by rounding
    # (~1e-12), which the square root below would turn into 1e-6 on CLs
    qmuA = torch.where(mus == asimov_mu, torch.zeros_like(qmuA), qmuA)
    pv = asymptotic_pvalues(qmu, qmuA, test_stat, calc_base_dist)
    ts, clsb, clb, cls, cls_exp = pv["teststat"], pv["CLsb"], pv["CLb"], pv["CLs_obs"], pv["CLs_exp"]
    status = torch.cat([r1.status, r2.status])
    return dict(mu=mus, CLs_obs=cls, CLs_exp=cls_exp, CLsb=clsb, CLb=clb, qmu=qmu,
                qmuA=qmuA, teststat=ts, muhat=r1.x[0, poi], asimov_data=asimov.reshape(-1),
                fits_data=r1, fits_asimov=r2, n_fits=2 * M + 3, n_failed=int((status != 0).sum()))


def upper_limit_from_scan(mus, cls, level=0.05):
    """observed / expected limits by the interpolation of upper_limits.py:16-18
    (``np.interp(level, cls[::-1], mus[::-1])`` on the scan)."""
    mus = np.asarray(mus.detach().cpu() if hasattr(mus, "detach") else mus, dtype=float)
    cls = np.asarray(cls.detach().cpu() if hasattr(cls, "detach") else cls, dtype=float)
    if cls.ndim == 1:
        return float(np.interp(level, cls[::-1], mus[::-1]))
    return [float(np.interp(level, cls[::-1, i], mus[::-1])) for i in range(cls.shape[1])]


def fp64_peak_tflops():
    lib = _lib.require()
    v = C.c_double(0.0)
    _lib.check(lib.hf_fp64_peak(C.byref(v), _stream()))
    return v.value
