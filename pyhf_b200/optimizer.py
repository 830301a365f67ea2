"""``b200_optimizer`` -- pyhf optimizer plugin (drop-in boundary #2).

Mirrors the surface of the reference's ``OptimizerMixin.minimize``
(``src/pyhf/optimize/mixins.py:124-209``: same arguments, same return conventions, same
exceptions) but cannot inherit it: ``shim()`` refuses any backend not named numpy/jax
(``optimize/common.py:32-50``).  Instead of wrapping the Python objective it recognises the only
objective pyhf ever passes on this path -- ``pyhf.infer.mle.twice_nll`` on a ``pyhf.Model``
(``infer/mle.py:134-136``) -- compiles the model into a device plan once, and runs the batched
CUDA minimiser (K3).  Anything else raises ``pyhf.exceptions.Unsupported``: there is no CPU
fallback on this path.

Toy batching under an UNCHANGED ``ToyCalculator``: the calculator loops
``for sample in signal_sample: teststat_func(..., sample, ...)`` (``infer/calculators.py:838-866``).
``sample`` is a row view of the toy batch the backend created, so the first ``minimize`` call of
a loop fits *all* rows of that batch in one K3 launch and the following calls are answered from
the cache.
"""
from __future__ import annotations

import weakref

import numpy as np
import torch

from . import batched
from .plan import PlanError

__all__ = ["B200Optimizer", "b200_optimizer"]


class _Result(dict):
    """scipy.optimize.OptimizeResult-like record (attribute access)."""

    __getattr__ = dict.get
    __setattr__ = dict.__setitem__


class B200Optimizer:
    """Batched on-device minimiser behind pyhf's optimizer socket."""

    __slots__ = ["name", "maxiter", "verbose", "tolerance", "_cache", "_stats", "_consts", "__weakref__"]

    def __init__(self, *args, **kwargs):
        # reference: optimize/mixins.py:19-32 (unknown kwargs -> Unsupported)
        self.name = "b200"
        self.maxiter = kwargs.pop("maxiter", 200)
        self.verbose = kwargs.pop("verbose", 0)
        self.tolerance = kwargs.pop("tolerance", None)
        if kwargs:
            import pyhf

            raise pyhf.exceptions.Unsupported(f"Unsupported kwargs were passed in: {list(kwargs)}.")
        self._cache = {}
        self._consts = {}
        self._stats = {"batched_fits": 0, "cache_hits": 0, "single_fits": 0}

    # ------------------------------------------------------------------ helpers
    @staticmethod
    def _toy_batch_of(data, D):
        """If ``data`` is a row view of an [N, D] toy batch return (key, matrix, row, base) else None."""
        if not isinstance(data, torch.Tensor) or data.dim() != 1 or data.shape[0] != D:
            return None
        base = data._base
        if base is None or base.numel() == data.numel() or base.numel() % D != 0:
            return None
        if not getattr(base, "_hf_toys", False):
            # only storage the backend's samplers created is a toy batch (backend.mark_toys): a row
            # of a user's own data matrix is fitted on its own
            return None
        N = base.numel() // D
        stride = data.stride(0)
        off = data.storage_offset() - base.storage_offset()
        if stride == 1 and off % D == 0:        # base is [N, D] row-major
            row, strides = off // D, (D, 1)
        elif stride == N and 0 <= off < N:      # base is [D, N] (stitched / transposed batch)
            row, strides = off, (1, N)
        else:
            return None
        key = (base.data_ptr(), base._version, N, D)
        matrix = torch.as_strided(base, (N, D), strides, base.storage_offset())
        return key, matrix, int(row), base

    def _fit(self, plan, data, init, lo, hi, fixed, opts):
        # ``tolerance``: the reference hands it to scipy as ``tol`` (optimize/opt_scipy.py:96-105), which
        # SLSQP reads as ftol -- stop when 2NLL changes by less than that between two iterates.  The
        # Newton iteration here stops on the same quantity, predicted instead of observed: the
        # estimated distance of 2NLL to the minimum (``tol_edm``, and a step below 1e-8).  So an
        # explicit tolerance keeps its meaning; only the DEFAULT differs on purpose (1e-11 instead
        # of scipy's 1e-6, which leaves the reference 1e-5 .. 0.5 above the minimum in 2NLL and
        # cannot meet the 1e-6 bar on q~ / CLs).
        tol = opts.get("tolerance", self.tolerance)
        kw = {"max_iter": opts.get("maxiter", self.maxiter), "verbose": opts.get("verbose", self.verbose)}
        if tol is not None:
            kw["tol_edm"] = float(tol)
        # bounds and masks repeat from call to call (five fits per hypotest, one call per toy): their
        # device copies are kept, so a fit uploads its start point and data only
        lo, hi, fixed = (self._on_device(plan, a) for a in (lo, hi, fixed))
        return plan.fit(data, init, lo, hi, fixed, **kw)

    def _on_device(self, plan, arr):
        key = (id(plan), arr.dtype.str, arr.tobytes())
        t = self._consts.get(key)
        if t is None:
            if len(self._consts) > 64:
                self._consts.clear()
            t = self._consts[key] = (torch.as_tensor(np.ascontiguousarray(arr), device=plan.device), plan)
        return t[0]

    # ------------------------------------------------------------------ the plugin entry point
    def minimize(
        self,
        objective,
        data,
        pdf,
        init_pars,
        par_bounds,
        fixed_vals=None,
        return_fitted_val=False,
        return_result_obj=False,
        return_uncertainties=False,
        return_correlations=False,
        do_grad=None,
        do_stitch=False,
        **kwargs,
    ):
        import pyhf

        opts = dict(kwargs)
        known = {"maxiter", "verbose", "tolerance"}
        unknown = [k for k in opts if k not in known]
        if unknown:  # optimize/opt_scipy.py:81-83
            raise pyhf.exceptions.Unsupported(f"Unsupported options were passed in: {unknown}.")
        if objective is not pyhf.infer.mle.twice_nll or not isinstance(pdf, pyhf.pdf.Model):
            raise pyhf.exceptions.Unsupported(
                "b200_optimizer minimises pyhf.infer.mle.twice_nll of a pyhf.Model on the GPU; "
                "arbitrary Python objectives have no CPU fallback on this path"
            )
        tensorlib, _ = pyhf.get_backend()
        try:
            plan = batched.plan_for(pdf)
        except PlanError as exc:
            raise pyhf.exceptions.Unsupported(str(exc)) from exc
        dev = plan.device
        P, D = plan.P, plan.D

        init = np.asarray(tensorlib.tolist(init_pars) if isinstance(init_pars, torch.Tensor) else init_pars,
                          dtype=np.float64)
        bounds = np.asarray(par_bounds, dtype=np.float64)
        fixed = np.zeros(P, dtype=np.uint8)
        for idx, val in fixed_vals or []:  # infer/mle.py:128-132; opt_scipy.py:85-94
            fixed[idx] = 1
            init[idx] = float(val)
        # (the plan object is kept alive by the cache entry, so its id cannot be recycled)
        sig = (id(plan), init.tobytes(), bounds.tobytes(), fixed.tobytes(),
               tuple(sorted((k, str(v)) for k, v in opts.items())))

        toy = self._toy_batch_of(data, D)
        if toy is not None:
            key, matrix, row, base = toy
            ckey = (key, sig)
            hit = self._cache.get(ckey)
            if hit is None:
                res = self._fit(plan, matrix.contiguous(), init, bounds[:, 0], bounds[:, 1], fixed, opts)
                # the entry keeps ``base`` alive so its address cannot be recycled while cached
                hit = (res.x, res.fun, res.status.cpu().numpy(), res.niter.cpu().numpy(), base, plan)
                if len(self._cache) > 8:
                    self._cache.pop(next(iter(self._cache)))
                self._cache[ckey] = hit
                self._stats["batched_fits"] += matrix.shape[0]
            else:
                self._stats["cache_hits"] += 1
            x, fun, status, nit = hit[0][row], hit[1][row], int(hit[2][row]), int(hit[3][row])
        else:
            d = data if isinstance(data, torch.Tensor) else np.asarray(data, dtype=np.float64)
            res = self._fit(plan, d, init, bounds[:, 0], bounds[:, 1], fixed, opts)
            self._stats["single_fits"] += 1
            x, fun = res.x[0], res.fun[0]
            status, nit = (int(v) for v in torch.stack([res.status[0], res.niter[0]]).tolist())  # one read-back

        result = _Result(x=x, fun=fun, success=(status == 0), status=status, nit=nit,
                         message=("converged", "iteration limit reached", "numerical failure")[status],
                         unc=None, corr=None)
        if not result.success:  # optimize/mixins.py:64-68
            raise pyhf.exceptions.FailedMinimization(result)
        if return_uncertainties or return_correlations:
            # minuit-only in the reference (optimize/opt_minuit.py:136-152); here from the Hessian of
            # 2*NLL at the minimum (hf_hessian).  Same output layout as optimize/mixins.py:83-120:
            # x becomes [P, 2] = (value, uncertainty), fixed parameters get uncertainty 0
            d_row = matrix[row] if toy is not None else d
            _, unc, corr = plan.covariance(x, d_row, bounds[:, 0], bounds[:, 1], fixed)
            result.unc, result.corr = unc, corr
            if return_uncertainties:
                result.x = torch.stack([result.x, unc], dim=1)
        if do_stitch:
            # the reference strips fixed parameters before and stitches them back after the fit
            # (optimize/common.py:108-125); the full vector is fitted here, nothing to stitch
            pass
        _returns = [result.x]
        if return_correlations:
            _returns.append(result.corr)
        if return_fitted_val:
            _returns.append(result.fun)
        if return_result_obj:
            _returns.append(result)
        return tuple(_returns) if len(_returns) > 1 else _returns[0]


def b200_optimizer(**kwargs):
    """``pyhf.set_backend(b200_backend(), b200_optimizer())``"""
    return B200Optimizer(**kwargs)
