"""``_setup()`` hook: put the fused K1 kernel under an UNCHANGED ``model.logpdf(pars, data)``.

pyhf never asks a backend for "logpdf": ``Model.logpdf`` decomposes into 30-69 generic tensor ops
(SURVEY.md section 3.1), so the op-granular backend socket cannot see a whole evaluation.
``set_backend`` calls ``new_backend._setup()`` -- "run any global setups for the lib"
(reference tensor/manager.py:183-184, tensor/numpy_backend.py:78-81).  The b200 backend uses it
to install a thin, reversible wrapper around ``pyhf.pdf.Model.logpdf`` that

* only acts while the *active* backend is the b200 backend and the model is plannable
  (unbatched, hot-path modifiers/interpolators only, no autograd request on ``pars``);
* performs the same shape checks / exceptions as the reference (pdf.py:969-981);
* otherwise defers to the original method.

The wrapper removes itself when ``tensorlib_changed`` fires for another backend.
"""
from __future__ import annotations

import logging

log = logging.getLogger(__name__)

_state = {"orig": None, "subscribed": False}


def _is_b200(tensorlib):
    from .backend import B200Backend

    return isinstance(tensorlib, B200Backend)


def install(backend):
    import pyhf
    from pyhf import exceptions

    from . import batched
    from .plan import PlanError

    if _state["orig"] is None:
        orig = pyhf.pdf.Model.logpdf

        def logpdf(self, pars, data):  # same signature as pdf.py:957
            tensorlib, _ = pyhf.get_backend()
            if not _is_b200(tensorlib) or self.batch_size is not None:
                return orig(self, pars, data)
            if getattr(pars, "requires_grad", False):
                return orig(self, pars, data)  # autograd through the generic op route
            try:
                plan = batched.plan_for(self, tensorlib.device)
            except PlanError:
                return orig(self, pars, data)
            pars, data = tensorlib.astensor(pars), tensorlib.astensor(data)
            if pars.shape[-1] != self.config.npars:
                msg = f"eval failed as pars has len {pars.shape[-1]} but {self.config.npars} was expected"
                log.error(msg)
                raise exceptions.InvalidPdfParameters(msg)
            if data.shape[-1] != self.config.nmaindata + self.config.nauxdata:
                msg = (f"eval failed as data has len {data.shape[-1]} but "
                       f"{self.config.nmaindata + self.config.nauxdata} was expected")
                log.error(msg)
                raise exceptions.InvalidPdfData(msg)
            if pars.dim() != 1 or data.dim() != 1:
                return orig(self, pars, data)
            return plan.logpdf(pars, data)  # shape (1,), pdf.py:985-988

        logpdf.__wrapped__ = orig
        logpdf.__doc__ = orig.__doc__
        pyhf.pdf.Model.logpdf = logpdf
        _state["orig"] = orig

    if not _state["subscribed"]:
        # module-level function: pyhf.events keeps plain functions by weak reference as well
        pyhf.events.subscribe("tensorlib_changed")(_on_tensorlib_changed)
        _state["subscribed"] = True


def uninstall():
    import pyhf

    if _state["orig"] is not None:
        pyhf.pdf.Model.logpdf = _state["orig"]
        _state["orig"] = None


def _on_tensorlib_changed():
    import pyhf

    tensorlib, _ = pyhf.get_backend()
    if not _is_b200(tensorlib):
        uninstall()
