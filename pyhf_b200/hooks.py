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

``Model.expected_data`` (pdf.py:886-906) gets the same treatment (``hf_expected_data``): it is what
``generate_asimov_data`` calls between the fits of an asymptotic hypotest.

The wrappers remove themselves when ``tensorlib_changed`` fires for another backend.
"""
from __future__ import annotations

import logging

log = logging.getLogger(__name__)

_state = {"orig": None, "orig_expected": None, "subscribed": False}


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

        orig_expected = pyhf.pdf.Model.expected_data

        def expected_data(self, pars, include_auxdata=True):  # same signature as pdf.py:886
            # Asimov data (infer/calculators.py:45-97) and the toy-generation points go through here:
            # one hf_expected_data launch instead of ~30 eager tensor ops
            tensorlib, _ = pyhf.get_backend()
            if not _is_b200(tensorlib) or self.batch_size is not None or getattr(pars, "requires_grad", False):
                return orig_expected(self, pars, include_auxdata)
            try:
                plan = batched.plan_for(self, tensorlib.device)
            except PlanError:
                return orig_expected(self, pars, include_auxdata)
            pars = tensorlib.astensor(pars)
            if pars.shape[-1] != self.config.npars:  # pdf.py:899-901
                msg = (f"Evaluation failed as parameters have length {pars.shape[-1]} but model requires "
                       f"{self.config.npars}.")
                raise exceptions.InvalidPdfParameters(msg)
            if pars.dim() != 1:
                return orig_expected(self, pars, include_auxdata)
            out = tensorlib.astensor(plan.expected_data(pars))
            return out if include_auxdata else out[: self.config.nmaindata]

        expected_data.__wrapped__ = orig_expected
        expected_data.__doc__ = orig_expected.__doc__
        pyhf.pdf.Model.expected_data = expected_data
        _state["orig_expected"] = orig_expected

    if not _state["subscribed"]:
        # module-level function: pyhf.events keeps plain functions by weak reference as well
        pyhf.events.subscribe("tensorlib_changed")(_on_tensorlib_changed)
        _state["subscribed"] = True


def uninstall():
    import pyhf

    if _state["orig"] is not None:
        pyhf.pdf.Model.logpdf = _state["orig"]
        _state["orig"] = None
    if _state["orig_expected"] is not None:
        pyhf.pdf.Model.expected_data = _state["orig_expected"]
        _state["orig_expected"] = None


def _on_tensorlib_changed():
    import pyhf

    tensorlib, _ = pyhf.get_backend()
    if not _is_b200(tensorlib):
        uninstall()
