"""``b200_backend`` -- pyhf tensor-backend plugin (drop-in boundary #1).

Implements the op vocabulary of the reference's ``numpy_backend``
(``src/pyhf/tensor/numpy_backend.py:54-647``: same method names, keyword names and return
conventions) on fp64 ``torch`` tensors that live in B200 HBM.  torch is the device-memory
carrier only: the hot path (``Model.logpdf`` through the ``_setup`` hook, every fit through
``b200_optimizer``, toy sampling through ``poisson_dist/normal_dist(...).sample``) runs in the
hand-written CUDA library ``libhf_b200.so`` reached over the C-ABI of ``include/hf_b200.h``.

The generic op route stays correct for everything pyhf does outside the fused path
(``expected_data``, ``make_pdf``, the test-statistic arithmetic, ``EmpiricalDistribution``).

``TorchCarrierBackend(device="cpu")`` is the same vocabulary on CPU tensors; it exists for the
CPU test-suite and as the autograd *gradient oracle*  and is never what
``b200_backend()`` returns: the product object insists on CUDA + the native library.
"""
from __future__ import annotations

import math

import numpy as np
import torch

__all__ = ["TorchCarrierBackend", "B200Backend", "b200_backend", "DeviceArray"]


class DeviceArray(torch.Tensor):
    """torch.Tensor that behaves where pyhf treats backend tensors as numpy-like:

    * ``np.asarray(t)`` / numpy ``__setitem__`` with a device tensor on the right-hand side
      (model construction under a CUDA backend: reference modifiers/shapesys.py:168,
      staterror.py:189 assign backend index tensors into numpy arrays);
    * negative-step slices (``result_array[idx][::-1]``, infer/intervals/upper_limits.py:16-18).
    """

    @classmethod
    def __torch_function__(cls, func, types, args=(), kwargs=None):
        # the stock Tensor.__torch_function__ spends ~8 us per op in subclass checks and a generic
        # result conversion; pyhf's p-value arithmetic is ~250 tiny ops per hypotest, more time than
        # its five fits.  Same semantics (results come back as DeviceArray), leaner.
        with torch._C.DisableTorchFunctionSubclass():
            ret = func(*args, **(kwargs or {}))
        if func in _NOWRAP:  # ._base, .grad, ...: the very object, not a fresh alias (toy tags live on it)
            return ret
        if type(ret) is torch.Tensor:
            return ret.as_subclass(cls)
        if type(ret) in (tuple, list):
            return type(ret)(r.as_subclass(cls) if type(r) is torch.Tensor else r for r in ret)
        return ret

    def __array__(self, dtype=None, copy=None):
        a = self.detach().as_subclass(torch.Tensor).cpu().numpy()
        return a if dtype is None else a.astype(dtype, copy=False)

    def __getitem__(self, key):
        if isinstance(key, slice) and key.step is not None and key.step < 0:
            idx = range(*key.indices(self.shape[0]))
            index = torch.as_tensor(list(idx), dtype=torch.long, device=self.device)
            return torch.index_select(self, 0, index)
        return super().__getitem__(key)


_NOWRAP = torch.overrides.get_default_nowrap_functions()


def _wrap(t):
    return t.as_subclass(DeviceArray)


# Toy batches are TAGGED where they are created: ``poisson_dist / normal_dist(...).sample`` mark the
# storage they return, and the three ops ``Simultaneous.sample`` sends samples through on their way
# to the toy matrix (``_TensorViewer.stitch``: concatenate, einsum, gather -- reference
# tensor/common.py:40-52) pass the mark on.  ``b200_optimizer`` only batches row views of marked
# storage, so a user's own sliced data tensor never triggers a batch of unrelated fits.
def _root(t):
    b = getattr(t, "_base", None)
    return t if b is None else b


def mark_toys(t):
    if isinstance(t, torch.Tensor):
        _root(t)._hf_toys = True
    return t


def is_toys(t):
    return isinstance(t, torch.Tensor) and getattr(_root(t), "_hf_toys", False)


class _Poisson:
    """``poisson_dist`` object (numpy_backend.py:24-35)."""

    def __init__(self, backend, rate):
        self._b = backend
        self.rate = rate

    def sample(self, sample_shape):
        return self._b._sample_poisson(self.rate, tuple(sample_shape))

    def log_prob(self, value):
        return self._b.poisson_logpdf(value, self.rate)


class _Normal:
    """``normal_dist`` object (numpy_backend.py:38-51)."""

    def __init__(self, backend, loc, scale):
        self._b = backend
        self.loc = loc
        self.scale = scale

    def sample(self, sample_shape):
        return self._b._sample_normal(self.loc, self.scale, tuple(sample_shape))

    def log_prob(self, value):
        return self._b.normal_logpdf(value, self.loc, self.scale)


class TorchCarrierBackend:
    """numpy_backend's vocabulary on torch tensors of one device."""

    array_type = DeviceArray
    array_subtype = torch.Tensor

    def __init__(self, device="cpu", name="torch_carrier", **kwargs):
        self.name = name
        self.precision = kwargs.get("precision", "64b")
        if self.precision != "64b":
            raise ValueError("the b200 path computes in fp64 only (precision='64b')")
        self.device = torch.device(device)
        self.dtypemap = {"float": torch.float64, "int": torch.int64, "bool": torch.bool}
        self.default_do_grad = True
        self._generator = None

    def _setup(self):
        """Run global setups (numpy_backend.py:78-81); the product subclass installs its hook."""

    # ---------------------------------------------------------------- construction
    def astensor(self, tensor_in, dtype="float"):
        try:
            dt = self.dtypemap[dtype]
        except KeyError:
            raise KeyError("Invalid dtype: dtype must be float, int, or bool.") from None
        if isinstance(tensor_in, torch.Tensor):
            return _wrap(tensor_in.to(device=self.device, dtype=dt))
        if isinstance(tensor_in, (list, tuple)) and tensor_in and any(
            isinstance(t, torch.Tensor) for t in tensor_in
        ):
            # gotcha G2: lists of 0-d tensors (infer/calculators.py:838-869)
            return _wrap(torch.stack(
                [torch.as_tensor(t, dtype=dt, device=self.device) for t in tensor_in]
            ))
        return _wrap(torch.as_tensor(np.asarray(tensor_in), device=self.device).to(dt))

    def ones(self, shape, dtype="float"):
        return _wrap(torch.ones(shape, dtype=self.dtypemap[dtype], device=self.device))

    def zeros(self, shape, dtype="float"):
        return _wrap(torch.zeros(shape, dtype=self.dtypemap[dtype], device=self.device))

    def tolist(self, tensor_in):
        try:
            return tensor_in.tolist()
        except AttributeError:
            if isinstance(tensor_in, list):
                return tensor_in
            raise

    def to_numpy(self, tensor_in):
        if isinstance(tensor_in, torch.Tensor):
            return tensor_in.detach().as_subclass(torch.Tensor).cpu().numpy()
        return np.asarray(tensor_in)

    # ---------------------------------------------------------------- shape ops
    def shape(self, tensor):
        return tuple(tensor.shape)

    def reshape(self, tensor, newshape):
        return torch.reshape(tensor, newshape)

    def ravel(self, tensor):
        return torch.ravel(tensor)

    def tile(self, tensor_in, repeats):
        return torch.tile(tensor_in, tuple(repeats) if not isinstance(repeats, int) else (repeats,))

    def concatenate(self, sequence, axis=0):
        sequence = list(sequence)
        out = torch.cat(sequence, dim=axis)
        return mark_toys(out) if any(is_toys(t) for t in sequence) else out

    def stack(self, sequence, axis=0):
        return torch.stack(list(sequence), dim=axis)

    def transpose(self, tensor_in):
        return tensor_in.permute(*reversed(range(tensor_in.dim())))

    def simple_broadcast(self, *args):
        args = [a if a.dim() else a.reshape(1) for a in map(self.astensor, args)]
        return list(torch.broadcast_tensors(*args))

    def gather(self, tensor, indices):
        out = tensor[indices.long()]
        return mark_toys(out) if is_toys(tensor) else out

    def boolean_mask(self, tensor, mask):
        return tensor[mask]

    # ---------------------------------------------------------------- arithmetic
    def sum(self, tensor_in, axis=None):
        return torch.sum(tensor_in) if axis is None else torch.sum(tensor_in, dim=axis)

    def product(self, tensor_in, axis=None):
        return torch.prod(tensor_in) if axis is None else torch.prod(tensor_in, dim=axis)

    def abs(self, tensor):
        return torch.abs(tensor)

    def power(self, tensor_in_1, tensor_in_2):
        return torch.pow(tensor_in_1, tensor_in_2)

    def sqrt(self, tensor_in):
        return torch.sqrt(tensor_in)

    def divide(self, tensor_in_1, tensor_in_2):
        return torch.div(tensor_in_1, tensor_in_2)

    def log(self, tensor_in):
        return torch.log(tensor_in)

    def exp(self, tensor_in):
        return torch.exp(tensor_in)

    def outer(self, tensor_in_1, tensor_in_2):
        return torch.outer(tensor_in_1, tensor_in_2)

    def erf(self, tensor_in):
        return torch.erf(tensor_in)

    def erfinv(self, tensor_in):
        return torch.erfinv(tensor_in)

    def isfinite(self, tensor):
        return torch.isfinite(tensor)

    def clip(self, tensor_in, min_value, max_value):
        return torch.clamp(tensor_in, min_value, max_value)

    def where(self, mask, tensor_in_1, tensor_in_2):
        return torch.where(mask.bool() if isinstance(mask, torch.Tensor) else mask,
                           tensor_in_1, tensor_in_2)

    def conditional(self, predicate, true_callable, false_callable):
        return true_callable() if predicate else false_callable()

    def einsum(self, subscripts, *operands):
        ops = [o if isinstance(o, torch.Tensor) else self.astensor(o) for o in operands]
        dt = torch.float64 if any(o.is_floating_point() for o in ops) else ops[0].dtype
        out = torch.einsum(subscripts, *[o.to(dt) for o in ops])
        return mark_toys(out) if len(ops) == 1 and is_toys(ops[0]) else out

    def percentile(self, tensor_in, q, axis=None, method="linear"):
        # calculators.py:697-699, :970-972 (numpy semantics: q in percent)
        qq = torch.as_tensor(q, dtype=torch.float64, device=tensor_in.device) / 100.0
        return torch.quantile(tensor_in, qq, dim=axis, interpolation=method)

    # ---------------------------------------------------------------- probability
    def poisson_logpdf(self, n, lam):
        n = self.astensor(n)
        lam = self.astensor(lam)
        return torch.xlogy(n, lam) - lam - torch.lgamma(n + 1.0)  # numpy_backend.py:435-436

    def poisson(self, n, lam):
        return torch.exp(self.poisson_logpdf(n, lam))

    def normal_logpdf(self, x, mu, sigma):
        # numpy_backend.py:482-490
        root2 = math.sqrt(2)
        root2pi = math.sqrt(2 * math.pi)
        prefactor = -torch.log(sigma * root2pi)
        summand = -torch.square(torch.div((x - mu), (root2 * sigma)))
        return prefactor + summand

    def normal(self, x, mu, sigma):
        x, mu, sigma = map(self.astensor, (x, mu, sigma))
        return torch.exp(self.normal_logpdf(x, mu, sigma))

    def normal_cdf(self, x, mu=0.0, sigma=1):
        x = self.astensor(x)
        return 0.5 * (1.0 + torch.erf((x - mu) / (sigma * math.sqrt(2))))

    def poisson_dist(self, rate):
        return _Poisson(self, rate)

    def normal_dist(self, mu, sigma):
        return _Normal(self, mu, sigma)

    # ---------------------------------------------------------------- sampling (carrier: torch RNG)
    def _sample_poisson(self, rate, sample_shape):
        r = rate.as_subclass(torch.Tensor).expand(sample_shape + tuple(rate.shape))
        return mark_toys(_wrap(torch.poisson(r, generator=self._generator)))

    def _sample_normal(self, loc, scale, sample_shape):
        loc, scale = torch.broadcast_tensors(self.astensor(loc), self.astensor(scale))
        shape = sample_shape + tuple(loc.shape)
        z = torch.randn(shape, dtype=torch.float64, device=loc.device, generator=self._generator)
        return mark_toys(_wrap(loc + scale * z))


class B200Backend(TorchCarrierBackend):
    """The product backend: CUDA only, native library mandatory (no CPU fallback)."""

    def __init__(self, device=None, seed=0, toy_rng="philox", **kwargs):
        """``toy_rng``: "philox" (default) draws toys with the on-device Philox sampler (K4);
        "numpy" is the exact-parity mode -- toys are drawn on the host exactly as the reference's
        numpy backend draws them (``scipy.stats.poisson / norm(...).rvs`` on numpy's global RNG,
        tensor/numpy_backend.py:24-51) and uploaded, so that an UNCHANGED
        ``hypotest(..., calctype="toybased")`` under ``np.random.seed(s)`` fits bit-identical toys to
        the reference's and returns its CLs.  Only where the toys come from changes; every fit
        still runs in the CUDA minimiser."""
        from . import _lib

        _lib.require()  # raises loudly when libhf_b200.so is missing / not loadable
        if not torch.cuda.is_available():
            raise RuntimeError(
                "b200_backend needs a CUDA device; there is no CPU fallback on this path"
            )
        if device is None:
            device = torch.device("cuda", torch.cuda.current_device())
        super().__init__(device=device, name="b200", **kwargs)
        self.seed = int(seed)
        self._toy_offset = 0
        if toy_rng not in ("philox", "numpy"):
            raise ValueError("toy_rng must be 'philox' or 'numpy'")
        self.toy_rng = toy_rng

    def _setup(self):
        from . import hooks

        hooks.install(self)

    # toys come from the Philox sampler kernels (K4), not from torch's RNG
    def _sample_poisson(self, rate, sample_shape):
        if self.toy_rng == "numpy":
            from scipy.stats import poisson

            r = self.to_numpy(rate)
            return mark_toys(self.astensor(poisson(r).rvs(size=tuple(sample_shape) + r.shape)))
        from . import batched

        return mark_toys(_wrap(batched.sample_poisson(self, rate, sample_shape)))

    def _sample_normal(self, loc, scale, sample_shape):
        if self.toy_rng == "numpy":
            from scipy.stats import norm

            lo_, sc_ = np.broadcast_arrays(self.to_numpy(self.astensor(loc)), self.to_numpy(self.astensor(scale)))
            return mark_toys(self.astensor(norm(lo_, sc_).rvs(size=tuple(sample_shape) + lo_.shape)))
        from . import batched

        return mark_toys(_wrap(batched.sample_normal(self, loc, scale, sample_shape)))


def b200_backend(**kwargs):
    """``pyhf.set_backend(b200_backend(), b200_optimizer())``"""
    return B200Backend(**kwargs)
