"""pyhf_b200 -- B200-native (sm_100a) hot path for pyhf: batched HistFactory NLL + gradient,
batched bounded MLE fits and toy sampling behind pyhf's backend / optimizer plugin sockets.

    import pyhf, pyhf_b200
    pyhf.set_backend(pyhf_b200.b200_backend(), pyhf_b200.b200_optimizer())

Hand-written CUDA in ``libhf_b200.so`` (C-ABI: include/hf_b200.h); no CPU fallback.
"""
from .backend import B200Backend, TorchCarrierBackend, b200_backend  # noqa: F401
from .optimizer import B200Optimizer, b200_optimizer  # noqa: F401
from .plan import HostPlan, PlanError, extract_plan  # noqa: F401
from .compile import compile_spec  # noqa: F401


def __getattr__(name):
    # device-side API is imported lazily so that ``import pyhf_b200`` works on a CPU-only box
    if name in ("DevicePlan", "plan_for", "qmu_tilde_batch", "toy_hypotest", "FitResult",
                "teststat", "pvalue_counts", "fp64_peak_tflops", "asymptotic_scan",
                "upper_limit_from_scan", "empirical_expected_pvalues", "empirical_expected_value",
                "asymptotic_pvalues"):
        from . import batched

        return getattr(batched, name)
    if name in ("patchset_hypotest", "patchset_upper_limits", "apply_patch", "build_signal_group"):
        from . import patchset

        return getattr(patchset, name)
    raise AttributeError(name)
