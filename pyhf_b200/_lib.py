"""ctypes binding of libhf_b200.so (include/hf_b200.h).  There is no fallback: if the library
is missing or does not load, every entry point of the package raises."""
from __future__ import annotations

import ctypes as C
import os

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "libhf_b200.so")

c_dp = C.POINTER(C.c_double)
c_ip = C.POINTER(C.c_int32)


class PlanDesc(C.Structure):
    """mirror of ``hf_plan_desc``"""

    _fields_ = [
        ("n_pars", C.c_int32), ("n_bins", C.c_int32), ("n_samples", C.c_int32),
        ("n_aux", C.c_int32), ("n_segs", C.c_int32), ("n_hs", C.c_int32),
        ("n_hs_rows", C.c_int32), ("hs_code", C.c_int32), ("n_sf", C.c_int32),
        ("n_bf_slots", C.c_int32), ("n_gauss", C.c_int32), ("n_pois", C.c_int32),
        ("has_clip_sample", C.c_int32), ("has_clip_bin", C.c_int32),
        ("clip_sample", C.c_double), ("clip_bin", C.c_double),
        ("nominal", C.c_void_p), ("bin_seg", C.c_void_p), ("hs_par", C.c_void_p),
        ("hs_row_sample", C.c_void_p), ("hs_t1", C.c_void_p), ("hs_t2", C.c_void_p),
        ("sf_ptr", C.c_void_p), ("sf_par", C.c_void_p), ("sf_kind", C.c_void_p),
        ("sf_coef", C.c_void_p), ("bf_par", C.c_void_p), ("g_par", C.c_void_p),
        ("g_aux", C.c_void_p), ("g_sigma", C.c_void_p), ("p_par", C.c_void_p),
        ("p_aux", C.c_void_p), ("p_factor", C.c_void_p),
    ]


class FitOpts(C.Structure):
    """mirror of ``hf_fit_opts``"""

    _fields_ = [
        ("max_iter", C.c_int32), ("reserved0", C.c_int32), ("tol_edm", C.c_double),
        ("tol_step", C.c_double), ("verbose", C.c_int32), ("reserved", C.c_int32),
    ]


# every symbol include/hf_b200.h declares: name -> (restype, argtypes)
_vp, _i64, _u64, _i32 = C.c_void_p, C.c_int64, C.c_uint64, C.c_int32
SIGNATURES = {
    "hf_version": (C.c_int, []),
    "hf_last_error": (C.c_char_p, []),
    "hf_fp64_peak": (C.c_int, [c_dp, _vp]),
    "hf_plan_create": (C.c_int, [C.POINTER(PlanDesc), C.POINTER(_vp)]),
    "hf_plan_destroy": (None, [_vp]),
    "hf_plan_launch_count": (_i64, [_vp]),
    "hf_plan_set_timing": (C.c_int, [_vp, C.c_int]),
    "hf_plan_main_kernel_ms": (C.c_int, [_vp, c_dp, C.POINTER(C.c_int64)]),
    "hf_plan_fit_info": (C.c_int, [_vp, C.POINTER(C.c_int32)]),
    "hf_logpdf": (C.c_int, [_vp, _vp, _vp, _i64, _i64, _vp, _vp]),
    "hf_logpdf_grad": (C.c_int, [_vp, _vp, _vp, _i64, _i64, _vp, _vp, _vp]),
    "hf_expected_data": (C.c_int, [_vp, _vp, _i64, _vp, _vp]),
    "hf_logpdf_grad_host": (C.c_int, [_vp, _vp, _vp, _i64, _i64, _vp, _vp, _vp]),
    "hf_fit_default_opts": (None, [C.POINTER(FitOpts)]),
    "hf_fit_batch": (C.c_int, [_vp, _vp, _i64, _vp, _i64, _vp, _vp, _vp, _i64,
                               C.POINTER(FitOpts), _vp, _vp, _vp, _vp, _vp]),
    "hf_fit_batch_mixed": (C.c_int, [_vp, _vp, _i64, _vp, _vp, _i64, _vp, _vp, _vp, _vp, _vp, _i64,
                                     C.POINTER(FitOpts), _vp, _vp, _vp, _vp, _vp]),
    "hf_fit_batch_masks": (C.c_int, [_vp, _vp, _i64, _vp, _vp, _i64, _vp, _vp, _vp, _i32, _vp, _i64,
                                     C.POINTER(FitOpts), _vp, _vp, _vp, _vp, _vp]),
    "hf_fit_batch_host": (C.c_int, [_vp, _vp, _i64, _vp, _i64, _vp, _vp, _vp, _i64,
                                    C.POINTER(FitOpts), _vp, _vp, _vp, _vp, _vp]),
    "hf_hessian": (C.c_int, [_vp, _vp, _vp, _i64, _i64, _vp, _vp, _vp, _vp, _vp]),
    "hf_hessian_analytic": (C.c_int, [_vp, _vp, _vp, _i64, _i64, _vp, _i32, _vp, _vp]),
    "hf_sample_toys": (C.c_int, [_vp, _vp, _u64, _u64, _i64, _vp, _vp]),
    "hf_sample_poisson": (C.c_int, [_vp, _i64, _i64, _u64, _u64, _vp, _vp]),
    "hf_sample_normal": (C.c_int, [_vp, _vp, _i64, _i64, _u64, _u64, _vp, _vp]),
    "hf_teststat": (C.c_int, [_vp, _vp, _vp, C.c_double, _i32, _i64, _vp, _vp]),
    "hf_pvalue_counts": (C.c_int, [_vp, _i64, _vp, _i64, _vp, _vp]),
    "hf_pvalue_counts_sorted": (C.c_int, [_vp, _i64, _vp, _i64, _vp, _vp]),
}

_lib = None


class NativeLibraryError(RuntimeError):
    pass


def require():
    """Load (once) and return the native library; raise loudly when it is unavailable."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise NativeLibraryError(
            f"{LIB_PATH} is missing: build it with `python -m pyhf_b200.build` "
            "(or __graft_entry__.build()). The b200 path has no CPU fallback."
        )
    try:
        lib = C.CDLL(LIB_PATH)
    except OSError as exc:  # pragma: no cover - depends on the box
        raise NativeLibraryError(f"cannot load {LIB_PATH}: {exc}") from exc
    for name, (res, args) in SIGNATURES.items():
        try:
            fn = getattr(lib, name)
        except AttributeError as exc:
            raise NativeLibraryError(f"{LIB_PATH} does not export {name}") from exc
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


class HFError(RuntimeError):
    pass


def check(rc):
    if rc != 0:
        msg = require().hf_last_error()
        raise HFError(f"libhf_b200 error {rc}: {msg.decode() if msg else ''}")
