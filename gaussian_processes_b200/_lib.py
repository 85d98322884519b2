"""ctypes binding of libgpb200.so (the C ABI declared in include/gpb200.h).

The product path has no CPU fallback: if the CUDA library is missing this module
raises ImportError with the build command, and every entry point raises
``GpbError`` on a non-zero status (e.g. when no GPU is present).
"""
import ctypes
import os
from ctypes import POINTER, c_char_p, c_double, c_int, c_int64, c_size_t, c_uint, c_void_p

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "libgpb200.so")


class GpbError(RuntimeError):
    pass


def _load():
    if not os.path.exists(LIB_PATH):
        raise ImportError(
            "gaussian_processes_b200: %s not found. Build it with "
            "`python -c 'import __graft_entry__ as g; g.build()'` or "
            "`make -C gaussian_processes_b200/csrc` (needs nvcc; there is no CPU fallback)." % LIB_PATH)
    return ctypes.CDLL(LIB_PATH)


lib = _load()

dp = POINTER(c_double)
ip = POINTER(c_int)
vp = c_void_p          # device pointers and streams travel as integers

_i64 = c_int64
_SIG_G = [vp, vp, _i64, vp, _i64, c_double, c_double]                  # host pointers as void*
_SIG_P = [vp, vp, _i64, vp, _i64, c_double, c_double, c_double]

_PROTOS = {
    "gpb_version": (c_int, []),
    "gpb_last_error": (c_char_p, []),
    "gpb_min_log": (c_double, []),
    "gpb_launch_count": (_i64, []),
    "gpb_kernel_build": (c_int, [c_int, dp, c_double, vp, _i64, vp, _i64, _i64, _i64, c_uint, vp, _i64,
                                 _i64, c_int, c_int, vp]),
    "gpb_kernel_matvec": (c_int, [c_int, dp, vp, _i64, vp, _i64, c_int, ip, ip, dp, POINTER(vp), c_int,
                                  POINTER(vp), vp]),
    "gpb_post_var": (c_int, [c_int, dp, vp, _i64, _i64, _i64, vp, vp]),
    "gpb_potrf": (c_int, [vp, _i64, _i64, _i64, c_int, vp, _i64, _i64, vp, _i64, _i64, vp, vp]),
    "gpb_potrs": (c_int, [vp, vp, _i64, _i64, _i64, _i64, _i64, c_int, vp, _i64, vp, vp, _i64, vp, vp]),
    "gpb_trtri": (c_int, [vp, _i64, _i64, _i64, c_int, vp, _i64, _i64, vp, _i64, _i64, vp, _i64, _i64, vp]),
    "gpb_lauum": (c_int, [vp, _i64, _i64, _i64, c_int, vp, _i64, _i64, vp]),
    "gpb_tril": (c_int, [vp, _i64, _i64, vp]),
    "gpb_tril_copy": (c_int, [vp, _i64, vp, _i64, _i64, vp]),
    "gpb_copy2d": (c_int, [vp, _i64, vp, _i64, _i64, _i64, vp]),
    "gpb_download_2d": (c_int, [vp, _i64, vp, _i64, _i64, _i64, c_int, vp]),
    "gpb_gemm_nt": (c_int, [vp, _i64, vp, _i64, vp, _i64, vp, _i64, _i64, _i64, _i64, c_double, c_double,
                            c_int, c_int, c_int, vp]),
    "gpb_loglh": (c_int, [vp, _i64, _i64, vp, vp, vp, vp, vp]),
    "gpb_slice_reduce": (c_int, [c_int, dp, vp, _i64, vp, _i64, vp, c_int, ip, vp, vp, vp]),
    "gpb_grad_partial_doubles": (_i64, [_i64]),
    "gpb_gemv": (c_int, [vp, _i64, _i64, _i64, vp, vp, c_double, c_double, vp]),
    "gpb_trace_prod": (c_int, [vp, _i64, vp, _i64, _i64, vp, vp, vp]),
    "gpb_quadform": (c_int, [vp, vp, _i64, vp, _i64, vp, vp, vp]),
    "gpb_eval_workspace_bytes": (c_size_t, [_i64, c_int, c_int]),
    "gpb_gp_eval": (c_int, [c_int, dp, c_int, vp, vp, _i64, c_int, vp, c_size_t, vp, vp]),
    "gpb_gp_eval_host": (c_int, [c_int, dp, c_int, dp, dp, _i64, c_int, dp]),
    "gpb_eval_layout": (c_int, [_i64, POINTER(c_int64), c_int]),
    "gpb_gp_stages": (c_int, [c_int, dp, vp, vp, _i64, c_uint, vp, c_size_t, vp, vp]),
    "gpb_post_mean_host": (c_int, [c_int, dp, vp, _i64, vp, _i64, vp, vp, vp, vp]),
    "gpb_post_cov_scratch_doubles": (c_size_t, [_i64, _i64]),
    "gpb_post_cov_host": (c_int, [c_int, dp, vp, _i64, vp, _i64, vp, _i64, vp, vp, _i64, vp]),
    "gpb_kernel_slices_host": (c_int, [c_int, c_uint, vp, vp, _i64, vp, _i64, dp]),
    "gpb_gp_c_log_lh": (c_int, [vp, vp, vp, _i64, dp]),
    "gpb_gp_c_dloglh_dtheta": (c_int, [vp, vp, vp, vp, c_double, _i64, _i64, vp]),
    "gpb_gp_c_dlh_dtheta": (c_int, [vp, vp, vp, vp, c_double, c_double, _i64, _i64, vp]),
    "gpb_gp_c_d2lh_dtheta2": (c_int, [vp, vp, vp, vp, vp, c_double, c_double, vp, _i64, _i64, vp]),
    "gpb_gp_c_dm_dtheta": (c_int, [vp, vp, vp, vp, vp, c_double, _i64, _i64, _i64, vp]),
    "gpb_microbench_fp64": (c_int, [c_int, c_int, dp, dp]),
    "gpb_microbench_latency": (c_int, [dp]),
    "gpb_set_option": (c_int, [c_char_p, c_int]),
    "gpb_profile_enable": (None, [c_int]),
    "gpb_profile_read": (c_int, [c_int, dp, POINTER(c_int64)]),
}
for _n in ("K", "jacobian", "hessian", "dK_dh", "dK_dw", "d2K_dhdh", "d2K_dhdw", "d2K_dwdh", "d2K_dwdw"):
    _PROTOS["gpb_gaussian_" + _n] = (c_int, _SIG_G)
for _n in ("K", "jacobian", "hessian", "dK_dh", "dK_dw", "dK_dp", "d2K_dhdh", "d2K_dhdw", "d2K_dhdp",
           "d2K_dwdh", "d2K_dwdw", "d2K_dwdp", "d2K_dpdh", "d2K_dpdw", "d2K_dpdp"):
    _PROTOS["gpb_periodic_" + _n] = (c_int, _SIG_P)

#: every symbol include/gpb200.h declares (tests check the library exports them all)
EXPORTS = sorted(_PROTOS)

for _name, (_res, _args) in _PROTOS.items():
    _f = getattr(lib, _name)
    _f.restype = _res
    _f.argtypes = _args


def last_error():
    msg = lib.gpb_last_error()
    return msg.decode("utf-8", "replace") if msg else ""


def check(status, what=""):
    """Raise on a non-zero C-ABI status."""
    if status != 0:
        raise GpbError("%s failed (status %d): %s" % (what or "libgpb200", status, last_error()))


def call(name, *args):
    check(getattr(lib, name)(*args), name)


def set_option(name, value):
    """Tuning knobs: eval_streams, gemm_bm, potrf_inner (0 = default)."""
    call("gpb_set_option", name.encode(), int(value))


def darr(values):
    """ctypes double array from a python sequence (host-side parameter vectors)."""
    vals = [float(v) for v in values]
    return (c_double * len(vals))(*vals)


def iarr(values):
    vals = [int(v) for v in values]
    return (c_int * len(vals))(*vals)


def parr(values):
    vals = [int(v) for v in values]
    return (vp * len(vals))(*vals)
