"""Likelihood / derivative reductions with the reference's Cython signatures
(reference: gp/ext/gp_c.pyx:17,34,52,70,114): numpy arrays in, outputs written in place.

Each function validates its buffers the way the Cython declarations do and makes ONE call into
libgpb200.so (``gpb_gp_c_*``, include/gpb200.h): the library stages the host arrays into its own
device pool and runs the CUDA kernels -- there is no torch operation on this path.  ``GP`` itself
never goes through here: it keeps everything on the device and regenerates kernel tiles instead of
reading ``Kj`` / ``Kh``.
"""
import ctypes

import numpy as np

from .._lib import call
from ._host import carray, out_array

__all__ = ["log_lh", "dloglh_dtheta", "dlh_dtheta", "d2lh_dtheta2", "dm_dtheta"]

MIN = float(np.log(np.exp2(np.float64(np.finfo(np.float64).minexp + 4))))


def _p(a):
    return a.ctypes.data


def _same(name, shape, expected):
    if tuple(shape) != tuple(expected):
        raise ValueError("%s has shape %s, expected %s" % (name, tuple(shape), tuple(expected)))


def log_lh(y, K, Kiy):
    """gp_c.log_lh (gp_c.pyx:17-31).  log|K| comes from a Cholesky of K instead of the
    reference's LU ``slogdet``; a K that is not positive definite returns ``-inf`` (for
    such K the reference returns -inf too whenever det K <= 0)."""
    carray(y, 1, "y"); carray(K, 2, "K"); carray(Kiy, 1, "Kiy")
    n = y.size
    _same("K", K.shape, (n, n)); _same("Kiy", Kiy.shape, (n,))
    out = ctypes.c_double()
    call("gpb_gp_c_log_lh", _p(y), _p(K), _p(Kiy), n, ctypes.byref(out))
    return float(out.value)


def _check_common(y, Ki, Kj, Kiy):
    carray(y, 1, "y"); carray(Ki, 2, "Ki"); carray(Kj, 3, "Kj"); carray(Kiy, 1, "Kiy")
    n_p, n = Kj.shape[0], y.size
    _same("Ki", Ki.shape, (n, n)); _same("Kj", Kj.shape, (n_p, n, n)); _same("Kiy", Kiy.shape, (n,))
    return n_p, n


def dloglh_dtheta(y, Ki, Kj, Kiy, s, dloglh):
    """gp_c.dloglh_dtheta (gp_c.pyx:34-49)."""
    n_p, n = _check_common(y, Ki, Kj, Kiy)
    out_array(dloglh, (n_p + 1,), "dloglh")
    call("gpb_gp_c_dloglh_dtheta", _p(y), _p(Ki), _p(Kj), _p(Kiy), float(s), n_p, n, _p(dloglh))


def dlh_dtheta(y, Ki, Kj, Kiy, s, lh, dlh):
    """gp_c.dlh_dtheta (gp_c.pyx:52-67)."""
    n_p, n = _check_common(y, Ki, Kj, Kiy)
    out_array(dlh, (n_p + 1,), "dlh")
    call("gpb_gp_c_dlh_dtheta", _p(y), _p(Ki), _p(Kj), _p(Kiy), float(s), float(lh), n_p, n, _p(dlh))


def d2lh_dtheta2(y, Ki, Kj, Kh, Kiy, s, lh, dlh, d2lh):
    """gp_c.d2lh_dtheta2 (gp_c.pyx:70-111): the n_p products Ki dK_i on the DMMA GEMM, everything
    else as O(N^2) traces / quadratic forms over the caller's dense Ki, Kj, Kh."""
    n_p, n = _check_common(y, Ki, Kj, Kiy)
    carray(Kh, 4, "Kh"); carray(dlh, 1, "dlh")
    _same("Kh", Kh.shape, (n_p, n_p, n, n)); _same("dlh", dlh.shape, (n_p + 1,))
    out_array(d2lh, (n_p + 1, n_p + 1), "d2lh")
    call("gpb_gp_c_d2lh_dtheta2", _p(y), _p(Ki), _p(Kj), _p(Kh), _p(Kiy), float(s), float(lh), _p(dlh), n_p, n, _p(d2lh))


def dm_dtheta(y, Ki, Kj, Kjxo, Kxox, s, dm):
    """gp_c.dm_dtheta (gp_c.pyx:114-131) with mat-vecs:
    dm[i] = dKxox_i (Ki y) - Kxox (Ki (dK_i (Ki y)))."""
    carray(y, 1, "y"); carray(Ki, 2, "Ki"); carray(Kj, 3, "Kj"); carray(Kjxo, 3, "Kjxo"); carray(Kxox, 2, "Kxox")
    n_p, n, m = Kj.shape[0], y.size, Kxox.shape[0]
    _same("Ki", Ki.shape, (n, n)); _same("Kj", Kj.shape, (n_p, n, n))
    _same("Kjxo", Kjxo.shape, (n_p, m, n)); _same("Kxox", Kxox.shape, (m, n))
    out_array(dm, (n_p + 1, m), "dm")
    call("gpb_gp_c_dm_dtheta", _p(y), _p(Ki), _p(Kj), _p(Kjxo), _p(Kxox), float(s), n_p, n, m, _p(dm))
