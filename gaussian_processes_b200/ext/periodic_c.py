"""Periodic kernel builders with the reference's Cython signatures
(reference: gp/ext/periodic_c.pyx): ``f(out, x1, x2, h, w, p) -> None`` with a
caller-allocated C-contiguous float64 ``out``.  Each call binds the C-ABI entry point
``gpb_periodic_<name>`` of libgpb200.so (host buffers in, host buffers out; the CUDA
builder computes every requested slice from one exp + one sincos per element)."""
import ctypes

from .. import _lib
from ._host import carray, out_array

__all__ = ['K', 'jacobian', 'hessian', 'dK_dh', 'dK_dw', 'dK_dp', 'd2K_dhdh', 'd2K_dhdw', 'd2K_dhdp', 'd2K_dwdh', 'd2K_dwdw', 'd2K_dwdp', 'd2K_dpdh', 'd2K_dpdw', 'd2K_dpdp']


def _call(name, lead, out, x1, x2, h, w, p):
    carray(x1, 1, "x1")
    carray(x2, 1, "x2")
    out_array(out, lead + (x1.size, x2.size))
    _lib.call("gpb_periodic_" + name, out.ctypes.data, x1.ctypes.data, x1.size, x2.ctypes.data, x2.size,
              float(h), float(w), float(p))


def K(out, x1, x2, h, w, p):
    """periodic_c.pyx:18 -> gpb_periodic_K"""
    _call("K", (), out, x1, x2, h, w, p)

def jacobian(out, x1, x2, h, w, p):
    """periodic_c.pyx:33 -> gpb_periodic_jacobian"""
    _call("jacobian", (3,), out, x1, x2, h, w, p)

def hessian(out, x1, x2, h, w, p):
    """periodic_c.pyx:39 -> gpb_periodic_hessian"""
    _call("hessian", (3, 3), out, x1, x2, h, w, p)

def dK_dh(out, x1, x2, h, w, p):
    """periodic_c.pyx:53 -> gpb_periodic_dK_dh"""
    _call("dK_dh", (), out, x1, x2, h, w, p)

def dK_dw(out, x1, x2, h, w, p):
    """periodic_c.pyx:68 -> gpb_periodic_dK_dw"""
    _call("dK_dw", (), out, x1, x2, h, w, p)

def dK_dp(out, x1, x2, h, w, p):
    """periodic_c.pyx:83 -> gpb_periodic_dK_dp"""
    _call("dK_dp", (), out, x1, x2, h, w, p)

def d2K_dhdh(out, x1, x2, h, w, p):
    """periodic_c.pyx:99 -> gpb_periodic_d2K_dhdh"""
    _call("d2K_dhdh", (), out, x1, x2, h, w, p)

def d2K_dhdw(out, x1, x2, h, w, p):
    """periodic_c.pyx:114 -> gpb_periodic_d2K_dhdw"""
    _call("d2K_dhdw", (), out, x1, x2, h, w, p)

def d2K_dhdp(out, x1, x2, h, w, p):
    """periodic_c.pyx:129 -> gpb_periodic_d2K_dhdp"""
    _call("d2K_dhdp", (), out, x1, x2, h, w, p)

def d2K_dwdh(out, x1, x2, h, w, p):
    """periodic_c.pyx:145 -> gpb_periodic_d2K_dwdh"""
    _call("d2K_dwdh", (), out, x1, x2, h, w, p)

def d2K_dwdw(out, x1, x2, h, w, p):
    """periodic_c.pyx:160 -> gpb_periodic_d2K_dwdw"""
    _call("d2K_dwdw", (), out, x1, x2, h, w, p)

def d2K_dwdp(out, x1, x2, h, w, p):
    """periodic_c.pyx:175 -> gpb_periodic_d2K_dwdp"""
    _call("d2K_dwdp", (), out, x1, x2, h, w, p)

def d2K_dpdh(out, x1, x2, h, w, p):
    """periodic_c.pyx:191 -> gpb_periodic_d2K_dpdh"""
    _call("d2K_dpdh", (), out, x1, x2, h, w, p)

def d2K_dpdw(out, x1, x2, h, w, p):
    """periodic_c.pyx:207 -> gpb_periodic_d2K_dpdw"""
    _call("d2K_dpdw", (), out, x1, x2, h, w, p)

def d2K_dpdp(out, x1, x2, h, w, p):
    """periodic_c.pyx:223 -> gpb_periodic_d2K_dpdp"""
    _call("d2K_dpdp", (), out, x1, x2, h, w, p)
