"""Gaussian kernel builders with the reference's Cython signatures
(reference: gp/ext/gaussian_c.pyx): ``f(out, x1, x2, h, w) -> None`` with a
caller-allocated C-contiguous float64 ``out``.  Each call binds the C-ABI entry point
``gpb_gaussian_<name>`` of libgpb200.so (host buffers in, host buffers out; the CUDA
builder computes every requested slice from one exp per element)."""
import ctypes

from .. import _lib
from ._host import carray, out_array

__all__ = ['K', 'jacobian', 'hessian', 'dK_dh', 'dK_dw', 'd2K_dhdh', 'd2K_dhdw', 'd2K_dwdh', 'd2K_dwdw']


def _call(name, lead, out, x1, x2, h, w):
    carray(x1, 1, "x1")
    carray(x2, 1, "x2")
    out_array(out, lead + (x1.size, x2.size))
    _lib.call("gpb_gaussian_" + name, out.ctypes.data, x1.ctypes.data, x1.size, x2.ctypes.data, x2.size,
              float(h), float(w))


def K(out, x1, x2, h, w):
    """gaussian_c.pyx:18 -> gpb_gaussian_K"""
    _call("K", (), out, x1, x2, h, w)

def jacobian(out, x1, x2, h, w):
    """gaussian_c.pyx:39 -> gpb_gaussian_jacobian"""
    _call("jacobian", (2,), out, x1, x2, h, w)

def hessian(out, x1, x2, h, w):
    """gaussian_c.pyx:44 -> gpb_gaussian_hessian"""
    _call("hessian", (2, 2), out, x1, x2, h, w)

def dK_dh(out, x1, x2, h, w):
    """gaussian_c.pyx:51 -> gpb_gaussian_dK_dh"""
    _call("dK_dh", (), out, x1, x2, h, w)

def dK_dw(out, x1, x2, h, w):
    """gaussian_c.pyx:72 -> gpb_gaussian_dK_dw"""
    _call("dK_dw", (), out, x1, x2, h, w)

def d2K_dhdh(out, x1, x2, h, w):
    """gaussian_c.pyx:95 -> gpb_gaussian_d2K_dhdh"""
    _call("d2K_dhdh", (), out, x1, x2, h, w)

def d2K_dhdw(out, x1, x2, h, w):
    """gaussian_c.pyx:116 -> gpb_gaussian_d2K_dhdw"""
    _call("d2K_dhdw", (), out, x1, x2, h, w)

def d2K_dwdh(out, x1, x2, h, w):
    """gaussian_c.pyx:139 -> gpb_gaussian_d2K_dwdh"""
    _call("d2K_dwdh", (), out, x1, x2, h, w)

def d2K_dwdw(out, x1, x2, h, w):
    """gaussian_c.pyx:143 -> gpb_gaussian_d2K_dwdw"""
    _call("d2K_dwdw", (), out, x1, x2, h, w)
