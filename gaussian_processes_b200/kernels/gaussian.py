"""Gaussian kernel (reference: gp/kernels/gaussian.py:14-144)."""
import numpy as np

from .base import Kernel, _add_slice_methods

__all__ = ["GaussianKernel"]


@_add_slice_methods
class GaussianKernel(Kernel):
    r"""
    Gaussian kernel function,

    .. math:: K(x_1, x_2) = \frac{h^2}{\sqrt{2\pi w^2}}\exp\left(-\frac{(x_1-x_2)^2}{2w^2}\right)

    Parameters
    ----------
    h : float
        Output scale kernel parameter
    w : float
        Input scale kernel parameter
    """
    _names = ("h", "w")
    KIND = 0

    def __init__(self, h, w):
        self.h = None
        self.w = None
        self.set_param("h", h)
        self.set_param("w", w)

    @staticmethod
    def _ext():
        from ..ext import gaussian_c
        return gaussian_c

    @property
    def sym_K(self):
        """Symbolic form of the kernel (gaussian.py:77-87)."""
        import sympy as sym
        h, w, d = sym.Symbol("h"), sym.Symbol("w"), sym.Symbol("d")
        return h ** 2 * (1. / sym.sqrt(2 * sym.pi * w ** 2)) * sym.exp(-d ** 2 / (2.0 * w ** 2))
