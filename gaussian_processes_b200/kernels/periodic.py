"""Periodic kernel (reference: gp/kernels/periodic.py:14-190)."""
import numpy as np

from .base import Kernel, _add_slice_methods

__all__ = ["PeriodicKernel"]


@_add_slice_methods
class PeriodicKernel(Kernel):
    r"""
    Periodic kernel function (Eq. 4.31 of Rasmussen & Williams),

    .. math:: K(x_1, x_2) = h^2\exp\left(\frac{-2\sin^2\left(\frac{x_1-x_2}{2p}\right)}{w^2}\right)

    Parameters
    ----------
    h : float
        Output scale kernel parameter
    w : float
        Input scale kernel parameter
    p : float
        Period kernel parameter
    """
    _names = ("h", "w", "p")
    KIND = 1

    def __init__(self, h, w, p):
        self.h = None
        self.w = None
        self.p = None
        self.set_param("h", h)
        self.set_param("w", w)
        self.set_param("p", p)

    @staticmethod
    def _ext():
        from ..ext import periodic_c
        return periodic_c

    @property
    def sym_K(self):
        """Symbolic form of the kernel (periodic.py:87-97)."""
        import sympy as sym
        h, w, p, d = sym.Symbol("h"), sym.Symbol("w"), sym.Symbol("p"), sym.Symbol("d")
        return h ** 2 * sym.exp(-2. * (sym.sin(d / (2. * p)) ** 2) / w ** 2)
