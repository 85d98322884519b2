from .base import Kernel
from .periodic import PeriodicKernel
from .gaussian import GaussianKernel
from .symbolic import SymbolicKernel

__all__ = ["Kernel", "PeriodicKernel", "GaussianKernel", "SymbolicKernel"]
