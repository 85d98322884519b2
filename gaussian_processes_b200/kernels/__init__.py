from .base import Kernel
from .periodic import PeriodicKernel
from .gaussian import GaussianKernel

__all__ = ["Kernel", "PeriodicKernel", "GaussianKernel"]
