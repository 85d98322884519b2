"""gaussian_processes_b200 -- B200 (sm_100a) implementation of the GP-regression hot
path of jhamrick/gaussian_processes v1.0.5, behind the reference's own Python API
(reference: gp/__init__.py:1-8 exports ``ext``, ``GP``, ``Kernel``, ``PeriodicKernel``,
``GaussianKernel``).

    from gaussian_processes_b200 import GP, GaussianKernel
    gp = GP(GaussianKernel(1.0, 0.5), x, y, s=1.0)
    gp.log_lh, gp.dloglh_dtheta, gp.mean(xo), gp.cov(xo)

``install_as_gp()`` registers the package under the reference's import name ``gp``.
There is no CPU fallback: importing needs the built ``libgpb200.so`` and computing
needs a CUDA device.
"""
from . import _lib            # noqa: F401  (fails loudly when the CUDA library is missing)
from . import ext
from .gp import GP
from .kernels import Kernel, PeriodicKernel, GaussianKernel, SymbolicKernel
from .mlii import fit_MLII, MLIIResult
from .posterior import sharded_posterior

__all__ = ["ext", "GP", "Kernel", "PeriodicKernel", "GaussianKernel", "SymbolicKernel", "fit_MLII", "MLIIResult",
           "sharded_posterior", "install_as_gp"]
__version__ = "0.1.0"


def install_as_gp():
    """Make ``import gp`` resolve to this package (drop-in for the reference's name)."""
    import sys
    me = sys.modules[__name__]
    sys.modules.setdefault("gp", me)
    for sub in ("ext", "kernels", "gp"):
        sys.modules.setdefault("gp." + sub, sys.modules[__name__ + "." + sub])
    return me
