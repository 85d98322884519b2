"""``gp.ext`` of the reference (gp/ext/__init__.py:1-5): the native layer's namespace.
Same three module names and function signatures; the bodies are the sm_100a CUDA path."""
from . import gaussian_c
from . import periodic_c
from . import gp_c

__all__ = ["gaussian_c", "periodic_c", "gp_c"]
