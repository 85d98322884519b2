"""Kernel base class: the reference's public kernel interface
(reference: gp/kernels/base.py:7-121), kept verbatim at the API level --
``K/__call__/jacobian/hessian(x1, x2, out=None)``, ``params``, ``copy`` and
pickling by parameters -- over the CUDA builders."""
import copy as _copy

import numpy as np

__all__ = ["Kernel"]

DTYPE = np.float64
EPS = np.finfo(DTYPE).eps


class Kernel(object):
    #: ordered parameter names, set by subclasses
    _names = ()
    #: module of gaussian_processes_b200.ext holding the builders for this kernel
    _ext = None

    # -- state / copies: a kernel is fully described by its parameters (base.py:9-32)
    def __getstate__(self):
        return {"params": self.params}

    def __setstate__(self, state):
        self.params = state["params"]

    def __copy__(self):
        return type(self)(*self.params)

    def __deepcopy__(self, memo):
        return type(self)(*self.params)

    def copy(self):
        """New kernel object of the same type with the same parameters."""
        return _copy.copy(self)

    # -- parameters (gaussian.py:44-73, periodic.py:48-83)
    @property
    def params(self):
        return np.array([getattr(self, n) for n in self._names], dtype=DTYPE)

    @params.setter
    def params(self, val):
        for n, v in zip(self._names, val):
            self.set_param(n, v)

    def set_param(self, name, val):
        if name not in self._names:
            raise ValueError("unknown parameter: %s" % name)
        if val < EPS:
            raise ValueError("invalid value for %s: %s" % (name, val))
        setattr(self, name, DTYPE(val))

    @property
    def sym_K(self):
        raise NotImplementedError

    # -- builders
    def _build(self, fname, lead, x1, x2, out):
        if out is None:
            out = np.empty(lead + (x1.size, x2.size), dtype=DTYPE)
        getattr(self._ext(), fname)(out, x1, x2, *[getattr(self, n) for n in self._names])
        return out

    def K(self, x1, x2, out=None):
        r"""Kernel function evaluated at `x1` and `x2`: :math:`n\times m` array."""
        return self._build("K", (), x1, x2, out)

    def __call__(self, x1, x2, out=None):
        return self.K(x1, x2, out=out)

    def jacobian(self, x1, x2, out=None):
        r"""Jacobian w.r.t. the kernel parameters: :math:`n_p\times n\times m` array."""
        return self._build("jacobian", (len(self._names),), x1, x2, out)

    def hessian(self, x1, x2, out=None):
        r"""Hessian w.r.t. the kernel parameters: :math:`n_p\times n_p\times n\times m` array."""
        return self._build("hessian", (len(self._names),) * 2, x1, x2, out)


def _add_slice_methods(cls):
    """dK_d<a> and d2K_d<a>d<b> per-slice methods (gaussian.py:110-144, periodic.py:120-190)."""
    def make(fname):
        def method(self, x1, x2, out=None):
            return self._build(fname, (), x1, x2, out)
        method.__name__ = fname
        return method
    for a in cls._names:
        setattr(cls, "dK_d%s" % a, make("dK_d%s" % a))
        for b in cls._names:
            setattr(cls, "d2K_d%sd%s" % (a, b), make("d2K_d%sd%s" % (a, b)))
    return cls
