"""``GP``: the reference's Gaussian-process object (reference: gp/gp.py:44-700) with the
same constructor, properties, methods, memo/invalidate semantics, copy/pickle state and
error behaviour -- computed by the sm_100a CUDA path instead of numpy/scipy/Cython.

Residency: matrix-valued intermediates (Kxx, its Cholesky factor, the triangular
inverse, K^-1) live on the device in ``self._dev`` and are only copied to the host when
the user reads a matrix-valued property.  ``gp.log_lh`` / ``gp.dloglh_dtheta`` move
``n_theta + 1`` doubles over PCIe, never an N x N array.  ``_memoized`` holds exactly
what the reference's holds (the values the properties returned), so its keys, its
``del`` semantics and its place in the pickled state are unchanged
(gp/tests/test_gp.py:245-279).
"""
import copy as _copy

import numpy as np

from . import engine as _engine

__all__ = ["GP"]

DTYPE = np.float64
EPS = np.finfo(DTYPE).eps
MIN = np.log(np.exp2(DTYPE(np.finfo(DTYPE).minexp + 4)))      # gp.py:17


def fused_kind(kernel):
    """``KIND`` of the fused CUDA functor that evaluates ``kernel``, or None.  A subclass of a
    built-in kernel inherits ``KIND`` but may override ``K`` / ``jacobian`` / ``hessian`` (the
    reference's GP calls exactly these methods, gp.py:264,271,276): the fused functor is only
    taken when the element formulas are the built-in ones."""
    from .kernels import GaussianKernel, PeriodicKernel
    cls = type(kernel)
    kind = getattr(cls, "KIND", None)
    if kind is None:
        return None
    for base in (GaussianKernel, PeriodicKernel):
        if isinstance(kernel, base):
            for name in ("K", "__call__", "jacobian", "hessian", "_build", "_ext"):
                if getattr(cls, name, None) is not getattr(base, name, None):
                    return None
            return kind
    return None


class memoprop(object):
    """Property whose value is computed once and kept in ``obj._memoized[name]``;
    ``del obj.name`` forgets it (semantics of gp.py:20-41)."""

    def __init__(self, fn):
        self.fn = fn
        self.name = fn.__name__
        self.__doc__ = fn.__doc__

    def __get__(self, obj, objtype=None):
        if obj is None:
            return self
        cache = obj._memoized
        if self.name not in cache:
            cache[self.name] = self.fn(obj)
        return cache[self.name]

    def __set__(self, obj, value):
        raise AttributeError("can't set attribute")

    def __delete__(self, obj):
        del obj._memoized[self.name]


class GP(object):
    r"""GP regression on fixed observations, every linear-algebra result held on the B200.

    ``GP(K, x, y, s=0)`` -- same constructor, attributes and memoised properties as the reference's
    ``gp.GP`` (gp/gp.py:44-127):

    * ``K``: a :class:`~gaussian_processes_b200.kernels.Kernel` (``GaussianKernel``, ``PeriodicKernel``,
      a ``SymbolicKernel`` or any user subclass);
    * ``x``, ``y``: the :math:`n` training inputs and targets, 1-D float64 (stored read-only);
    * ``s``: noise standard deviation added to the diagonal of ``Kxx`` as :math:`s^2`, ``s >= 0``.

    Assigning ``x``, ``y``, ``s``, ``K`` or a parameter drops exactly the cached results that depend on it.
    """

    _STATE = ("K", "_x", "_y", "_s", "_memoized")

    def __init__(self, K, x, y, s=0):
        self.K = K
        self._x = None
        self._y = None
        self._s = None
        self._memoized = {}
        self._dev = None
        self._data_gen = 0
        self.x = x
        self.y = y
        self.s = s

    # ------------------------------------------------------------------ state (gp.py:78-127)
    def __getstate__(self):
        return dict((k, getattr(self, k)) for k in self._STATE)

    def __setstate__(self, state):
        for k in self._STATE:
            setattr(self, k, state[k])
        self._dev = None
        self._data_gen = 0

    def __copy__(self):
        new = type(self).__new__(type(self))
        new.__setstate__(self.__getstate__())
        return new

    def __deepcopy__(self, memo):
        new = type(self).__new__(type(self))
        new.__setstate__(_copy.deepcopy(self.__getstate__(), memo))
        return new

    def copy(self, deep=True):
        """Deep (default) or shallow copy of the GP."""
        return _copy.deepcopy(self) if deep else _copy.copy(self)

    def _reset(self, data=False):
        """Every setter empties the memo (gp.py:231-240); only new observations drop the device
        copy of x / y -- new hyperparameters rebind the resident engine."""
        self._memoized = {}
        if data:
            self._dev = None
            # generation of the observations: device-side copies (the batched evaluator's x / y) are
            # keyed on it -- object identities are recycled by the allocator and cannot be
            self._data_gen = getattr(self, "_data_gen", 0) + 1

    # ------------------------------------------------------------------ inputs (gp.py:129-240)
    @property
    def x(self):
        r"""Vector of input locations (read-only array)."""
        return self._x

    @x.setter
    def x(self, val):
        if np.any(val != self._x):
            self._reset(data=True)
            self._x = np.array(val, copy=True, dtype=DTYPE)
            self._x.flags.writeable = False

    @property
    def y(self):
        r"""Vector of input observations (read-only array)."""
        return self._y

    @y.setter
    def y(self, val):
        if np.any(val != self._y):
            self._reset(data=True)
            self._y = np.array(val, copy=True, dtype=DTYPE)
            self._y.flags.writeable = False
            if self._y.shape != self._x.shape:
                raise ValueError("invalid shape for y: %s" % str(self._y.shape))

    @property
    def s(self):
        r"""Standard deviation of the observation noise (numpy.float64)."""
        return self._s

    @s.setter
    def s(self, val):
        if val < 0:
            raise ValueError("invalid value for s: %s" % val)
        if val != self._s:
            self._reset()
            self._s = DTYPE(val)

    @property
    def params(self):
        r"""Kernel parameters followed by the noise parameter :math:`s`."""
        kp = self.K.params
        out = np.empty(kp.size + 1)
        out[:-1] = kp
        out[-1] = self._s
        return out

    @params.setter
    def params(self, val):
        if np.any(self.params != val):
            self._reset()
            self.K.params = val[:-1]
            self.s = val[-1]

    def get_param(self, name):
        return self.s if name == "s" else getattr(self.K, name)

    def set_param(self, name, val):
        if name == "s":
            self.s = val
        elif getattr(self.K, name) != val:
            self._reset()
            self.K.set_param(name, val)

    # ------------------------------------------------------------------ device state
    def _engine(self):
        """Device-side twin of the current (K, x, y, s); dropped by every setter.  The built-in
        kernels have fused CUDA functors (``KIND``); any other ``Kernel`` subclass evaluates its
        matrices in its own Python methods and everything after that runs on the device."""
        kp = tuple(float(v) for v in self.K.params)
        kind = fused_kind(self.K)
        if kind is None:
            key = (("host", id(self.K)), kp, float(self._s))
            if self._dev is None or self._dev[0][0] != key[0]:
                from .generic import HostKernelEngine
                self._dev = (key, HostKernelEngine(self.K, key[2], self._x, self._y))
            elif self._dev[0] != key:
                self._dev = (key, self._dev[1].rebind(kp, key[2]))
            return self._dev[1]
        key = (kind, kp, float(self._s))
        if self._dev is None or self._dev[0][0] != key[0]:
            self._dev = (key, _engine.Engine(key[0], kp, key[2], self._x, self._y))
        elif self._dev[0] != key:           # same observations and kernel family, new hyperparameters
            self._dev = (key, self._dev[1].rebind(kp, key[2]))
        return self._dev[1]

    def _n_theta(self):
        return self.K.params.size + 1

    def _nan(self, *shape):
        out = np.empty(shape)
        out.fill(np.nan)
        return out

    # ------------------------------------------------------------------ matrices (gp.py:242-335)
    @memoprop
    def Kxx(self):
        r"""Kernel covariance matrix :math:`K(x_i, x_j) + s^2\delta_{ij}` (:math:`n\times n`)."""
        e = self._engine()
        from . import device as D
        return D.download_2d(e.Kxx(), e.n, e.n)

    @memoprop
    def Kxx_J(self):
        x = self._x
        return self.K.jacobian(x, x)

    @memoprop
    def Kxx_H(self):
        x = self._x
        return self.K.hessian(x, x)

    @memoprop
    def Lxx(self):
        r"""Lower Cholesky factor of ``Kxx``; raises ``numpy.linalg.LinAlgError`` when
        ``Kxx`` is not positive definite."""
        return self._engine().Lxx_host()

    @memoprop
    def inv_Kxx(self):
        r"""Inverse kernel covariance matrix :math:`\mathbf{K}_{xx}^{-1}` (from the Cholesky factor)."""
        e = self._engine()
        e.require_pd()
        from . import device as D
        return D.download_2d(e.Ki(), e.n, e.n)

    @memoprop
    def inv_Kxx_y(self):
        r""":math:`\mathbf{K}_{xx}^{-1}\mathbf{y}` by Cholesky solves."""
        e = self._engine()
        from . import device as D
        return D.to_host(e.alpha()[:e.n]).copy()

    # ------------------------------------------------------------------ likelihood (gp.py:337-502)
    @memoprop
    def log_lh(self):
        r"""Marginal log likelihood (Eq. 5.8 of Rasmussen & Williams); ``-inf`` when the
        Cholesky fails or :math:`\log|\mathbf{K}_{xx}|` is below ``MIN``."""
        e = self._engine()
        try:
            e.require_pd()
        except np.linalg.LinAlgError:
            return -np.inf
        return DTYPE(e.loglh3()[0])

    @memoprop
    def lh(self):
        r"""Marginal likelihood; integer 0 when ``log_lh`` is below ``MIN``."""
        llh = self.log_lh
        if llh < MIN:
            return 0
        return np.exp(self.log_lh)

    def _grad_bracket(self):
        """(y^T Ki dK_i Ki y, tr(Ki dK_i)) per parameter, or None when Kxx is not PD."""
        e = self._engine()
        try:
            return e.grad_terms()
        except np.linalg.LinAlgError:
            return None

    @memoprop
    def dloglh_dtheta(self):
        r"""Derivative of the marginal log likelihood w.r.t. ``params`` (Eq. 5.9 of R&W)."""
        terms = self._grad_bracket()
        if terms is None:
            return self._nan(self._n_theta())
        t0, t1 = terms
        return 0.5 * t0 + -0.5 * t1                                  # gp_c.pyx:47-49

    @memoprop
    def dlh_dtheta(self):
        r"""Derivative of the marginal likelihood w.r.t. ``params``."""
        terms = self._grad_bracket()
        if terms is None:
            return self._nan(self._n_theta())
        t0, t1 = terms
        return 0.5 * self.lh * (t0 - t1)                             # gp_c.pyx:65-67

    def _d2lh(self, lh, dlh):
        e = self._engine()
        t0, t1 = e.grad_terms()
        G, Q, TP, TH = e.d2_terms()
        nth = self._n_theta()
        out = np.empty((nth, nth))
        for i in range(nth):
            r_i = t0[i] - t1[i]                                      # gp_c.pyx:92-93
            for j in range(nth):
                t1a = -G[i, j]                                       # y^T dKi_j dK_i Kiy
                t1c = -G[j, i]                                       # Kiy^T dK_i dKi_j y
                tr = -TP[i, j] + TH[i, j]                            # trace(dKi_j dK_i + Ki d2k)
                out[i, j] = 0.5 * (dlh[j] * r_i + lh * (t1a + Q[i, j] + t1c - tr))
        return out

    @memoprop
    def d2lh_dtheta2(self):
        r"""Second derivative (Hessian) of the marginal likelihood w.r.t. ``params``."""
        nth = self._n_theta()
        if self._grad_bracket() is None:
            return self._nan(nth, nth)
        return self._d2lh(self.lh, self.dlh_dtheta)

    def d2loglh_normalised(self):
        r"""``gp_c.d2lh_dtheta2`` evaluated with ``lh = 1`` and ``dlh = dloglh_dtheta``: the
        likelihood-normalised second derivative :math:`(\partial^2 p/\partial\theta^2)/p`, usable at
        sizes where ``lh`` underflows to 0 (additive API)."""
        if self._grad_bracket() is None:
            nth = self._n_theta()
            return self._nan(nth, nth)
        return self._d2lh(1.0, self.dloglh_dtheta)

    def d2loglh_dtheta2(self):
        r"""Hessian of the marginal **log** likelihood w.r.t. ``params``,
        :math:`\partial^2 \log p / \partial\theta_i\partial\theta_j = (\partial^2 p)/p - (\partial_i p)(\partial_j p)/p^2`,
        assembled from the likelihood-normalised second derivative and ``dloglh_dtheta`` -- what a
        Newton step on the hyperparameters needs, finite at sizes where ``lh`` itself underflows
        (additive API)."""
        g = self.dloglh_dtheta
        return self.d2loglh_normalised() - np.outer(g, g)

    # ------------------------------------------------------------------ posterior (gp.py:504-662)
    def Kxoxo(self, xo):
        r"""Kernel covariance matrix of new sample locations (:math:`m\times m`)."""
        return self.K(xo, xo)

    def Kxxo(self, xo):
        r"""Kernel covariance between given and new locations (:math:`n\times m`)."""
        return self.K(self._x, xo)

    def Kxox(self, xo):
        r"""Kernel covariance between new and given locations (:math:`m\times n`)."""
        return self.K(xo, self._x)

    def mean(self, xo):
        r"""Predictive mean :math:`K(\mathbf{x^*},\mathbf{x})\mathbf{K}_{xx}^{-1}\mathbf{y}` (Eq. 2.23 of R&W)."""
        return self._engine().mean(np.asarray(xo, dtype=DTYPE))

    def cov(self, xo):
        r"""Predictive covariance (Eq. 2.24 of R&W), :math:`m\times m`."""
        return self._engine().cov(np.asarray(xo, dtype=DTYPE))

    def var(self, xo):
        r"""Predictive variance, the diagonal of ``cov(xo)``, without forming the :math:`m\times m`
        matrix (additive API; what ``plot`` draws)."""
        return self._engine().var(np.asarray(xo, dtype=DTYPE))

    def cov_rows(self, xo, lo, hi):
        r"""Rows ``lo:hi`` of ``cov(xo)`` (:math:`(hi-lo)\times m`): the shard one GPU owns when
        test points are partitioned across devices (additive API)."""
        return self._engine().cov_rows(np.asarray(xo, dtype=DTYPE), int(lo), int(hi))

    def dm_dtheta(self, xo):
        r"""Derivative of the predictive mean w.r.t. ``params``: :math:`n_\theta\times m`."""
        return self._engine().dm(np.asarray(xo, dtype=DTYPE))

    def plot(self, ax=None, xlim=None, color="k", markercolor="r"):
        """Predictive mean with a one-standard-deviation band over ``xlim`` plus the training points (the picture
        of gp/gp.py:664-697; needs matplotlib).  The band comes from :meth:`var` -- the diagonal only, the
        1000 x 1000 covariance the reference forms for it is never built."""
        import matplotlib.pyplot as plt
        axes = plt.gca() if ax is None else ax
        lo, hi = (float(self._x.min()), float(self._x.max())) if xlim is None else xlim
        grid = np.linspace(lo, hi, 1000)
        mu = self.mean(grid)
        sd = np.sqrt(self.var(grid))
        axes.fill_between(grid, mu - sd, mu + sd, color=color, alpha=0.3)
        axes.plot(grid, mu, lw=2, color=color)
        axes.plot(self._x, self._y, "o", ms=5, color=markercolor)
        axes.set_xlim(lo, hi)

    # ------------------------------------------------------------------ batched search (additive)
    def batch_eval(self, thetas, grad=True):
        """log_lh (and dloglh_dtheta) for many parameter vectors on this GP's (x, y).

        thetas : [B, n_theta] rows ordered like ``params``.  Returns ``(log_lh[B],
        dloglh[B, n_theta])``; a candidate whose Kxx is not positive definite gets
        ``-inf`` / NaN exactly like the scalar properties."""
        from .mlii import batch_eval
        return batch_eval(self, thetas, grad=grad)

    def fit_MLII(self, candidates, **kw):
        """Type-II maximum likelihood by batched multi-restart search (see mlii.fit_MLII)."""
        from .mlii import fit_MLII
        return fit_MLII(self, candidates, **kw)
