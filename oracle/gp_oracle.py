"""oracle/gp_oracle.py -- TEST INFRASTRUCTURE ONLY (never imported by the product).

CPU restatement of the GP-regression hot path of jhamrick/gaussian_processes
v1.0.5, written from the reference's algorithm (not its text) with every
function citing the reference ``file:line`` it follows.  numpy/scipy do the
linear algebra exactly where the reference calls them; the element loops of the
kernel builders are the plain-C restatement in ``oracle/kernels_oracle.c``
(built to ``oracle/libgporacle.so``), or -- ``impl="ref"`` -- the reference's
own Cython compiled in place into ``oracle/_ref/`` by ``oracle/build_ref.py``.

Parity status: PINNED.  ``tests/test_oracle.py`` checks this file against
(a) ``oracle/_ref`` (reference's compiled native layer) on seeded inputs,
(b) the golden vectors in ``tests/golden/*.npz`` that were generated in the
build container by importing the *unmodified* Python reference from
``/root/reference`` (``tests/golden/make_golden.py``), and
(c) the one hard-coded known-answer case of the reference suite
(``gp/tests/test_gp.py:298-333``).

Unpinned: ``fit_MLII`` (absent from the reference tree, CHANGELOG.md:16-19);
``oracle_fit_mlii`` below is the CPU loop the product is compared against.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s cpu_baseline /
``--impl reference`` legs may import this module.
"""
import ctypes
import os
import subprocess

import numpy as np
import scipy.linalg

HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(HERE, "libgporacle.so")

DTYPE = np.float64
#: gp/gp.py:17, gp_c.pyx:14, gaussian_c.pyx:15
MIN = np.log(np.exp2(DTYPE(np.finfo(DTYPE).minexp + 4)))
EPS = np.finfo(DTYPE).eps

GAUSSIAN, PERIODIC = 0, 1
N_KP = {GAUSSIAN: 2, PERIODIC: 3}


def build_c(force=False):
    """gcc the C restatement -> oracle/libgporacle.so."""
    src = os.path.join(HERE, "kernels_oracle.c")
    if force or not os.path.exists(_LIB_PATH) or \
            os.path.getmtime(_LIB_PATH) < os.path.getmtime(src):
        subprocess.check_call(["gcc", "-O2", "-shared", "-fPIC", src, "-o", _LIB_PATH, "-lm"])
    return _LIB_PATH


_lib = None


def _clib():
    global _lib
    if _lib is None:
        build_c()
        lib = ctypes.CDLL(_LIB_PATH)
        dp = ctypes.POINTER(ctypes.c_double)
        lib.gpo_gaussian_slice.argtypes = [ctypes.c_int, dp, dp, ctypes.c_int64, dp,
                                           ctypes.c_int64, ctypes.c_double, ctypes.c_double]
        lib.gpo_periodic_slice.argtypes = [ctypes.c_int, dp, dp, ctypes.c_int64, dp,
                                           ctypes.c_int64, ctypes.c_double, ctypes.c_double,
                                           ctypes.c_double]
        lib.gpo_min_log.restype = ctypes.c_double
        _lib = lib
    return _lib


_ref_mods = None


def _ref():
    global _ref_mods
    if _ref_mods is None:
        import importlib.util
        spec = importlib.util.spec_from_file_location("_gpo_build_ref",
                                                      os.path.join(HERE, "build_ref.py"))
        br = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(br)
        _ref_mods = br.load()
    return _ref_mods


def have_ref():
    try:
        _ref()
        return True
    except Exception:
        return False


def _dptr(a):
    return a.ctypes.data_as(ctypes.POINTER(ctypes.c_double))


# ----------------------------------------------------------------------------
# kernel builders  (gaussian_c.pyx / periodic_c.pyx; dispatch as in
# gp/kernels/gaussian.py:90-144 and periodic.py:100-190)
# ----------------------------------------------------------------------------

_G_NAMES = ["K", "dK_dh", "dK_dw", "d2K_dhdh", "d2K_dhdw", "d2K_dwdh", "d2K_dwdw"]
_P_NAMES = ["K", "dK_dh", "dK_dw", "dK_dp",
            "d2K_dhdh", "d2K_dhdw", "d2K_dhdp",
            "d2K_dwdh", "d2K_dwdw", "d2K_dwdp",
            "d2K_dpdh", "d2K_dpdw", "d2K_dpdp"]


def slice_names(kind):
    return _G_NAMES if kind == GAUSSIAN else _P_NAMES


def kernel_slice(kind, sl, x1, x2, kparams, impl="c"):
    """One [n1, n2] slice.  Slice ids: 0 = K, 1..n_p = jacobian, then hessian row-major."""
    x1 = np.ascontiguousarray(x1, dtype=DTYPE)
    x2 = np.ascontiguousarray(x2, dtype=DTYPE)
    out = np.empty((x1.size, x2.size), dtype=DTYPE)
    kp = [float(v) for v in kparams]
    if impl == "ref":
        g, p, _ = _ref()
        getattr(g if kind == GAUSSIAN else p, slice_names(kind)[sl])(out, x1, x2, *kp)
    elif kind == GAUSSIAN:
        _clib().gpo_gaussian_slice(sl, _dptr(out), _dptr(x1), x1.size, _dptr(x2), x2.size, *kp)
    else:
        _clib().gpo_periodic_slice(sl, _dptr(out), _dptr(x1), x1.size, _dptr(x2), x2.size, *kp)
    return out


def K(kind, x1, x2, kparams, impl="c"):
    """gaussian_c.K (:18-36) / periodic_c.K (:18-30)."""
    return kernel_slice(kind, 0, x1, x2, kparams, impl)


def jacobian(kind, x1, x2, kparams, impl="c"):
    """gaussian_c.jacobian (:39-41) / periodic_c.jacobian (:33-36): [n_p, n1, n2]."""
    n_p = N_KP[kind]
    return np.stack([kernel_slice(kind, 1 + i, x1, x2, kparams, impl) for i in range(n_p)])


def hessian(kind, x1, x2, kparams, impl="c"):
    """gaussian_c.hessian (:44-48) / periodic_c.hessian (:39-50): [n_p, n_p, n1, n2]."""
    n_p = N_KP[kind]
    x1 = np.asarray(x1)
    x2 = np.asarray(x2)
    out = np.empty((n_p, n_p, x1.size, x2.size), dtype=DTYPE)
    for i in range(n_p):
        for j in range(n_p):
            out[i, j] = kernel_slice(kind, 1 + n_p + i * n_p + j, x1, x2, kparams, impl)
    return out


def gaussian_closed_form(x1, x2, h, w):
    """Independent closed form the reference suite pins K against
    (gp/tests/test_gaussian_kernel.py:44-64): h^2 * normal_pdf(dx; 0, w)."""
    dx = np.subtract.outer(x1, x2)
    return h ** 2 * np.exp(-0.5 * (dx / w) ** 2) / (w * np.sqrt(2 * np.pi))


def periodic_closed_form(x1, x2, h, w, p):
    """gp/tests/test_periodic_kernel.py:47-64."""
    dx = np.subtract.outer(x1, x2)
    return (h ** 2) * np.exp(-2. * (np.sin(dx / (2. * p)) ** 2) / (w ** 2))


# ----------------------------------------------------------------------------
# gp_c reductions  (gp/ext/gp_c.pyx)
# ----------------------------------------------------------------------------

def log_lh_reduce(y, Kxx, Kiy):
    """gp_c.log_lh (gp_c.pyx:17-31): LU slogdet, -inf clamp, three-term sum."""
    sign, logdet = np.linalg.slogdet(Kxx)
    if sign != 1 or logdet < MIN:
        return -np.inf
    fit = -0.5 * float(np.dot(y, Kiy))
    penalty = -0.5 * logdet
    const = -0.5 * y.size * np.log(2 * np.pi)
    return fit + penalty + const


def _dK_list(Kj, s):
    """dK_i for i < n_p from the jacobian, 2*s*I for the noise row (gp_c.pyx:41-45)."""
    n = Kj.shape[1]
    return [Kj[i] for i in range(Kj.shape[0])] + [np.eye(n) * 2 * s]


def dloglh_reduce(y, Ki, Kj, Kiy, s):
    """gp_c.dloglh_dtheta (gp_c.pyx:34-49)."""
    out = np.empty(Kj.shape[0] + 1)
    for i, dK in enumerate(_dK_list(Kj, s)):
        k = np.dot(Ki, dK)
        out[i] = 0.5 * np.dot(y, np.dot(k, Kiy)) + -0.5 * np.trace(k)
    return out


def dlh_reduce(y, Ki, Kj, Kiy, s, lh):
    """gp_c.dlh_dtheta (gp_c.pyx:52-67)."""
    out = np.empty(Kj.shape[0] + 1)
    for i, dK in enumerate(_dK_list(Kj, s)):
        k = np.dot(Ki, dK)
        out[i] = 0.5 * lh * (np.dot(y, np.dot(k, Kiy)) - np.trace(k))
    return out


def d2lh_reduce(y, Ki, Kj, Kh, Kiy, s, lh, dlh):
    """gp_c.d2lh_dtheta2 (gp_c.pyx:70-111)."""
    n_p = Kj.shape[0]
    m = Kj.shape[1]
    dK = _dK_list(Kj, s)
    dKi = [np.dot(-Ki, np.dot(d, Ki)) for d in dK]              # :88
    out = np.empty((n_p + 1, n_p + 1))
    for i in range(n_p + 1):
        KidK = np.dot(Ki, dK[i])                                # :91
        r_i = np.dot(y, np.dot(KidK, Kiy)) - np.trace(KidK)     # :92-93
        for j in range(n_p + 1):
            if i < n_p and j < n_p:                             # :97-102
                d2k = Kh[i, j]
            elif i == n_p and j == n_p:
                d2k = np.eye(m) * 2
            else:
                d2k = np.zeros((m, m))
            G = np.dot(dKi[j], dK[i])                           # :104
            t0 = dlh[j] * r_i
            t1a = np.dot(y, np.dot(G, Kiy))
            t1b = np.dot(Kiy, np.dot(d2k, Kiy))
            t1c = np.dot(Kiy, np.dot(dK[i], np.dot(dKi[j], y)))
            t1 = lh * (t1a + t1b + t1c - np.trace(G + np.dot(Ki, d2k)))
            out[i, j] = 0.5 * (t0 + t1)                         # :111
    return out


def dm_reduce(y, Ki, Kj, Kjxo, Kxox, s):
    """gp_c.dm_dtheta (gp_c.pyx:114-131)."""
    n_p = Kj.shape[0]
    m2 = Kjxo.shape[1]
    out = np.empty((n_p + 1, m2))
    dK = _dK_list(Kj, s)
    for i in range(n_p + 1):
        dKxo = Kjxo[i] if i < n_p else np.zeros_like(Kxox)
        out[i] = np.dot(dKxo, np.dot(Ki, y))
        out[i] -= np.dot(Kxox, np.dot(np.dot(Ki, np.dot(dK[i], Ki)), y))
    return out


# ----------------------------------------------------------------------------
# GP object  (gp/gp.py)
# ----------------------------------------------------------------------------

class OracleGP(object):
    """Restatement of gp.GP's numerics (gp/gp.py:242-662) without the memo plumbing.

    ``impl="c"`` uses the C restatement for kernel matrices and the numpy
    restatement above for the reductions; ``impl="ref"`` routes both through the
    reference's compiled Cython in oracle/_ref (what gp.py itself calls).
    """

    def __init__(self, kind, kparams, x, y, s=0.0, impl="c"):
        self.kind = kind
        self.kparams = np.array(kparams, dtype=DTYPE)
        self.x = np.array(x, dtype=DTYPE)
        self.y = np.array(y, dtype=DTYPE)
        self.s = DTYPE(s)
        self.impl = impl
        self._c = {}

    @property
    def params(self):
        return np.concatenate([self.kparams, [self.s]])

    def _memo(self, name, fn):
        if name not in self._c:
            self._c[name] = fn()
        return self._c[name]

    # gp.py:242-266 -- index-diagonal s^2 (np.eye), not x-equality
    @property
    def Kxx(self):
        def f():
            Kx = K(self.kind, self.x, self.x, self.kparams, self.impl)
            Kx += np.eye(self.x.size, dtype=DTYPE) * (self.s ** 2)
            return Kx
        return self._memo("Kxx", f)

    @property
    def Kxx_J(self):   # gp.py:268-271
        return self._memo("Kxx_J", lambda: jacobian(self.kind, self.x, self.x,
                                                    self.kparams, self.impl))

    @property
    def Kxx_H(self):   # gp.py:273-276
        return self._memo("Kxx_H", lambda: hessian(self.kind, self.x, self.x,
                                                   self.kparams, self.impl))

    @property
    def Lxx(self):     # gp.py:278-294 (raises np.linalg.LinAlgError when not PD)
        return self._memo("Lxx", lambda: scipy.linalg.cholesky(
            self.Kxx, lower=True, overwrite_a=False, check_finite=True))

    @property
    def inv_Kxx(self):  # gp.py:296-312 -- inv(L)^T inv(L)
        def f():
            iL = np.linalg.inv(self.Lxx)
            return np.dot(iL.T, iL)
        return self._memo("inv_Kxx", f)

    @property
    def inv_Kxx_y(self):  # gp.py:314-335
        return self._memo("inv_Kxx_y", lambda: scipy.linalg.cho_solve(
            (self.Lxx, True), self.y, overwrite_b=False, check_finite=True))

    @property
    def log_lh(self):   # gp.py:337-367
        def f():
            try:
                Kiy = self.inv_Kxx_y
            except np.linalg.LinAlgError:
                return -np.inf
            if self.impl == "ref":
                return DTYPE(_ref()[2].log_lh(self.y, self.Kxx, Kiy))
            return DTYPE(log_lh_reduce(self.y, self.Kxx, Kiy))
        return self._memo("log_lh", f)

    @property
    def lh(self):       # gp.py:369-396 -- int 0 below MIN
        def f():
            llh = self.log_lh
            return 0 if llh < MIN else np.exp(llh)
        return self._memo("lh", f)

    def _nan(self, *shape):
        out = np.empty(shape)
        out.fill(np.nan)
        return out

    @property
    def dloglh_dtheta(self):  # gp.py:398-433
        def f():
            try:
                Ki = self.inv_Kxx
            except np.linalg.LinAlgError:
                return self._nan(len(self.params))
            if self.impl == "ref":
                out = np.empty(len(self.params))
                _ref()[2].dloglh_dtheta(self.y, Ki, self.Kxx_J, self.inv_Kxx_y, self.s, out)
                return out
            return dloglh_reduce(self.y, Ki, self.Kxx_J, self.inv_Kxx_y, self.s)
        return self._memo("dloglh_dtheta", f)

    def dlh_dtheta_with(self, lh):
        """gp_c.dlh_dtheta with an explicit ``lh`` (lh=1.0 gives the
        lh-normalised derivative where the real lh underflows to 0)."""
        Ki = self.inv_Kxx
        if self.impl == "ref":
            out = np.empty(len(self.params))
            _ref()[2].dlh_dtheta(self.y, Ki, self.Kxx_J, self.inv_Kxx_y, self.s, lh, out)
            return out
        return dlh_reduce(self.y, Ki, self.Kxx_J, self.inv_Kxx_y, self.s, lh)

    @property
    def dlh_dtheta(self):     # gp.py:435-466
        def f():
            try:
                self.inv_Kxx
            except np.linalg.LinAlgError:
                return self._nan(len(self.params))
            return self.dlh_dtheta_with(self.lh)
        return self._memo("dlh_dtheta", f)

    def d2lh_dtheta2_with(self, lh, dlh):
        Ki = self.inv_Kxx
        if self.impl == "ref":
            n = len(self.params)
            out = np.empty((n, n))
            _ref()[2].d2lh_dtheta2(self.y, Ki, self.Kxx_J, self.Kxx_H, self.inv_Kxx_y,
                                   self.s, lh, np.ascontiguousarray(dlh, dtype=DTYPE), out)
            return out
        return d2lh_reduce(self.y, Ki, self.Kxx_J, self.Kxx_H, self.inv_Kxx_y, self.s, lh, dlh)

    @property
    def d2lh_dtheta2(self):   # gp.py:468-502
        def f():
            try:
                self.inv_Kxx
            except np.linalg.LinAlgError:
                return self._nan(len(self.params), len(self.params))
            return self.d2lh_dtheta2_with(self.lh, self.dlh_dtheta)
        return self._memo("d2lh_dtheta2", f)

    # gp.py:504-572
    def Kxoxo(self, xo):
        return K(self.kind, xo, xo, self.kparams, self.impl)

    def Kxxo(self, xo):
        return K(self.kind, self.x, xo, self.kparams, self.impl)

    def Kxox(self, xo):
        return K(self.kind, xo, self.x, self.kparams, self.impl)

    def mean(self, xo):       # gp.py:574-597
        return np.dot(self.Kxox(xo), self.inv_Kxx_y)

    def cov(self, xo):        # gp.py:599-625 -- explicit inverse
        return self.Kxoxo(xo) - np.dot(self.Kxox(xo), np.dot(self.inv_Kxx, self.Kxxo(xo)))

    def dm_dtheta(self, xo):  # gp.py:627-662
        Kjxo = jacobian(self.kind, xo, self.x, self.kparams, self.impl)
        Kxox = self.Kxox(xo)
        if self.impl == "ref":
            out = np.empty((len(self.params), np.asarray(xo).size))
            _ref()[2].dm_dtheta(self.y, self.inv_Kxx, self.Kxx_J, Kjxo, Kxox, self.s, out)
            return out
        return dm_reduce(self.y, self.inv_Kxx, self.Kxx_J, Kjxo, Kxox, self.s)


def oracle_fit_mlii(kind, x, y, candidates, impl="c"):
    """CPU definition of the batched MLII search (no reference implementation
    exists -- CHANGELOG.md:16-19 -- so this loop over reference-equivalent GP
    evaluations IS the specification; parity of the *search* is unpinned, parity
    of every per-candidate value is pinned through OracleGP).

    candidates: [B, n_theta] rows (kernel params..., s).  Returns
    (best_index, log_lh[B], dloglh[B, n_theta]); the best index is the lowest
    index attaining the maximum finite-or--inf log_lh, NaN ordered last.
    """
    cand = np.asarray(candidates, dtype=DTYPE)
    B, nth = cand.shape
    llh = np.empty(B)
    grad = np.empty((B, nth))
    for b in range(B):
        g = OracleGP(kind, cand[b, :-1], x, y, cand[b, -1], impl)
        llh[b] = g.log_lh
        grad[b] = g.dloglh_dtheta
    key = np.where(np.isnan(llh), -np.inf, llh)
    best = int(np.argmax(key))
    return best, llh, grad
