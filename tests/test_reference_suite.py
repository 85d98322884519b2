"""The reference's own test-suite, restated for pytest >= 4 / python 3 and run against
the CUDA implementation through the unchanged public API.

Source of the checks (same grids, seed 2348, dtheta = 1e-5, rtol = 1e-5, <5 % allowed
failures): gp/tests/test_kernels.py:13-112, test_gaussian_kernel.py:29-152,
test_periodic_kernel.py:30-202, test_gp.py:37-333 (yield-tests became loops /
parametrize; ``test_plot`` needs matplotlib, which this image lacks)."""
import numpy as np
import pytest
import scipy.stats

import gaussian_processes_b200 as gpb
from gaussian_processes_b200 import GP, GaussianKernel, PeriodicKernel
from suite_util import (OPT, DTHETA, seed, rand_params, central, make_xy, make_xo, allclose,
                        INVALID_X, INVALID_Y, INVALID_H, INVALID_W)

pytestmark = pytest.mark.gpu
DTYPE = np.float64
KERNELS = [(GaussianKernel, "hw"), (PeriodicKernel, "hwp")]


def random_kernel(cls, names):
    return cls(*rand_params(*names))


# ------------------------------------------------------------------ kernels
def test_gaussian_K_closed_form():
    """test_gaussian_kernel.py:44-64: K == h^2 * normal pdf."""
    seed()
    x = np.linspace(-2, 2, 10)
    dx = x[:, None] - x[None, :]
    for _ in range(OPT["n_big"]):
        k = random_kernel(GaussianKernel, "hw")
        K1 = k(x, x)
        K2 = np.empty_like(K1)
        assert k(x, x, out=K2) is K2
        h, w = k.params
        pdx = h ** 2 * scipy.stats.norm.pdf(dx, loc=0, scale=w)
        assert allclose(pdx, K1) and allclose(pdx, K2) and allclose(K1, K2)


def test_periodic_K_closed_form():
    """test_periodic_kernel.py:47-64."""
    seed()
    x = np.linspace(-2 * np.pi, 2 * np.pi, 16)
    dx = x[:, None] - x[None, :]
    for _ in range(OPT["n_big"]):
        k = random_kernel(PeriodicKernel, "hwp")
        K1 = k(x, x)
        K2 = np.empty_like(K1)
        k(x, x, out=K2)
        h, w, p = k.params
        ref = (h ** 2) * np.exp(-2. * (np.sin(dx / (2. * p)) ** 2) / (w ** 2))
        assert allclose(ref, K1) and allclose(ref, K2) and allclose(K1, K2)


@pytest.mark.parametrize("cls,params", [(GaussianKernel, (1, 1)), (GaussianKernel, (0.5, 0.5)),
                                        (PeriodicKernel, (1, 1, 1)), (PeriodicKernel, (0.5, 0.5, 2))])
def test_sym_K(cls, params):
    """test_gaussian_kernel.py:67-88 / test_periodic_kernel.py:67-90."""
    x = np.linspace(-2, 2, 3)
    dx = x[:, None] - x[None, :]
    k = cls(*params)
    K = k(x, x)
    Ks = np.empty_like(K)
    for i in range(x.size):
        for j in range(x.size):
            Ks[i, j] = k.sym_K.evalf(subs=dict(zip(("d",) + cls._names, (dx[i, j],) + tuple(params))))
    assert allclose(Ks, K)


@pytest.mark.parametrize("cls,names", KERNELS)
def test_jacobian_fd(cls, names):
    """test_kernels.py:29-48."""
    seed()
    x = np.linspace(-2, 2, 10) if cls is GaussianKernel else np.linspace(-2 * np.pi, 2 * np.pi, 16)
    for _ in range(OPT["n_small"]):
        k = random_kernel(cls, names)
        params = k.params.copy()
        jac1 = k.jacobian(x, x)
        jac2 = np.empty_like(jac1)
        k.jacobian(x, x, out=jac2)
        approx = np.empty(jac1.shape)
        for i in range(len(params)):
            p0, p1 = list(params), list(params)
            p0[i] -= DTHETA
            p1[i] += DTHETA
            approx[i] = central(cls(*p0)(x, x), cls(*p1)(x, x), DTHETA)
        assert allclose(jac1, approx) and allclose(jac2, approx) and allclose(jac1, jac2)


@pytest.mark.parametrize("cls,names", KERNELS)
def test_hessian_fd(cls, names):
    """test_kernels.py:72-91."""
    seed()
    x = np.linspace(-2, 2, 10) if cls is GaussianKernel else np.linspace(-2 * np.pi, 2 * np.pi, 16)
    for _ in range(OPT["n_small"]):
        k = random_kernel(cls, names)
        params = k.params.copy()
        h1 = k.hessian(x, x)
        h2 = np.empty_like(h1)
        k.hessian(x, x, out=h2)
        approx = np.empty(h1.shape)
        for i in range(len(params)):
            p0, p1 = list(params), list(params)
            p0[i] -= DTHETA
            p1[i] += DTHETA
            approx[:, i] = central(cls(*p0).jacobian(x, x), cls(*p1).jacobian(x, x), DTHETA)
        assert allclose(h1, approx) and allclose(h2, approx) and allclose(h1, h2)


@pytest.mark.parametrize("cls,names", KERNELS)
def test_slice_methods_fd(cls, names):
    """test_kernels.py:51-69, 94-112: every dK_d* / d2K_d*d* method, return and out= paths."""
    seed()
    x = np.linspace(-2, 2, 10) if cls is GaussianKernel else np.linspace(-2 * np.pi, 2 * np.pi, 16)
    for _ in range(3):
        k = random_kernel(cls, names)
        params = k.params.copy()
        J, H = k.jacobian(x, x), k.hessian(x, x)
        for i, a in enumerate(names):
            f = getattr(k, "dK_d%s" % a)
            d1 = f(x, x)
            d2 = np.empty_like(d1)
            f(x, x, out=d2)
            p0, p1 = list(params), list(params)
            p0[i] -= DTHETA
            p1[i] += DTHETA
            approx = central(cls(*p0)(x, x), cls(*p1)(x, x), DTHETA)
            assert allclose(d1, approx) and allclose(d2, approx) and np.array_equal(d1, J[i])
            for j, b in enumerate(names):
                f2 = getattr(k, "d2K_d%sd%s" % (a, b))
                e1 = f2(x, x)
                e2 = np.empty_like(e1)
                f2(x, x, out=e2)
                q0, q1 = list(params), list(params)
                q0[j] -= DTHETA
                q1[j] += DTHETA
                approx2 = central(getattr(cls(*q0), "dK_d%s" % a)(x, x),
                                  getattr(cls(*q1), "dK_d%s" % a)(x, x), DTHETA)
                assert allclose(e1, approx2) and allclose(e2, approx2) and np.array_equal(e1, H[i, j])


# ------------------------------------------------------------------ GP
def make_gp():
    x, y = make_xy()
    return GP(GaussianKernel(1, 1), x, y, s=1)


def make_random_gp():
    x, y = make_xy()
    h, w, s = rand_params("h", "w", "s")
    return GP(GaussianKernel(h, w), x, y, s=s)


def count_failures(check, n):
    """test_gp.py:37-54."""
    seed()
    failures = []
    for _ in range(n):
        gp = make_random_gp()
        try:
            check(gp)
        except AssertionError:
            failures.append(tuple(gp.params))
    pfail = 100 * len(failures) / n
    assert pfail < OPT["pct_fail"], "%s failed %d/%d (%.1f%%) times: %s" % (
        check.__name__, len(failures), n, pfail, failures[:3])


def fd_params(gp, f):
    """Central difference of f(gp) over every parameter, via gp.copy(); gp.params = ..."""
    params = gp.params
    cols = []
    for i in range(len(params)):
        p0, p1 = list(params), list(params)
        p0[i] -= DTHETA
        p1[i] += DTHETA
        g0, g1 = gp.copy(), gp.copy()
        g0.params = p0
        g1.params = p1
        cols.append(central(f(g0), f(g1), DTHETA))
    return cols


def test_mean():
    def check_mean(gp):                                    # test_gp.py:59-64
        gp.s = 0
        assert allclose(gp.mean(gp.x), gp.y)
    count_failures(check_mean, OPT["n_big"])


def test_inv():
    def check_inv(gp):                                     # test_gp.py:67-72
        I = np.dot(gp.Kxx, gp.inv_Kxx)
        assert allclose(I, np.eye(I.shape[0]))
    count_failures(check_inv, OPT["n_small"])


def test_dloglh():
    def check_dloglh(gp):                                  # test_gp.py:75-98
        assert allclose(gp.dloglh_dtheta, np.array(fd_params(gp, lambda g: g.log_lh)))
    count_failures(check_dloglh, OPT["n_big"])


def test_dlh():
    def check_dlh(gp):                                     # test_gp.py:101-124
        assert allclose(gp.dlh_dtheta, np.array(fd_params(gp, lambda g: g.lh)))
    count_failures(check_dlh, OPT["n_big"])


def test_d2lh():
    def check_d2lh(gp):                                    # test_gp.py:127-150
        approx = np.stack(fd_params(gp, lambda g: g.dlh_dtheta), axis=1)
        assert allclose(gp.d2lh_dtheta2, approx)
    count_failures(check_d2lh, OPT["n_big"])


def test_dm():
    xo = make_xo()

    def check_dm(gp):                                      # test_gp.py:153-174
        assert allclose(gp.dm_dtheta(xo), np.array(fd_params(gp, lambda g: g.mean(xo))))
    count_failures(check_dm, OPT["n_big"])


def test_dtypes_and_shapes():
    """test_gp.py:177-242."""
    gp = make_gp()
    xo = make_xo()
    n, m, n_p = gp.x.size, xo.size, gp.params.size
    shapes = dict(x=(n,), y=(n,), params=(n_p,), Kxx=(n, n), Kxx_J=(n_p - 1, n, n),
                  Kxx_H=(n_p - 1, n_p - 1, n, n), Lxx=(n, n), inv_Kxx=(n, n), inv_Kxx_y=(n,),
                  dloglh_dtheta=(n_p,), dlh_dtheta=(n_p,), d2lh_dtheta2=(n_p, n_p))
    for name, shape in shapes.items():
        v = getattr(gp, name)
        assert isinstance(v, np.ndarray) and v.dtype == DTYPE and v.shape == shape, name
    for name in ("s", "log_lh", "lh"):
        assert type(getattr(gp, name)) == DTYPE, name
    fshapes = dict(Kxoxo=(m, m), Kxxo=(n, m), Kxox=(m, n), mean=(m,), cov=(m, m), dm_dtheta=(n_p, m))
    for name, shape in fshapes.items():
        v = getattr(gp, name)(xo)
        assert isinstance(v, np.ndarray) and v.dtype == DTYPE and v.shape == shape, name


def test_memoprop_del_and_reset():
    """test_gp.py:245-279 with real values."""
    gp = make_gp()
    for prop in ("Kxx", "Kxx_J", "Kxx_H", "Lxx", "inv_Kxx", "inv_Kxx_y", "log_lh", "lh",
                 "dloglh_dtheta", "dlh_dtheta", "d2lh_dtheta2"):
        getattr(gp, prop)
        assert prop in gp._memoized
        delattr(gp, prop)
        assert prop not in gp._memoized
    for prop, val in (("x", gp.x.copy() + 1), ("y", gp.y.copy() + 1), ("s", gp.s + 1), ("params", gp.params + 1)):
        gp.Kxx
        assert gp._memoized != {}
        setattr(gp, prop, val)
        assert gp._memoized == {}


def test_invalid_params():
    """test_gp.py:298-333 -- the suite's hard-coded known-answer case."""
    gp = GP(GaussianKernel(INVALID_H, INVALID_W), INVALID_X, INVALID_Y, s=0)
    for prop in ("Lxx", "inv_Kxx", "inv_Kxx_y"):
        with pytest.raises(np.linalg.LinAlgError):
            getattr(gp, prop)
    assert gp.log_lh == -np.inf
    assert gp.lh == 0
    assert np.isnan(gp.dloglh_dtheta).all()
    assert np.isnan(gp.dlh_dtheta).all()
    assert np.isnan(gp.d2lh_dtheta2).all()


def test_pickle_roundtrip_with_values():
    import pickle
    gp1 = make_gp()
    llh = gp1.log_lh
    gp2 = pickle.loads(pickle.dumps(gp1))
    assert gp2._memoized["log_lh"] == llh
    del gp2.log_lh
    assert gp2.log_lh == llh                               # recomputed on the device after unpickling
