"""Host-side semantics that need no GPU: parameter validation, copy / pickle state,
memo bookkeeping, argument checking of the ext shims, sharding helpers.
(reference behaviours: SURVEY Appendix C; gp/tests/test_gp.py:245-432,
gp/tests/test_gaussian_kernel.py:155-189, test_periodic_kernel.py:206-246)"""
import pickle
from copy import copy, deepcopy

import numpy as np
import pytest

import gaussian_processes_b200 as gpb
from gaussian_processes_b200 import GP, GaussianKernel, PeriodicKernel
from gaussian_processes_b200.mlii import shard_bounds, select_best
from suite_util import make_xy, rand_params, seed


def make_gp():
    x, y = make_xy()
    return GP(GaussianKernel(1, 1), x, y, s=1)


@pytest.mark.parametrize("cls,names", [(GaussianKernel, "hw"), (PeriodicKernel, "hwp")])
def test_kernel_params(cls, names):
    seed()
    for _ in range(20):
        params = rand_params(*names)
        k = cls(*params)
        assert (k.params == np.array(params)).all()
        k.params = params
        assert (k.params == np.array(params)).all()
        assert k.params is not k.params and k.params.dtype == np.float64
    good = rand_params(*names)
    for i in range(len(names)):
        bad = list(good)
        bad[i] = 0
        with pytest.raises(ValueError):
            cls(*bad)
        k = cls(*good)
        with pytest.raises(ValueError):
            k.params = bad
    with pytest.raises(ValueError):
        cls(*good).set_param("nope", 1.0)


@pytest.mark.parametrize("cls,params", [(GaussianKernel, (1.0, 0.5)), (PeriodicKernel, (1.0, 0.5, 2.0))])
def test_kernel_copy_pickle(cls, params):
    k1 = cls(*params)
    for k2 in (k1.copy(), copy(k1), deepcopy(k1), pickle.loads(pickle.dumps(k1))):
        assert k2 is not k1 and type(k2) is cls
        assert (k2.params == k1.params).all()
        for n in cls._names:
            assert getattr(k1, n) is not getattr(k2, n)
            assert type(getattr(k2, n)) is np.float64


def test_kernel_api_surface():
    g, p = GaussianKernel(1, 1), PeriodicKernel(1, 1, 1)
    for n in ("K", "jacobian", "hessian", "dK_dh", "dK_dw", "d2K_dhdh", "d2K_dhdw", "d2K_dwdh", "d2K_dwdw"):
        assert callable(getattr(g, n))
    for a in "hwp":
        assert callable(getattr(p, "dK_d%s" % a))
        for b in "hwp":
            assert callable(getattr(p, "d2K_d%sd%s" % (a, b)))
    import sympy
    assert isinstance(g.sym_K, sympy.Expr) and isinstance(p.sym_K, sympy.Expr)
    assert set(gpb.__all__) >= {"ext", "GP", "Kernel", "PeriodicKernel", "GaussianKernel"}
    assert set(gpb.ext.__all__) == {"gaussian_c", "periodic_c", "gp_c"}


def test_gp_inputs_and_errors():
    gp = make_gp()
    assert gp.x.dtype == np.float64 and not gp.x.flags.writeable and not gp.y.flags.writeable
    assert type(gp.s) is np.float64
    assert (gp.params == np.array([1, 1, 1.0])).all()
    with pytest.raises(ValueError):
        gp.y = gp.y.copy()[:, None]                        # test_gp.py:282-286
    x, y = make_xy()
    with pytest.raises(ValueError):
        GP(GaussianKernel(1, 1), x, y, s=-1)               # test_gp.py:289-295
    gp = make_gp()
    gp.set_param("h", 1)
    gp.set_param("w", 0.2)
    gp.set_param("s", 0.01)
    assert gp.get_param("w") == 0.2 and gp.get_param("s") == 0.01
    with pytest.raises(AttributeError):
        gp.set_param("p", 1.1)                             # test_gp.py:425-432


def test_memo_bookkeeping():
    """del removes one key; every setter empties the cache (test_gp.py:245-279)."""
    gp = make_gp()
    for prop in ("Kxx", "Kxx_J", "Kxx_H", "Lxx", "inv_Kxx", "inv_Kxx_y", "log_lh", "lh",
                 "dloglh_dtheta", "dlh_dtheta", "d2lh_dtheta2"):
        assert isinstance(getattr(GP, prop), gpb.gp.memoprop)
        gp._memoized[prop] = "cached"
        assert getattr(gp, prop) == "cached"
        delattr(gp, prop)
        assert prop not in gp._memoized
        with pytest.raises(AttributeError):
            setattr(gp, prop, 1)
    for prop, val in (("x", gp.x.copy() + 1), ("y", gp.y.copy() + 1), ("s", gp.s + 1), ("params", gp.params + 1)):
        gp._memoized["Kxx"] = 1
        setattr(gp, prop, val)
        assert gp._memoized == {}
    gp._memoized["Kxx"] = 1
    gp.s = gp.s                                            # equal value: no reset
    gp.x = gp.x.copy()
    assert gp._memoized == {"Kxx": 1}
    gp.set_param("w", 3.0)
    assert gp._memoized == {}


def test_gp_copy_semantics():
    gp1 = make_gp()
    gp1._memoized["log_lh"] = np.float64(-1.0)
    for gp2 in (gp1.copy(deep=False), copy(gp1)):          # test_gp.py:351-370
        assert gp1 is not gp2 and gp1._x is gp2._x and gp1._y is gp2._y
        assert gp1._s is gp2._s and gp1.K is gp2.K
    for gp2 in (gp1.copy(deep=True), deepcopy(gp1), pickle.loads(pickle.dumps(gp1))):   # :373-422
        assert gp1 is not gp2 and gp1._x is not gp2._x and gp1._y is not gp2._y
        assert gp1._s is not gp2._s and gp1.K is not gp2.K and gp1.K.params is not gp2.K.params
        assert (gp1._x == gp2._x).all() and (gp1._y == gp2._y).all() and gp1._s == gp2._s
        assert (gp1.K.params == gp2.K.params).all()
        assert gp2._memoized == {"log_lh": -1.0}           # the memo is part of the state (gp.py:78-92)


def test_ext_argument_validation():
    """Cython-style buffer validation raises before any device work."""
    from gaussian_processes_b200.ext import gaussian_c, periodic_c, gp_c
    x = np.linspace(0, 1, 4)
    with pytest.raises(ValueError):
        gaussian_c.K(np.empty((4, 4), dtype=np.float32), x, x, 1.0, 1.0)
    with pytest.raises(ValueError):
        gaussian_c.K(np.empty((4, 4)), x.astype(np.float32), x, 1.0, 1.0)
    with pytest.raises(ValueError):
        gaussian_c.jacobian(np.empty((4, 4)), x, x, 1.0, 1.0)
    with pytest.raises(ValueError):
        periodic_c.hessian(np.empty((3, 3, 4, 5)), x, x, 1.0, 1.0, 1.0)
    with pytest.raises(ValueError):
        gaussian_c.K(np.empty((8, 4))[::2], x, x, 1.0, 1.0)
    with pytest.raises(ValueError):
        gp_c.log_lh(x, np.eye(4, dtype=np.float32), x)
    with pytest.raises(TypeError):
        gaussian_c.K([[0.0]], x, x, 1.0, 1.0)


def test_sharding_helpers():
    for total in (0, 1, 7, 4096):
        for world in (1, 2, 3, 8):
            b = [shard_bounds(total, world, r) for r in range(world)]
            assert b[0][0] == 0 and b[-1][1] == total
            assert all(b[i][1] == b[i + 1][0] for i in range(world - 1))
            sizes = [hi - lo for lo, hi in b]
            assert max(sizes) - min(sizes) <= 1
    assert select_best(np.array([-3.0, np.nan, -1.0, -1.0, -np.inf])) == 2
    assert select_best(np.array([np.nan, -np.inf])) in (0, 1)
    assert select_best(np.array([-np.inf, -np.inf])) == 0


def test_install_as_gp():
    import sys
    saved = {k: sys.modules.pop(k) for k in list(sys.modules) if k == "gp" or k.startswith("gp.")}
    try:
        gpb.install_as_gp()
        import gp
        assert gp.GP is GP and gp.ext.gaussian_c is gpb.ext.gaussian_c
        from gp.kernels import GaussianKernel as G2
        assert G2 is GaussianKernel
    finally:
        for k in [k for k in sys.modules if k == "gp" or k.startswith("gp.")]:
            del sys.modules[k]
        sys.modules.update(saved)


def test_refine_batched_bfgs_on_analytic_surface():
    """mlii.refine: every restart converges to the maximiser of a concave (in log space) surface,
    infeasible trials (log_lh = -inf, NaN gradient) only shorten the step."""
    from gaussian_processes_b200 import mlii
    opt = np.array([1.3, 0.4, 0.9])
    A = np.array([[2, 0.3, 0], [0.3, 1, 0.2], [0, 0.2, 3.0]])
    ncalls = []

    def ev(th):
        ncalls.append(len(th))
        u = np.log(th) - np.log(opt)
        f = -0.5 * np.einsum("bi,ij,bj->b", u, A, u) - 0.1 * np.sum(u ** 4, axis=1)
        gu = -(u @ A) - 0.4 * u ** 3
        g = gu / th
        bad = th[:, 0] > 1.9                      # a "not positive definite" region
        f = np.where(bad, -np.inf, f)
        g[bad] = np.nan
        return f, g
    starts = np.array([[1.0, 1.0, 1.0], [1.8, 0.2, 0.5], [0.7, 0.9, 2.0]])
    th, f, g, n = mlii.refine(starts, ev, steps=40, gtol=1e-9)
    assert np.allclose(th, opt, rtol=1e-6)
    assert np.all(f > -1e-12) and n == sum(ncalls)
    assert max(ncalls) <= 3                      # one trial per restart per round
    with pytest.raises(ValueError):
        mlii.refine(np.array([[1.0, -1.0, 1.0]]), ev)
    th0, f0, g0, n0 = mlii.refine(np.empty((0, 3)), ev)
    assert th0.shape == (0, 3) and n0 == 0


# ------------------------------------------------------------------ generated functors
def test_symbolic_kernel_codegen_compiles_for_sm100a():
    """sym_K -> CUDA source (K, Jacobian, Hessian with shared sub-expressions) -> NVRTC cubin for
    sm_100a.  Compilation needs no GPU; running the builder does (tests/test_parity_gpu.py)."""
    pytest.importorskip("sympy")
    pytest.importorskip("cuda.bindings.nvrtc")
    import sympy as sym
    from gaussian_processes_b200 import GaussianKernel, PeriodicKernel, SymbolicKernel
    from gaussian_processes_b200.kernels.symbolic import _Module
    for k0 in (GaussianKernel(1.0, 0.5), PeriodicKernel(1.0, 0.5, 1.3)):
        k = SymbolicKernel.from_kernel(k0)
        assert (k.params == k0.params).all() and k._names == k0._names
        assert "symk_eval2" in k.cuda_source and k.cuda_source.count("exp(") <= 3     # one exp per level
        assert len(_Module.compile(k.cuda_source)) > 1000
        assert (k.copy().params == k.params).all() and k.copy() is not k
    h, w, a, d = sym.symbols("h w a d")
    rq = SymbolicKernel(h ** 2 * (1 + d ** 2 / (2 * a * w ** 2)) ** (-a), ("h", "w", "a"), (1.0, 0.5, 2.0))
    assert len(_Module.compile(rq.cuda_source)) > 1000
    with pytest.raises(ValueError):
        SymbolicKernel(h * sym.Symbol("zz") * d, ("h",), (1.0,))          # unknown symbol
    with pytest.raises(ValueError):
        rq.set_param("w", 0.0)                                             # same validation as the built-ins


# ------------------------------------------------------------------ round-2 host-logic fixes
def test_data_generation_counter_not_object_identity():
    """The batched evaluator re-uploads x / y when the GP's data generation moved.  Object ids cannot be
    the key: CPython recycles them (assign gp.y twice and the third array usually gets the first one's id)."""
    x, y = make_xy()
    gp = GP(GaussianKernel(1, 1), x, y, s=1)
    g0 = gp._data_gen
    ids = {id(gp._y)}
    gp.y = y + 1.0
    g1 = gp._data_gen
    gp.y = y + 2.0
    g2 = gp._data_gen
    assert g0 < g1 < g2
    gp.x = x + 0.5
    assert gp._data_gen > g2
    gen = gp._data_gen
    gp.s = 0.5                                   # hyperparameters do not touch the observations
    gp.set_param("w", 0.7)
    gp.y = gp.y.copy()                           # equal values: the setter is a no-op (gp.py:166-167)
    assert gp._data_gen == gen
    # a copy starts its own evaluator and its own generation count
    for g in (gp.copy(), copy(gp), pickle.loads(pickle.dumps(gp))):
        assert not hasattr(g, "_batch_ev") and g._data_gen == 0


def test_fused_kind_only_for_builtin_formulas():
    from gaussian_processes_b200.gp import fused_kind

    class Scaled(GaussianKernel):                # overrides the element formula: must not take the CUDA functor
        def K(self, x1, x2, out=None):
            return 2.0 * GaussianKernel.K(self, x1, x2, out)

    class Renamed(PeriodicKernel):               # same formulas, extra behaviour elsewhere: fused path is right
        def describe(self):
            return "periodic"

    assert fused_kind(GaussianKernel(1, 1)) == 0 and fused_kind(PeriodicKernel(1, 1, 1)) == 1
    assert fused_kind(Scaled(1, 1)) is None
    assert fused_kind(Renamed(1, 1, 1)) == 1
    assert fused_kind(object()) is None


def test_fit_mlii_validates_and_reports_nothing_found():
    import torch
    from gaussian_processes_b200 import mlii
    x, y = make_xy()
    gp = GP(GaussianKernel(1, 1), x, y, s=1)
    p0 = gp.params.copy()

    def all_bad(th):
        t = np.full((len(th), 5), np.nan)
        t[:, 0] = -np.inf
        t[:, 4] = 1
        return torch.from_numpy(t)
    cand = np.array([[1.0, 0.5, 0.1], [2.0, 0.3, 0.2]])
    res = mlii.fit_MLII(gp, cand, evaluate=all_bad, distributed=False)
    assert not res.found and res.best_index == -1 and res.best_params is None
    assert res.best_log_lh == -np.inf and (gp.params == p0).all()          # the GP did not move
    for bad in ([[0.0, 0.5, 0.1]], [[1.0, -1.0, 0.1]], [[1.0, 0.5, -0.1]], [[1.0, np.nan, 0.1]]):
        with pytest.raises(ValueError):
            mlii.fit_MLII(gp, np.array(bad), evaluate=all_bad, distributed=False)
    with pytest.raises(ValueError):
        mlii.fit_MLII(gp, np.empty((0, 3)), evaluate=all_bad, distributed=False)
    with pytest.raises(ValueError):
        mlii.fit_MLII(gp, np.ones((2, 4)), evaluate=all_bad, distributed=False)
