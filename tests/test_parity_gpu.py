"""Parity of the CUDA path (through the public API and the C ABI) against the CPU
oracle on the same seeded inputs, against the golden vectors produced by the
unmodified reference, and -- at BASELINE.json's full sizes -- through size-independent
properties.  Tolerance (BASELINE.json north_star): scalars |d|/|ref| <= 1e-9,
arrays ||d||_inf <= 1e-9 ||ref||_inf; identical -inf / 0 / NaN / exception behaviour."""
import numpy as np
import pytest

import gaussian_processes_b200 as gpb
from gaussian_processes_b200 import GP, GaussianKernel, PeriodicKernel
from conftest import golden, assert_parity, synth_xy, RTOL

pytestmark = pytest.mark.gpu


def make_kernel(params):
    return GaussianKernel(*params) if len(params) == 2 else PeriodicKernel(*params)


def kind_of(oracle, kparams):
    return oracle.GAUSSIAN if len(kparams) == 2 else oracle.PERIODIC


# ------------------------------------------------------------------ builders
@pytest.mark.parametrize("n1,n2", [(10, 10), (23, 16), (1, 7), (7, 1), (130, 67), (257, 300)])
@pytest.mark.parametrize("kparams", [(1.3, 0.4), (0.7, 0.9, 1.7)])
def test_builders_vs_oracle(oracle, n1, n2, kparams):
    rng = np.random.RandomState(n1 * 1000 + n2)
    x1, x2 = rng.uniform(-6, 6, n1), rng.uniform(-6, 6, n2)
    k, kind = make_kernel(kparams), kind_of(oracle, kparams)
    assert_parity(k(x1, x2), oracle.K(kind, x1, x2, kparams), 1e-13, "K")
    assert_parity(k.jacobian(x1, x2), oracle.jacobian(kind, x1, x2, kparams), 1e-13, "J")
    assert_parity(k.hessian(x1, x2), oracle.hessian(kind, x1, x2, kparams), 1e-12, "H")


def test_builders_empty_inputs():
    k = GaussianKernel(1.0, 1.0)
    e, x = np.empty(0), np.linspace(0, 1, 3)
    assert k(e, x).shape == (0, 3) and k(x, e).shape == (3, 0) and k.jacobian(e, e).shape == (2, 0, 0)


def test_builders_vs_golden():
    g = golden("kernels")
    for t in range(4):
        for tag, xa in (("g", g["x10"]), ("p", g["x16"])):
            k = make_kernel(g["%s%d_params" % (tag, t)])
            assert_parity(k(xa, xa), g["%s%d_K" % (tag, t)], 1e-13)
            assert_parity(k.jacobian(xa, xa), g["%s%d_J" % (tag, t)], 1e-13)
            assert_parity(k.hessian(xa, xa), g["%s%d_H" % (tag, t)], 1e-12)
            assert_parity(k.hessian(g["xr"], xa), g["%s%d_Hr" % (tag, t)], 1e-12)
    kz = GaussianKernel(*g["gz_params"])
    K = kz(g["gz_x"], g["gz_x"])
    assert ((K == 0) == (g["gz_K"] == 0)).all() and (K == 0).sum() > 0     # e < MIN -> exactly 0
    assert ((kz.hessian(g["gz_x"], g["gz_x"]) == 0) == (g["gz_H"] == 0)).all()


# ------------------------------------------------------------------ GP vs golden (reference outputs)
FULL_KEYS = ["Kxx", "Kxx_J", "Kxx_H", "Lxx", "inv_Kxx", "inv_Kxx_y", "log_lh", "dloglh_dtheta",
             "dlh_dtheta", "d2lh_dtheta2"]


@pytest.mark.parametrize("name", ["gp_suite_g0", "gp_suite_g1", "gp_suite_g2", "gp_suite_p0",
                                  "gp_suite_p1", "gp_suite_p2", "gp_c1"])
def test_gp_vs_golden_small(name):
    g = golden(name)
    gp = GP(make_kernel(g["params"][:-1]), g["x"], g["y"], s=g["params"][-1])
    cond = np.linalg.cond(g["Kxx"])
    assert cond < 1e4                       # fixtures are well conditioned: the 1e-9 bar applies as is
    for key in FULL_KEYS:
        assert_parity(getattr(gp, key), g[key], RTOL, "%s/%s" % (name, key))
    for key in ("mean", "cov", "dm_dtheta"):
        assert_parity(getattr(gp, key)(g["xo"]), g[key], RTOL, "%s/%s" % (name, key))
    assert_parity(gp.d2loglh_normalised(), g["d2lh_norm"], RTOL, "d2lh_norm")
    assert float(gp.lh) == pytest.approx(float(g["lh"]), rel=1e-9)


@pytest.mark.parametrize("name", ["gp_g300", "gp_p257"])
def test_gp_vs_golden_multiblock(name):
    g = golden(name)
    gp = GP(make_kernel(g["params"][:-1]), g["x"], g["y"], s=g["params"][-1])
    assert_parity(gp.log_lh, g["log_lh"])
    assert_parity(gp.dloglh_dtheta, g["dloglh_dtheta"])
    assert_parity(gp.inv_Kxx_y, g["inv_Kxx_y"])
    assert_parity(gp.mean(g["xo"]), g["mean"])
    c = gp.cov(g["xo"])
    assert np.max(np.abs(np.diag(c) - g["cov_diag"])) <= RTOL * float(g["cov_fro"])
    assert_parity(c[0], g["cov_row0"], 1e-9 * float(g["cov_fro"]) / np.max(np.abs(g["cov_row0"])))
    assert_parity(gp.dm_dtheta(g["xo"]), g["dm_dtheta"])
    assert_parity(gp.d2loglh_normalised(), g["d2lh_norm"])
    assert_parity(np.diag(gp.Lxx), g["Lxx_diag"])
    assert_parity(np.diag(gp.inv_Kxx), g["inv_Kxx_diag"])


@pytest.mark.parametrize("n", [1024, 4096])
def test_c2_vs_reference_values(n):
    """BASELINE config C2 (N=4096) and its N=1024 sibling against the reference's own numbers."""
    g = golden("gp_c2_n%d" % n)
    x, y = synth_xy(n, 0)
    gp = GP(GaussianKernel(*g["params"][:-1]), x, y, s=g["params"][-1])
    assert_parity(gp.log_lh, g["log_lh"])
    assert_parity(gp.dloglh_dtheta, g["dloglh_dtheta"])
    assert gp.lh == 0 and type(gp.lh) is int                       # underflow -> int 0 (gp.py:393-394)
    assert (gp.dlh_dtheta == 0).all() and (gp.d2lh_dtheta2 == 0).all()
    assert_parity(gp.inv_Kxx_y, g["inv_Kxx_y"])
    assert_parity(gp.mean(g["xo"]), g["mean"])
    assert_parity(gp.cov(g["xo"]), g["cov"])
    assert_parity(gp.dm_dtheta(g["xo"]), g["dm_dtheta"])
    assert_parity(gp.d2loglh_normalised(), g["d2lh_norm"])
    assert_parity(np.diag(gp.Lxx), g["Lxx_diag"])
    assert_parity(np.diag(gp.inv_Kxx), g["inv_Kxx_diag"])


# ------------------------------------------------------------------ GP vs oracle on seeded inputs
@pytest.mark.parametrize("n,kparams,s,m", [
    (1, (1.0, 0.5), 1.0, 3), (2, (1.0, 0.5), 0.5, 1), (127, (1.2, 0.3), 0.7, 50), (128, (1.2, 0.3), 0.7, 128),
    (129, (0.8, 0.6), 0.9, 130), (640, (1.0, 0.5), 1.0, 200), (513, (1.0, 1.0, 1.0), 1.0, 77)])
def test_gp_vs_oracle(oracle, n, kparams, s, m):
    x, y = synth_xy(n, n)
    xo = np.linspace(-2 * np.pi, 2 * np.pi, m)
    gp = GP(make_kernel(kparams), x, y, s=s)
    o = oracle.OracleGP(kind_of(oracle, kparams), kparams, x, y, s)
    for key in ("Kxx", "Lxx", "inv_Kxx", "inv_Kxx_y", "log_lh", "dloglh_dtheta"):
        assert_parity(getattr(gp, key), getattr(o, key), RTOL, key)
    assert_parity(gp.mean(xo), o.mean(xo), RTOL, "mean")
    assert_parity(gp.cov(xo), o.cov(xo), RTOL, "cov")
    assert_parity(gp.dm_dtheta(xo), o.dm_dtheta(xo), RTOL, "dm")
    assert_parity(gp.d2loglh_normalised(), o.d2lh_dtheta2_with(1.0, o.dloglh_dtheta), RTOL, "d2lh(lh=1)")
    assert gp.mean(np.empty(0)).shape == (0,) and gp.cov(np.empty(0)).shape == (0, 0)
    assert gp.dm_dtheta(np.empty(0)).shape == (len(kparams) + 1, 0)


@pytest.mark.parametrize("n", [3, 31, 32, 33, 50, 64, 65, 96, 97, 100])
@pytest.mark.parametrize("kparams,s", [((1.1, 0.35), 0.6), ((0.9, 0.8, 1.4), 0.8)])
def test_one_block_gp_vs_oracle(oracle, n, kparams, s):
    """N <= 128: factor + invert in one CTA (identity sub-blocks skipped), then solves, log_lh,
    K^-1 and the gradient brackets in ONE more launch (small.cu); every sub-block count and both
    sides of each 32-boundary.  The gradient is asked for first (cold: one library call)."""
    x, y = synth_xy(n, 100 + n)
    xo = np.linspace(-6, 6, 37)
    gp = GP(make_kernel(kparams), x, y, s=s)
    o = oracle.OracleGP(kind_of(oracle, kparams), kparams, x, y, s)
    assert_parity(gp.dloglh_dtheta, o.dloglh_dtheta, RTOL, "dloglh (cold)")
    for key in ("log_lh", "inv_Kxx_y", "inv_Kxx", "Lxx", "dlh_dtheta"):
        assert_parity(getattr(gp, key), getattr(o, key), RTOL, key)
    assert_parity(gp.mean(xo), o.mean(xo), RTOL, "mean")
    c37 = gp.cov(xo)                                   # one-launch covariance (small.cu)
    assert_parity(c37, o.cov(xo), RTOL, "cov") and np.array_equal(c37, c37.T)
    for m2 in (1, 32, 100, 128, 129):                  # every row-block count, and past the one-launch limit
        xo2 = np.linspace(-5.5, 6.1, m2)
        c2 = gp.cov(xo2)
        assert_parity(c2, o.cov(xo2), RTOL, "cov m=%d" % m2)
        assert np.array_equal(c2, c2.T)
    assert_parity(gp.var(xo), np.diag(o.cov(xo)), RTOL, "var")
    assert_parity(gp.dm_dtheta(xo), o.dm_dtheta(xo), RTOL, "dm")
    assert_parity(gp.d2loglh_normalised(), o.d2lh_dtheta2_with(1.0, o.dloglh_dtheta), RTOL, "d2lh(lh=1)")
    # same candidate through the batched evaluator (device KParams, batch > 1)
    th = np.array(list(kparams) + [s])
    llh, grad = gp.batch_eval(np.stack([th, th * 1.01, th]))
    assert_parity(llh[0], o.log_lh) and assert_parity(grad[0], o.dloglh_dtheta)
    assert llh[2] == llh[0] and np.array_equal(grad[2], grad[0])
    # new hyperparameters on the resident engine
    gp.set_param("w", kparams[1] * 1.3)
    o2 = oracle.OracleGP(kind_of(oracle, kparams), (kparams[0], kparams[1] * 1.3) + tuple(kparams[2:]), x, y, s)
    assert_parity(gp.log_lh, o2.log_lh) and assert_parity(gp.dloglh_dtheta, o2.dloglh_dtheta)


def test_one_block_not_pd_and_clamp(oracle):
    from suite_util import INVALID_X, INVALID_Y, INVALID_H, INVALID_W
    gp = GP(GaussianKernel(INVALID_H, INVALID_W), INVALID_X, INVALID_Y, s=0)
    assert np.isnan(gp.dloglh_dtheta).all() and gp.log_lh == -np.inf and gp.lh == 0
    with pytest.raises(np.linalg.LinAlgError):
        gp.inv_Kxx
    # logdet < MIN with a successful Cholesky: -inf, finite gradient (gp_c.pyx:22-23)
    x, y = synth_xy(120, 3)
    g2 = GP(GaussianKernel(1.0, 1.0), x, y, s=1e-2)           # cond(Kxx) = 1e5
    o = oracle.OracleGP(oracle.GAUSSIAN, (1.0, 1.0), x, y, 1e-2)
    assert o.log_lh == -np.inf and g2.log_lh == -np.inf
    assert_parity(g2.dloglh_dtheta, o.dloglh_dtheta, 1e-7, "dloglh at s=1e-2")


def test_logdet_clamp_matches_reference(oracle):
    """logdet(Kxx) < MIN -> log_lh = -inf although the Cholesky succeeds; the
    gradient is still finite (only LinAlgError gives NaN)."""
    x, y = synth_xy(1024, 0)
    gp = GP(GaussianKernel(1.0, 0.5), x, y, s=0.1)
    o = oracle.OracleGP(oracle.GAUSSIAN, (1.0, 0.5), x, y, 0.1)
    assert o.log_lh == -np.inf and gp.log_lh == -np.inf and gp.lh == 0
    assert np.isfinite(gp.dloglh_dtheta).all()
    assert_parity(gp.dloglh_dtheta, o.dloglh_dtheta, 1e-7, "dloglh at cond ~1e4")


def test_nonfinite_inputs_raise_like_scipy():
    x, y = synth_xy(16, 0)
    y2 = y.copy()
    y2[3] = np.nan
    gp = GP(GaussianKernel(1.0, 0.5), x, y2, s=1.0)
    with pytest.raises(ValueError):
        gp.Lxx


# ------------------------------------------------------------------ C ABI level
def test_host_abi_eval_matches_object_path():
    import ctypes
    from gaussian_processes_b200 import _lib
    x, y = synth_xy(300, 0)
    th = np.array([[1.0, 0.5, 1.0], [1.3, 0.4, 0.8]])
    res = np.empty((2, 8))
    _lib.call("gpb_gp_eval_host", 0, th.ctypes.data_as(_lib.dp), 2, x.ctypes.data_as(_lib.dp),
              y.ctypes.data_as(_lib.dp), 300, 1, res.ctypes.data_as(_lib.dp))
    for b in range(2):
        gp = GP(GaussianKernel(*th[b, :2]), x, y, s=th[b, 2])
        assert_parity(res[b, 0], gp.log_lh, 1e-13)
        assert_parity(res[b, 1:4], gp.dloglh_dtheta, 1e-12)
        assert res[b, 7] == 0


def test_ext_gp_c_dropins():
    from gaussian_processes_b200.ext import gp_c
    g = golden("gp_suite_g0")
    y, s = g["y"], float(g["params"][-1])
    assert_parity(gp_c.log_lh(y, g["Kxx"], g["inv_Kxx_y"]), g["log_lh"])
    out = np.empty(3)
    gp_c.dloglh_dtheta(y, g["inv_Kxx"], g["Kxx_J"], g["inv_Kxx_y"], s, out)
    assert_parity(out, g["dloglh_dtheta"])
    gp_c.dlh_dtheta(y, g["inv_Kxx"], g["Kxx_J"], g["inv_Kxx_y"], s, float(g["lh"]), out)
    assert_parity(out, g["dlh_dtheta"])
    o2 = np.empty((3, 3))
    gp_c.d2lh_dtheta2(y, g["inv_Kxx"], g["Kxx_J"], g["Kxx_H"], g["inv_Kxx_y"], s, float(g["lh"]),
                      g["dlh_dtheta"], o2)
    assert_parity(o2, g["d2lh_dtheta2"])
    k = GaussianKernel(*g["params"][:-1])
    dm = np.empty((3, g["xo"].size))
    gp_c.dm_dtheta(y, g["inv_Kxx"], g["Kxx_J"], k.jacobian(g["xo"], g["x"]), k(g["xo"], g["x"]), s, dm)
    assert_parity(dm, g["dm_dtheta"])
    gp = golden("gp_suite_p1")
    assert_parity(gp_c.log_lh(gp["y"], gp["Kxx"], gp["inv_Kxx_y"]), gp["log_lh"])


# ------------------------------------------------------------------ batched search
def test_batch_eval_and_fit_mlii_vs_oracle(oracle):
    x, y = synth_xy(200, 7)
    rng = np.random.RandomState(11)
    B = 24
    cand = np.stack([rng.uniform(0.5, 2, B), rng.uniform(np.pi / 32, np.pi / 2, B), rng.uniform(0.75, 1.5, B)], axis=1)
    gp = GP(GaussianKernel(1.0, 1.0), x, y, s=1.0)
    llh, grad = gp.batch_eval(cand)
    best, ollh, ograd = oracle.oracle_fit_mlii(oracle.GAUSSIAN, x, y, cand)
    assert_parity(llh, ollh)
    for b in range(B):
        assert_parity(grad[b], ograd[b], RTOL, "grad[%d]" % b)
    res = gp.fit_MLII(cand)
    assert res.best_index == best and np.array_equal(gp.params, cand[best])
    assert_parity(gp.log_lh, ollh[best])
    kp = np.stack([rng.uniform(0.5, 2, 5), rng.uniform(0.5, 1.5, 5), rng.uniform(0.5, 3, 5), rng.uniform(0.75, 1.5, 5)], axis=1)
    gpp = GP(PeriodicKernel(1.0, 1.0, 1.0), x, y, s=1.0)
    l2, g2 = gpp.batch_eval(kp)
    _, ol2, og2 = oracle.oracle_fit_mlii(oracle.PERIODIC, x, y, kp)
    assert_parity(l2, ol2)
    assert_parity(g2, og2)


def test_batch_eval_failure_rows(oracle):
    """A candidate whose Kxx is not PD gives -inf / NaN in its row only; the clamp row gives
    -inf with a finite gradient (reference semantics, gp.py:362-365, 424-428)."""
    from suite_util import INVALID_X, INVALID_Y, INVALID_H, INVALID_W
    gp = GP(GaussianKernel(1.0, 1.0), INVALID_X, INVALID_Y, s=0.5)
    cand = np.array([[1.0, 0.5, 0.5], [INVALID_H, INVALID_W, 0.0], [0.7, 0.3, 0.2]])
    llh, grad = gp.batch_eval(cand)
    assert llh[1] == -np.inf and np.isnan(grad[1]).all()
    for b in (0, 2):
        o = oracle.OracleGP(oracle.GAUSSIAN, cand[b, :2], INVALID_X, INVALID_Y, cand[b, 2])
        assert_parity(llh[b], o.log_lh)
        assert_parity(grad[b], o.dloglh_dtheta)


# ------------------------------------------------------------------ full-size properties
def test_full_size_properties_n4096():
    """At the headline size the oracle takes minutes; use size-independent identities:
    K * Ki = I, L L^T = K, K alpha = y, symmetric Ki, and the batched evaluator agreeing
    with the object path."""
    n = 4096
    x, y = synth_xy(n, 0)
    gp = GP(GaussianKernel(1.1, 0.45), x, y, s=1.0)    # s >= 0.92 keeps logdet above MIN at N=4096
    assert np.isfinite(gp.log_lh)
    K, L, Ki, a = gp.Kxx, gp.Lxx, gp.inv_Kxx, gp.inv_Kxx_y
    assert np.array_equal(L, np.tril(L))
    scale = np.max(np.abs(K))
    assert np.max(np.abs(L @ L.T - K)) <= 1e-12 * scale
    assert np.max(np.abs(K @ Ki - np.eye(n))) <= 1e-10
    assert np.max(np.abs(Ki - Ki.T)) <= 1e-13 * np.max(np.abs(Ki))
    assert np.max(np.abs(K @ a - y)) <= 1e-11 * np.max(np.abs(y))
    llh, grad = gp.batch_eval(np.array([gp.params, gp.params]))
    assert_parity(llh[0], gp.log_lh, 1e-13)
    assert_parity(grad[0], gp.dloglh_dtheta, 1e-12)
    assert llh[0] == llh[1] and np.array_equal(grad[0], grad[1])     # deterministic reductions
    # gradient against a central difference of the object's own log_lh (test_gp.py:75-98 at scale)
    for i, name in enumerate(("h", "w", "s")):
        g0, g1 = gp.copy(), gp.copy()
        g0.set_param(name, gp.params[i] - 1e-5)
        g1.set_param(name, gp.params[i] + 1e-5)
        fd = (g1.log_lh - g0.log_lh) / 2e-5
        assert abs(fd - gp.dloglh_dtheta[i]) <= 1e-5 * abs(fd) + 1e-4


def test_periodic_posterior_sharded_blocks(oracle):
    """C3-shaped (scaled to what the oracle finishes in seconds): Periodic GP, the posterior of a
    block of test points equals the same rows/cols of the full posterior (test points shard)."""
    n, m = 1536, 600
    x, y = synth_xy(n, 2)
    xo = np.linspace(-2 * np.pi, 2 * np.pi, m)
    gp = GP(PeriodicKernel(1.0, 1.0, 1.0), x, y, s=1.0)
    o = oracle.OracleGP(oracle.PERIODIC, (1.0, 1.0, 1.0), x, y, 1.0)
    mean, cov = gp.mean(xo), gp.cov(xo)
    assert_parity(mean, o.mean(xo))
    assert_parity(cov, o.cov(xo))
    lo, hi = 150, 420
    assert_parity(gp.mean(xo[lo:hi]), mean[lo:hi], 1e-13)
    assert_parity(gp.cov(xo[lo:hi]), cov[lo:hi, lo:hi], 1e-12)
    assert_parity(gp.Kxx_J, o.Kxx_J, 1e-13)


# ------------------------------------------------------------------ newer surface
def test_cov_rows_matches_cov(oracle):
    """Row blocks of the predictive covariance (the unit of test-point sharding) against the
    oracle's full covariance, including ragged block boundaries."""
    x, y = synth_xy(300, 5)
    xo = np.linspace(-6, 6, 211)
    gp = GP(GaussianKernel(1.0, 0.5), x, y, s=1.0)
    ref = oracle.OracleGP(oracle.GAUSSIAN, (1.0, 0.5), x, y, 1.0).cov(xo)
    for lo, hi in ((0, 211), (0, 1), (17, 150), (130, 211)):
        assert_parity(gp.cov_rows(xo, lo, hi), ref[lo:hi], RTOL, "rows %d:%d" % (lo, hi))
    assert gp.cov_rows(xo, 5, 5).shape == (0, 211)


def test_batch_eval_chunking_and_knobs():
    """Chunked evaluation (workspace smaller than the batch) and every tuning knob give the same
    numbers: the knobs change scheduling, never arithmetic order within a candidate."""
    from gaussian_processes_b200 import _lib, engine
    x, y = synth_xy(700, 9)
    rng = np.random.RandomState(2)
    th = np.stack([rng.uniform(0.5, 2, 11), rng.uniform(0.2, 1.0, 11), rng.uniform(0.75, 1.5, 11)], axis=1)
    base = engine.BatchEvaluator(engine.GAUSSIAN, x, y).eval(th)
    small = engine.BatchEvaluator(engine.GAUSSIAN, x, y, max_batch=3).eval(th)
    assert np.array_equal(base[0], small[0]) and np.array_equal(base[1], small[1])
    try:
        for name, vals in (("eval_streams", (1, 2)), ("gemm_bm", (128, 64)), ("potrf_inner", (1, 2))):
            for v in vals:
                _lib.set_option(name, v)
                got = engine.BatchEvaluator(engine.GAUSSIAN, x, y).eval(th)
                assert_parity(got[0], base[0], 1e-13, "%s=%d llh" % (name, v))
                assert_parity(got[1], base[1], 1e-11, "%s=%d grad" % (name, v))
                _lib.set_option(name, 0)
    finally:
        for name in ("eval_streams", "gemm_bm", "potrf_inner"):
            _lib.set_option(name, 0)
    with pytest.raises(_lib.GpbError):
        _lib.set_option("no_such_knob", 1)


def test_clamp_candidates_in_batch(oracle):
    """Reference-suite noise range s ~ U(0, 0.5) (tests/util.py:24-25) at N=1024: most candidates hit
    the logdet < MIN clamp; batched rows must reproduce -inf with finite gradients."""
    x, y = synth_xy(1024, 0)
    rng = np.random.RandomState(8)
    th = np.stack([rng.uniform(0.5, 2, 6), rng.uniform(0.3, 1.0, 6), rng.uniform(0.05, 0.5, 6)], axis=1)
    gp = GP(GaussianKernel(1.0, 0.5), x, y, s=1.0)
    llh, grad = gp.batch_eval(th)
    for b in range(3):
        o = oracle.OracleGP(oracle.GAUSSIAN, th[b, :2], x, y, th[b, 2])
        assert (llh[b] == -np.inf) == (o.log_lh == -np.inf)
        if np.isfinite(o.log_lh):
            assert_parity(llh[b], o.log_lh, 1e-8)
    assert (llh == -np.inf).any() and np.isfinite(grad).all()


# ------------------------------------------------------------------ additive APIs of this round
@pytest.mark.parametrize("kparams,n,m", [((1.0, 0.5), 300, 257), ((1.0, 1.0, 1.3), 131, 64), ((0.8, 0.3), 64, 1)])
def test_var_is_cov_diagonal(oracle, kparams, n, m):
    """GP.var (fused row norms of K(xo,x) L^-T) against the oracle's explicit-inverse cov diagonal,
    and the mean fast path (two rows per warp, 128-bit loads) at odd sizes."""
    x, y = synth_xy(n, 21)
    xo = np.linspace(-7, 7, m)
    gp = GP(make_kernel(kparams), x, y, s=0.7)
    o = oracle.OracleGP(kind_of(oracle, kparams), kparams, x, y, 0.7)
    ocov = o.cov(xo)
    v = gp.var(xo)
    assert v.shape == (m,)
    assert np.max(np.abs(v - np.diag(ocov))) <= RTOL * np.max(np.abs(ocov))
    assert_parity(gp.mean(xo), o.mean(xo))
    assert gp.var(np.empty(0)).shape == (0,)


def test_fit_mlii_refine_reaches_scipy_optimum(oracle):
    """Batched BFGS polish: from the best grid candidates the refined point is stationary, not worse
    than any start, and agrees with scipy's optimum of the ORACLE's log_lh (same surface)."""
    from scipy.optimize import minimize
    x, y = synth_xy(96, 5)
    rng = np.random.RandomState(3)
    B = 32
    cand = np.stack([rng.uniform(0.5, 2, B), rng.uniform(0.2, 1.5, B), rng.uniform(0.05, 0.8, B)], axis=1)
    gp = GP(GaussianKernel(1.0, 1.0), x, y, s=1.0)
    res = gp.fit_MLII(cand, refine_top=4, refine_steps=40, refine_gtol=1e-7)
    assert res.refined is not None and res.refined["params"].shape == (4, 3)
    assert res.best_log_lh >= res.log_lh[res.best_index]
    assert np.array_equal(gp.params, res.best_params)
    b = res.refined["best"]
    assert np.max(np.abs(res.refined["dloglh_dtheta"][b] * res.refined["params"][b])) < 1e-5
    o = oracle.OracleGP(oracle.GAUSSIAN, res.best_params[:2], x, y, res.best_params[2])
    assert_parity(res.best_log_lh, o.log_lh)

    def neg(u):
        t = np.exp(u)
        og = oracle.OracleGP(oracle.GAUSSIAN, t[:2], x, y, t[2])
        return -float(og.log_lh), -(og.dloglh_dtheta * t)
    ref = minimize(neg, np.log(cand[res.best_index]), jac=True, method="L-BFGS-B", options=dict(gtol=1e-8, maxiter=200))
    assert res.best_log_lh >= -ref.fun - 1e-6 * abs(ref.fun)


def test_cov_rows_aligned_shards_use_symmetry(oracle):
    """Tile-aligned shards (what M = 16384 over 2/4/8 GPUs gives) take the symmetric diagonal-block
    path of cov_rows; stitched together they reproduce the oracle's covariance and ``cov``."""
    x, y = synth_xy(200, 9)
    m = 1024
    xo = np.linspace(-6, 6, m)
    gp = GP(PeriodicKernel(1.0, 1.0, 1.3), x, y, s=0.8)
    ref = oracle.OracleGP(oracle.PERIODIC, (1.0, 1.0, 1.3), x, y, 0.8).cov(xo)
    full = gp.cov(xo)
    assert_parity(full, ref)
    for G in (2, 4):
        parts = [gp.cov_rows(xo, r * m // G, (r + 1) * m // G) for r in range(G)]
        st = np.concatenate(parts, axis=0)
        assert_parity(st, ref, RTOL, "G=%d" % G)
        assert np.max(np.abs(st - st.T)) <= 1e-12 * np.max(np.abs(ref))


@pytest.mark.parametrize("n1,n2", [(3, 1100000), (3000, 33), (1500, 700)])
def test_builders_large_and_skinny_downloads(oracle, n1, n2):
    """Host-buffer builders whose results take the different staged-download shapes: a single row
    longer than a pinned slot (column split), many short rows (re-tiled as one contiguous run) and
    multi-chunk slices; bit-level agreement is not required, the 1e-9 bar is."""
    rng = np.random.RandomState(n1 + n2)
    x1, x2 = rng.uniform(-6, 6, n1), rng.uniform(-6, 6, n2)
    k = GaussianKernel(1.1, 0.7)
    assert_parity(k.K(x1, x2), oracle.K(oracle.GAUSSIAN, x1, x2, (1.1, 0.7)), RTOL, "K")
    assert_parity(k.jacobian(x1, x2), oracle.jacobian(oracle.GAUSSIAN, x1, x2, (1.1, 0.7)), RTOL, "jacobian")
    out = np.full((n1, n2), np.nan)
    assert k.dK_dw(x1, x2, out=out) is out and not np.isnan(out).any()


# ------------------------------------------------------------------ user-defined kernels
class NumpyGaussianKernel(gpb.Kernel):
    """A user's own Kernel subclass (no CUDA functor): the Gaussian formulas of
    gaussian_c.pyx:18-164 in numpy, so its GP must agree with the built-in kernel's."""
    _names = ("h", "w")

    def __init__(self, h, w):
        self.set_param("h", h)
        self.set_param("w", w)

    def _parts(self, x1, x2):
        d2 = (np.asarray(x1)[:, None] - np.asarray(x2)[None, :]) ** 2
        c = np.sqrt(2.0 / np.pi)
        return d2, c, np.exp(-0.5 * d2 / self.w ** 2)

    def K(self, x1, x2, out=None):
        d2, c, e = self._parts(x1, x2)
        return 0.5 * c * self.h ** 2 / self.w * e

    def jacobian(self, x1, x2, out=None):
        d2, c, e = self._parts(x1, x2)
        h, w = self.h, self.w
        return np.stack([c * h / w * e, e * (0.5 * c * h ** 2 / w ** 4 * d2 - 0.5 * c * h ** 2 / w ** 2)])

    def hessian(self, x1, x2, out=None):
        d2, c, e = self._parts(x1, x2)
        h, w = self.h, self.w
        hh = c / w * e
        hw = e * (c * h / w ** 4 * d2 - c * h / w ** 2)
        ww = e * (0.5 * c * h ** 2 / w ** 7 * d2 ** 2 - 2.5 * c * h ** 2 / w ** 5 * d2 + c * h ** 2 / w ** 3)
        return np.stack([np.stack([hh, hw]), np.stack([hw, ww])])


@pytest.mark.parametrize("n,m", [(40, 25), (300, 130)])
def test_user_defined_kernel_subclass(oracle, n, m):
    """GP over a Kernel subclass that only exists in Python (gp.py calls K / jacobian / hessian on
    whatever kernel object it is given): matrices from the user's methods, linear algebra on the
    device; parity against the oracle's Gaussian GP on the same inputs."""
    x, y = synth_xy(n, 5)
    xo = np.linspace(-6, 6, m)
    kp, s = (1.2, 0.45), 0.8
    gp = GP(NumpyGaussianKernel(*kp), x, y, s=s)
    o = oracle.OracleGP(oracle.GAUSSIAN, kp, x, y, s)
    for key in ("Kxx", "Lxx", "inv_Kxx", "inv_Kxx_y", "log_lh", "dloglh_dtheta", "dlh_dtheta"):
        assert_parity(getattr(gp, key), getattr(o, key), RTOL, key)
    assert_parity(gp.mean(xo), o.mean(xo), RTOL, "mean")
    assert_parity(gp.cov(xo), o.cov(xo), RTOL, "cov")
    assert_parity(gp.var(xo), np.diag(o.cov(xo)), RTOL, "var")
    assert_parity(gp.dm_dtheta(xo), o.dm_dtheta(xo), RTOL, "dm")
    assert_parity(gp.d2loglh_normalised(), o.d2lh_dtheta2_with(1.0, o.dloglh_dtheta), RTOL, "d2lh(lh=1)")
    # setters rebind the resident engine; copies and pickles keep working
    gp.set_param("w", 0.6)
    o2 = oracle.OracleGP(oracle.GAUSSIAN, (1.2, 0.6), x, y, s)
    assert_parity(gp.log_lh, o2.log_lh)
    assert_parity(gp.dloglh_dtheta, o2.dloglh_dtheta)
    import pickle
    g3 = pickle.loads(pickle.dumps(gp))
    assert_parity(g3.mean(xo), o2.mean(xo))
    cand = np.array([[1.2, 0.6, 0.8], [1.0, 0.5, 1.0]])
    llh, grad = gp.batch_eval(cand)
    assert_parity(llh[0], o2.log_lh) and assert_parity(grad[0], o2.dloglh_dtheta)
    res = gp.fit_MLII(cand)
    assert res.best_index == int(np.argmax(llh))


# ------------------------------------------------------------------ generated CUDA functors from sym_K
@pytest.mark.parametrize("kparams", [(1.3, 0.4), (0.7, 0.9, 1.7)])
def test_symbolic_kernel_builders_vs_oracle(oracle, kparams):
    """The reference checks its Cython against sym_K (test_gaussian_kernel.py:67-88); here the code
    generated FROM sym_K (NVRTC, sm_100a) is checked against the oracle's element loops."""
    k = gpb.SymbolicKernel.from_kernel(make_kernel(kparams))
    kind = kind_of(oracle, kparams)
    rng = np.random.RandomState(3)
    for n1, n2 in ((10, 10), (1, 7), (130, 67), (257, 300)):
        x1, x2 = rng.uniform(-3, 3, n1), rng.uniform(-3, 3, n2)
        assert_parity(k(x1, x2), oracle.K(kind, x1, x2, kparams), 1e-13, "K")
        assert_parity(k.jacobian(x1, x2), oracle.jacobian(kind, x1, x2, kparams), 1e-12, "J")
        assert_parity(k.hessian(x1, x2), oracle.hessian(kind, x1, x2, kparams), 1e-11, "H")
    assert k(np.empty(0), x2).shape == (0, x2.size)


@pytest.mark.parametrize("n,kparams", [(60, (1.2, 0.45)), (300, (1.2, 0.45)), (200, (0.9, 0.8, 1.4))])
def test_gp_over_symbolic_kernel_vs_oracle(oracle, n, kparams):
    """GP whose kernel matrices are built on the device by generated code; everything else the
    library's kernels.  Parity against the oracle's GP for the same kernel function."""
    x, y = synth_xy(n, 9)
    xo = np.linspace(-6, 6, 70)
    s = 0.8
    gp = GP(gpb.SymbolicKernel.from_kernel(make_kernel(kparams)), x, y, s=s)
    o = oracle.OracleGP(kind_of(oracle, kparams), kparams, x, y, s)
    for key in ("Kxx", "Lxx", "inv_Kxx", "inv_Kxx_y", "log_lh", "dloglh_dtheta", "Kxx_J"):
        assert_parity(getattr(gp, key), getattr(o, key), RTOL, key)
    assert_parity(gp.mean(xo), o.mean(xo), RTOL, "mean")
    assert_parity(gp.cov(xo), o.cov(xo), RTOL, "cov")
    assert_parity(gp.dm_dtheta(xo), o.dm_dtheta(xo), RTOL, "dm")
    assert_parity(gp.d2loglh_normalised(), o.d2lh_dtheta2_with(1.0, o.dloglh_dtheta), RTOL, "d2lh(lh=1)")


def test_symbolic_kernel_new_function_vs_lambdified():
    """A kernel the reference does not have (rational quadratic): generated builders against the
    same sympy expressions evaluated by numpy, and the GP's log_lh / gradient against a direct
    scipy computation."""
    import sympy as sym
    from scipy.linalg import cho_factor, cho_solve
    h, w, a, d = sym.symbols("h w a d")
    expr = h ** 2 * (1 + d ** 2 / (2 * a * w ** 2)) ** (-a)
    vals = (1.1, 0.6, 1.7)
    k = gpb.SymbolicKernel(expr, ("h", "w", "a"), vals)
    x, y = synth_xy(150, 4)
    D_ = x[:, None] - x[None, :]
    f = sym.lambdify((d, h, w, a), expr, "numpy")
    K = f(D_, *vals)
    J = np.stack([sym.lambdify((d, h, w, a), sym.diff(expr, p), "numpy")(D_, *vals) for p in (h, w, a)])
    assert_parity(k(x, x), K, 1e-13) and assert_parity(k.jacobian(x, x), J, 1e-12)
    Hww = sym.lambdify((d, h, w, a), sym.diff(expr, w, w), "numpy")(D_, *vals)
    assert_parity(k.hessian(x, x)[1, 1], Hww, 1e-11)
    s = 0.7
    gp = GP(k, x, y, s=s)
    Kn = K + s ** 2 * np.eye(x.size)
    cf = cho_factor(Kn, lower=True)
    al = cho_solve(cf, y)
    llh = -0.5 * y @ al - np.sum(np.log(np.diag(cf[0]))) - 0.5 * x.size * np.log(2 * np.pi)
    Ki = cho_solve(cf, np.eye(x.size))
    grad = [0.5 * al @ Ji @ al - 0.5 * np.sum(Ki * Ji) for Ji in J] + [s * (al @ al) - s * np.trace(Ki)]
    assert_parity(gp.log_lh, llh) and assert_parity(gp.dloglh_dtheta, np.array(grad))


# ------------------------------------------------------------------ host-buffer posterior entry points
def test_mean_pinned_and_pageable_paths(oracle):
    """gpb_post_mean_host stages small test sets through a page-locked buffer and copies larger ones
    (> 64 MB of staging) straight from / to pageable memory: both against the oracle on a subset."""
    x, y = synth_xy(40, 2)
    gp = GP(GaussianKernel(1.0, 0.6), x, y, s=0.5)
    o = oracle.OracleGP(oracle.GAUSSIAN, (1.0, 0.6), x, y, 0.5)
    rng = np.random.RandomState(0)
    for m in (1, 33, 70000, 4300000):
        xo = rng.uniform(-7, 7, m)
        got = gp.mean(xo)
        assert got.shape == (m,)
        idx = rng.randint(0, m, size=min(m, 500))
        assert_parity(got[idx], o.mean(xo[idx]), RTOL, "mean m=%d" % m)


@pytest.mark.parametrize("m", [511, 512, 513])
def test_cov_host_call_and_panel_path_agree(oracle, m):
    """cov() is one gpb_post_cov_host call up to 512 test points and the pipelined panel path above."""
    x, y = synth_xy(130, 6)
    xo = np.linspace(-6.5, 6.5, m)
    gp = GP(PeriodicKernel(1.1, 0.9, 1.6), x, y, s=0.6)
    o = oracle.OracleGP(oracle.PERIODIC, (1.1, 0.9, 1.6), x, y, 0.6)
    c = gp.cov(xo)
    assert_parity(c, o.cov(xo), RTOL, "cov m=%d" % m)
    assert np.array_equal(c, c.T)


def test_log_likelihood_hessian_vs_finite_differences():
    """d2loglh_dtheta2 (additive API) against central differences of dloglh_dtheta, at a size where
    lh underflows to 0 and the reference's d2lh_dtheta2 is identically zero."""
    x, y = synth_xy(1500, 8)
    th0 = np.array([1.1, 0.45, 0.9])
    gp = GP(GaussianKernel(*th0[:2]), x, y, s=th0[2])
    assert gp.lh == 0 and not np.any(gp.d2lh_dtheta2)
    H = gp.d2loglh_dtheta2()
    assert np.allclose(H, H.T, rtol=1e-9, atol=1e-6 * np.abs(H).max())
    eps = 1e-5
    for j in range(3):
        tp, tm = th0.copy(), th0.copy()
        tp[j] += eps
        tm[j] -= eps
        gp.params = tp
        gpl = gp.dloglh_dtheta.copy()
        gp.params = tm
        gmi = gp.dloglh_dtheta.copy()
        fd = (gpl - gmi) / (2 * eps)
        assert np.allclose(H[:, j], fd, rtol=2e-5, atol=2e-5 * np.abs(H).max()), (j, H[:, j], fd)
