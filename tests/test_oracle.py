"""Pin the CPU oracle (oracle/) against the reference: golden vectors produced by
the unmodified reference (tests/golden/make_golden.py), the reference's compiled
Cython (oracle/_ref, when present), the closed forms the reference suite uses and
its one hard-coded known-answer case (gp/tests/test_gp.py:298-333)."""
import numpy as np
import pytest

from conftest import golden, assert_parity, synth_xy


def _kind(o, tag):
    return o.GAUSSIAN if tag == "g" else o.PERIODIC


@pytest.mark.parametrize("tag", ["g", "p"])
@pytest.mark.parametrize("t", range(4))
def test_kernel_slices_vs_golden(oracle, tag, t):
    g = golden("kernels")
    kind = _kind(oracle, tag)
    kp = g["%s%d_params" % (tag, t)]
    xa = g["x10"] if tag == "g" else g["x16"]
    for impl in ("c",) + (("ref",) if oracle.have_ref() else ()):
        assert_parity(oracle.K(kind, xa, xa, kp, impl), g["%s%d_K" % (tag, t)], 1e-14, "K")
        assert_parity(oracle.jacobian(kind, xa, xa, kp, impl), g["%s%d_J" % (tag, t)], 1e-14, "J")
        assert_parity(oracle.hessian(kind, xa, xa, kp, impl), g["%s%d_H" % (tag, t)], 1e-13, "H")
        # ragged n1 != n2
        assert_parity(oracle.K(kind, g["xr"], xa, kp, impl), g["%s%d_Kr" % (tag, t)], 1e-14, "Kr")
        assert_parity(oracle.jacobian(kind, g["xr"], xa, kp, impl), g["%s%d_Jr" % (tag, t)], 1e-14, "Jr")
        assert_parity(oracle.hessian(kind, g["xr"], xa, kp, impl), g["%s%d_Hr" % (tag, t)], 1e-13, "Hr")


def test_gaussian_min_rule(oracle):
    """Entries whose exponent is below MIN are exactly 0 in every slice (gaussian_c.pyx:33-34)."""
    g = golden("kernels")
    K = oracle.K(oracle.GAUSSIAN, g["gz_x"], g["gz_x"], g["gz_params"])
    assert (K == 0).sum() == (g["gz_K"] == 0).sum() > 0
    assert ((K == 0) == (g["gz_K"] == 0)).all()
    assert_parity(oracle.jacobian(oracle.GAUSSIAN, g["gz_x"], g["gz_x"], g["gz_params"]), g["gz_J"], 1e-14)
    assert_parity(oracle.hessian(oracle.GAUSSIAN, g["gz_x"], g["gz_x"], g["gz_params"]), g["gz_H"], 1e-14)
    assert abs(oracle.MIN - (-705.6238298100243)) < 1e-12


def test_closed_forms(oracle):
    """test_gaussian_kernel.py:44-64 and test_periodic_kernel.py:47-64."""
    rng = np.random.RandomState(2348)
    x = np.linspace(-2, 2, 10)
    for _ in range(20):
        h, w, p = rng.uniform(0, 2), rng.uniform(np.pi / 32., np.pi / 2.), rng.uniform(0.33, 3)
        assert np.allclose(oracle.K(oracle.GAUSSIAN, x, x, (h, w)),
                           oracle.gaussian_closed_form(x, x, h, w), rtol=1e-5)
        assert np.allclose(oracle.K(oracle.PERIODIC, x, x, (h, w, p)),
                           oracle.periodic_closed_form(x, x, h, w, p), rtol=1e-5)


def test_c_restatement_matches_ref_bitwise(oracle):
    """Same libm, same operation order: the C restatement reproduces oracle/_ref."""
    if not oracle.have_ref():
        pytest.skip("oracle/_ref not built")
    rng = np.random.RandomState(1)
    x1 = rng.uniform(-6, 6, 37)
    x2 = rng.uniform(-6, 6, 29)
    for kind, kp, ns in ((oracle.GAUSSIAN, (1.3, 0.4), 7), (oracle.PERIODIC, (0.7, 0.9, 1.7), 13)):
        for sl in range(ns):
            a = oracle.kernel_slice(kind, sl, x1, x2, kp, "c")
            b = oracle.kernel_slice(kind, sl, x1, x2, kp, "ref")
            assert_parity(a, b, 4e-16, "slice %d" % sl)


GP_KEYS = ["Kxx", "Kxx_J", "Kxx_H", "Lxx", "inv_Kxx", "inv_Kxx_y", "log_lh", "dloglh_dtheta",
           "dlh_dtheta", "d2lh_dtheta2", "mean", "cov", "dm_dtheta"]


@pytest.mark.parametrize("name", ["gp_suite_g0", "gp_suite_g1", "gp_suite_g2", "gp_suite_p0",
                                  "gp_suite_p1", "gp_suite_p2", "gp_c1"])
def test_gp_vs_golden(oracle, name):
    g = golden(name)
    kind = oracle.GAUSSIAN if g["params"].size == 3 else oracle.PERIODIC
    for impl in ("c",) + (("ref",) if oracle.have_ref() else ()):
        o = oracle.OracleGP(kind, g["params"][:-1], g["x"], g["y"], g["params"][-1], impl)
        for k in GP_KEYS:
            got = getattr(o, k)(g["xo"]) if k in ("mean", "cov", "dm_dtheta") else getattr(o, k)
            # the oracle is the same algorithm on the same BLAS: it must agree far
            # inside the 1e-9 product tolerance
            assert_parity(got, g[k], 1e-11, "%s/%s/%s" % (name, impl, k))
        assert float(o.lh) == float(g["lh"])
        assert_parity(o.d2lh_dtheta2_with(1.0, o.dloglh_dtheta), g["d2lh_norm"], 1e-11, "d2lh_norm")


@pytest.mark.parametrize("name", ["gp_g300", "gp_p257"])
def test_gp_medium_vs_golden(oracle, name):
    g = golden(name)
    kind = oracle.GAUSSIAN if g["params"].size == 3 else oracle.PERIODIC
    o = oracle.OracleGP(kind, g["params"][:-1], g["x"], g["y"], g["params"][-1])
    assert_parity(o.log_lh, g["log_lh"], 1e-12)
    assert_parity(o.dloglh_dtheta, g["dloglh_dtheta"], 1e-11)
    assert_parity(o.mean(g["xo"]), g["mean"], 1e-11)
    assert_parity(np.diag(o.cov(g["xo"])), g["cov_diag"], 1e-10)
    assert_parity(o.dm_dtheta(g["xo"]), g["dm_dtheta"], 1e-10)
    assert_parity(o.d2lh_dtheta2_with(1.0, o.dloglh_dtheta), g["d2lh_norm"], 1e-10)


def test_c2_scalars_n1024(oracle):
    g = golden("gp_c2_n1024")
    x, y = synth_xy(1024, 0)
    o = oracle.OracleGP(oracle.GAUSSIAN, g["params"][:-1], x, y, g["params"][-1])
    assert_parity(o.log_lh, g["log_lh"], 1e-12)
    assert_parity(o.dloglh_dtheta, g["dloglh_dtheta"], 1e-11)
    assert o.lh == 0 and float(g["lh"]) == 0            # underflow to int 0 (gp.py:393-394)


def test_invalid_params_known_answer(oracle):
    """gp/tests/test_gp.py:298-333 -- the reference suite's only hard-coded vector."""
    from suite_util import INVALID_X, INVALID_Y, INVALID_H, INVALID_W
    o = oracle.OracleGP(oracle.GAUSSIAN, (INVALID_H, INVALID_W), INVALID_X, INVALID_Y, 0.0)
    for prop in ("Lxx", "inv_Kxx", "inv_Kxx_y"):
        with pytest.raises(np.linalg.LinAlgError):
            getattr(o, prop)
    assert o.log_lh == -np.inf
    assert o.lh == 0
    assert np.isnan(o.dloglh_dtheta).all()
    assert np.isnan(o.dlh_dtheta).all()
    assert np.isnan(o.d2lh_dtheta2).all()


def test_oracle_fit_mlii_definition(oracle):
    x, y = synth_xy(64, 3)
    rng = np.random.RandomState(5)
    cand = np.stack([rng.uniform(0.5, 2, 6), rng.uniform(np.pi / 32, np.pi / 2, 6),
                     rng.uniform(0.75, 1.5, 6)], axis=1)
    best, llh, grad = oracle.oracle_fit_mlii(oracle.GAUSSIAN, x, y, cand)
    assert best == int(np.argmax(llh)) and grad.shape == (6, 3)
