import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


def load_oracle():
    """tests are one of the three places allowed to import oracle/."""
    import importlib.util
    spec = importlib.util.spec_from_file_location("gp_oracle", os.path.join(ROOT, "oracle", "gp_oracle.py"))
    mod = sys.modules.get("gp_oracle")
    if mod is None:
        mod = importlib.util.module_from_spec(spec)
        sys.modules["gp_oracle"] = mod
        spec.loader.exec_module(mod)
    return mod


@pytest.fixture(scope="session")
def oracle():
    return load_oracle()


def golden(name):
    return dict(np.load(os.path.join(GOLDEN, name + ".npz")))


def synth_xy(n, seed=0):
    """x = sort(U(-2pi, 2pi, n)), y = sin x + 0.1 N(0,1), RandomState(seed)."""
    rng = np.random.RandomState(seed)
    x = np.sort(rng.uniform(-2 * np.pi, 2 * np.pi, n))
    y = np.sin(x) + 0.1 * rng.randn(n)
    return x, y


#: parity tolerance (BASELINE.json north_star / SURVEY 8d): scalars |d|/|ref| <= 1e-9,
#: arrays ||d||_inf <= 1e-9 * ||ref||_inf
RTOL = 1e-9


def assert_parity(got, ref, rtol=RTOL, what=""):
    got = np.asarray(got, dtype=np.float64)
    ref = np.asarray(ref, dtype=np.float64)
    assert got.shape == ref.shape, "%s: shape %s vs %s" % (what, got.shape, ref.shape)
    nan_g, nan_r = np.isnan(got), np.isnan(ref)
    assert (nan_g == nan_r).all(), "%s: NaN pattern differs" % what
    inf_r = np.isinf(ref)
    assert (got[inf_r] == ref[inf_r]).all(), "%s: inf pattern differs" % what
    fin = ~(nan_r | inf_r)
    if not fin.any():
        return
    scale = np.max(np.abs(ref[fin]))
    err = np.max(np.abs(got[fin] - ref[fin]))
    assert err <= rtol * scale + 1e-300, "%s: |d|_inf=%.3e > %.1e * |ref|_inf=%.3e (rel %.3e)" % (
        what, err, rtol, scale, err / max(scale, 1e-300))
