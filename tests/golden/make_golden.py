"""Generate tests/golden/*.npz from the UNMODIFIED reference (build container only).

Runs the reference's own Python (``/root/reference/gp/gp.py``, ``gp/kernels/*.py``)
on top of the reference's own Cython compiled in place (``oracle/_ref``), and
stores input/output vectors as small fixtures.  ``/root/reference`` does not
exist on the GPU box, so the fixtures -- not this script -- travel.

The reference is Python-2 era: its package ``__init__`` files use implicit
relative imports (gp/ext/__init__.py:1-3, gp/kernels/__init__.py:1-3) and
gp/gp.py:4 imports matplotlib (only ``plot`` uses it).  Nothing is copied or
edited: the modules are exec'd from where they lie into a synthetic ``gp``
package whose ``gp.ext`` is the compiled oracle/_ref modules, with an empty
stand-in for ``matplotlib.pyplot``.

usage: python tests/golden/make_golden.py
"""
import importlib.util
import os
import sys
import types

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
REF = "/root/reference/gp"


def load_reference():
    """Returns the reference ``gp`` package object (GP, GaussianKernel, PeriodicKernel)."""
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import build_ref
    build_ref.build()
    gaussian_c, periodic_c, gp_c = build_ref.load()

    if "matplotlib" not in sys.modules:
        try:
            import matplotlib.pyplot  # noqa: F401
        except Exception:
            mpl = types.ModuleType("matplotlib")
            plt = types.ModuleType("matplotlib.pyplot")
            mpl.pyplot = plt
            sys.modules["matplotlib"] = mpl
            sys.modules["matplotlib.pyplot"] = plt

    def pkg(name, path):
        m = types.ModuleType(name)
        m.__path__ = [path]
        m.__package__ = name
        sys.modules[name] = m
        return m

    def load(name, path):
        spec = importlib.util.spec_from_file_location(name, path)
        m = importlib.util.module_from_spec(spec)
        sys.modules[name] = m
        spec.loader.exec_module(m)
        return m

    gp = pkg("gp", REF)
    ext = pkg("gp.ext", os.path.join(REF, "ext"))
    ext.gaussian_c, ext.periodic_c, ext.gp_c = gaussian_c, periodic_c, gp_c
    sys.modules["gp.ext.gaussian_c"] = gaussian_c
    sys.modules["gp.ext.periodic_c"] = periodic_c
    sys.modules["gp.ext.gp_c"] = gp_c
    gp.ext = ext
    kernels = pkg("gp.kernels", os.path.join(REF, "kernels"))
    gp.kernels = kernels
    base = load("gp.kernels.base", os.path.join(REF, "kernels", "base.py"))
    kernels.Kernel = base.Kernel
    gaussian = load("gp.kernels.gaussian", os.path.join(REF, "kernels", "gaussian.py"))
    periodic = load("gp.kernels.periodic", os.path.join(REF, "kernels", "periodic.py"))
    kernels.GaussianKernel = gaussian.GaussianKernel
    kernels.PeriodicKernel = periodic.PeriodicKernel
    gpmod = load("gp.gp", os.path.join(REF, "gp.py"))
    gp.GP = gpmod.GP
    gp.GaussianKernel = gaussian.GaussianKernel
    gp.PeriodicKernel = periodic.PeriodicKernel
    gp.Kernel = base.Kernel
    return gp


def synth_xy(n, seed=0):
    """SURVEY 8(d) input law: x = sort(U(-2pi, 2pi, n)), y = sin x + 0.1 N(0,1)."""
    rng = np.random.RandomState(seed)
    x = np.sort(rng.uniform(-2 * np.pi, 2 * np.pi, n))
    y = np.sin(x) + 0.1 * rng.randn(n)
    return x, y


def gp_bundle(gp, xo, ref_gp_c=None, full=True):
    """Everything the hot path exposes, from the reference object."""
    out = {
        "x": gp.x, "y": gp.y, "params": gp.params, "xo": xo,
        "log_lh": np.float64(gp.log_lh), "lh": np.float64(gp.lh),
        "inv_Kxx_y": gp.inv_Kxx_y,
        "dloglh_dtheta": gp.dloglh_dtheta, "dlh_dtheta": gp.dlh_dtheta,
        "d2lh_dtheta2": gp.d2lh_dtheta2,
        "mean": gp.mean(xo), "dm_dtheta": gp.dm_dtheta(xo),
    }
    # lh-normalised second derivative: gp_c.d2lh_dtheta2(lh=1, dlh=dloglh)
    n = gp.params.size
    d2n = np.empty((n, n))
    ref_gp_c.d2lh_dtheta2(gp.y, gp.inv_Kxx, gp.Kxx_J, gp.Kxx_H, gp.inv_Kxx_y, gp.s,
                          1.0, gp.dloglh_dtheta, d2n)
    out["d2lh_norm"] = d2n
    if full:
        out.update({"Kxx": gp.Kxx, "Kxx_J": gp.Kxx_J, "Kxx_H": gp.Kxx_H, "Lxx": gp.Lxx,
                    "inv_Kxx": gp.inv_Kxx, "cov": gp.cov(xo)})
    else:
        c = gp.cov(xo)
        out["cov_diag"] = np.diag(c).copy()
        out["cov_row0"] = c[0].copy()
        out["cov_fro"] = np.float64(np.linalg.norm(c))
        out["inv_Kxx_diag"] = np.diag(gp.inv_Kxx).copy()
        out["Lxx_diag"] = np.diag(gp.Lxx).copy()
    return out


def main():
    gp = load_reference()
    gp_c = sys.modules["gp.ext.gp_c"]
    G, P, GP = gp.GaussianKernel, gp.PeriodicKernel, gp.GP

    # --- kernel slices on the reference suite's grids (test_gaussian_kernel.py:44,
    #     test_periodic_kernel.py:47) plus a ragged n1 != n2 case
    rng = np.random.RandomState(2348)          # tests/util.py:47-48
    ker = {}
    x10 = np.linspace(-2, 2, 10)
    x16 = np.linspace(-2 * np.pi, 2 * np.pi, 16)
    xr = np.sort(rng.uniform(-7, 7, 23))
    for t in range(4):
        h, w, p = rng.uniform(0, 2), rng.uniform(np.pi / 32., np.pi / 2.), rng.uniform(0.33, 3)
        kg, kp = G(h, w), P(h, w, p)
        ker["g%d_params" % t] = kg.params
        ker["p%d_params" % t] = kp.params
        for tag, k, xa, xb in (("g", kg, x10, x10), ("p", kp, x16, x16)):
            ker["%s%d_K" % (tag, t)] = k(xa, xb)
            ker["%s%d_J" % (tag, t)] = k.jacobian(xa, xb)
            ker["%s%d_H" % (tag, t)] = k.hessian(xa, xb)
        ker["g%d_Kr" % t] = kg(xr, x10)
        ker["g%d_Jr" % t] = kg.jacobian(xr, x10)
        ker["g%d_Hr" % t] = kg.hessian(xr, x10)
        ker["p%d_Kr" % t] = kp(xr, x16)
        ker["p%d_Jr" % t] = kp.jacobian(xr, x16)
        ker["p%d_Hr" % t] = kp.hessian(xr, x16)
    # far-apart points: the e < MIN -> exact 0 rule (gaussian_c.pyx:33-34)
    kz = G(1.0, 0.05)
    xz = np.array([0.0, 0.5, 1.8, 1.9, 4.0])
    ker["gz_params"] = kz.params
    ker["gz_x"] = xz
    ker["gz_K"], ker["gz_J"], ker["gz_H"] = kz(xz, xz), kz.jacobian(xz, xz), kz.hessian(xz, xz)
    ker.update(x10=x10, x16=x16, xr=xr)
    np.savez_compressed(os.path.join(HERE, "kernels.npz"), **ker)

    # --- suite-scale GP (tests/util.py:36-44: 16 train pts, 32 test pts), seed 2348
    x = np.linspace(-2 * np.pi, 2 * np.pi, 16)
    y = np.sin(x)
    xo = np.linspace(-2 * np.pi, 2 * np.pi, 32)
    np.random.seed(2348)
    for t in range(3):
        h = np.random.uniform(0, 2)
        w = np.random.uniform(np.pi / 32., np.pi / 2.)
        s = np.random.uniform(0, 0.5)
        g = GP(G(h, w), x, y, s=s)
        np.savez_compressed(os.path.join(HERE, "gp_suite_g%d.npz" % t), **gp_bundle(g, xo, gp_c))
        p = np.random.uniform(0.33, 3)
        g = GP(P(h, w, p), x, y, s=s)
        np.savez_compressed(os.path.join(HERE, "gp_suite_p%d.npz" % t), **gp_bundle(g, xo, gp_c))

    # --- C1: N=50 linspace, Gaussian(1, 0.2), s=0, M=100
    x = np.linspace(-2 * np.pi, 2 * np.pi, 50)
    xo = np.linspace(-2 * np.pi, 2 * np.pi, 100)
    g = GP(G(1.0, 0.2), x, np.sin(x), s=0)
    np.savez_compressed(os.path.join(HERE, "gp_c1.npz"), **gp_bundle(g, xo, gp_c))

    # --- multi-block sizes (exercise the blocked Cholesky): N=300 Gaussian, N=257 Periodic
    x, y = synth_xy(300, 0)
    xo = np.linspace(-2 * np.pi, 2 * np.pi, 77)
    g = GP(G(1.0, 0.5), x, y, s=1.0)
    np.savez_compressed(os.path.join(HERE, "gp_g300.npz"), **gp_bundle(g, xo, gp_c, full=False))
    x, y = synth_xy(257, 1)
    g = GP(P(1.0, 1.0, 1.0), x, y, s=1.0)
    np.savez_compressed(os.path.join(HERE, "gp_p257.npz"), **gp_bundle(g, xo, gp_c, full=False))

    # --- C2-shaped scalars at N=1024 and N=4096 (inputs are regenerated from the seed)
    for n in (1024, 4096):
        x, y = synth_xy(n, 0)
        g = GP(G(1.0, 0.5), x, y, s=1.0)
        xo = np.linspace(-2 * np.pi, 2 * np.pi, 64)
        d2n = np.empty((3, 3))
        gp_c.d2lh_dtheta2(g.y, g.inv_Kxx, g.Kxx_J, g.Kxx_H, g.inv_Kxx_y, g.s, 1.0,
                          g.dloglh_dtheta, d2n)
        c = g.cov(xo)
        np.savez_compressed(
            os.path.join(HERE, "gp_c2_n%d.npz" % n), n=n, seed=0, params=g.params, xo=xo,
            log_lh=np.float64(g.log_lh), dloglh_dtheta=g.dloglh_dtheta,
            lh=np.float64(g.lh), d2lh_norm=d2n, mean=g.mean(xo), cov=c,
            dm_dtheta=g.dm_dtheta(xo), inv_Kxx_y=g.inv_Kxx_y,
            inv_Kxx_diag=np.diag(g.inv_Kxx).copy(), Lxx_diag=np.diag(g.Lxx).copy())
        print("N=%d log_lh=%.10f dloglh=%s" % (n, g.log_lh, g.dloglh_dtheta))

    # --- the suite's one hard-coded failure vector (test_gp.py:298-333)
    # (inputs are restated in tests/; only the expected behaviour is recorded here)
    print("golden fixtures written to", HERE)


if __name__ == "__main__":
    main()
