"""Host-boundary throughput: matrix-valued results crossing PCIe (cov, kernel builders with the
Cython signatures, Kxx / inv_Kxx properties).  Prints one JSON line."""
import json, os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch
import gaussian_processes_b200 as gpb
from gaussian_processes_b200.ext import gaussian_c, periodic_c
from conftest import synth_xy

out = {}
n = 4096
x, y = synth_xy(n, 0)
buf = np.empty((n, n))
for name, fn, ns in (("gaussian_K", lambda o: gaussian_c.K(o, x, x, 1.0, 0.5), 1),):
    fn(buf)
    t0 = time.perf_counter()
    for k in range(5):
        fn(buf)
    dt = (time.perf_counter() - t0) / 5
    out[name + "_host_ms"] = dt * 1e3
    out[name + "_host_GBps"] = ns * n * n * 8 / dt / 1e9
jb = np.empty((3, n, n))
periodic_c.jacobian(jb, x, x, 1.0, 1.0, 1.0)
t0 = time.perf_counter()
periodic_c.jacobian(jb, x, x, 1.0, 1.0, 1.0)
dt = time.perf_counter() - t0
out["periodic_jacobian_host_ms"] = dt * 1e3
out["periodic_jacobian_host_GBps"] = jb.nbytes / dt / 1e9
gp = gpb.GP(gpb.GaussianKernel(1.0, 0.5), x, y, s=1.0)
for m in (1024, 4096, 8192):
    xo = np.linspace(-6, 6, m)
    c = gp.cov(xo); del c
    ts = []
    for k in range(3):
        t0 = time.perf_counter()
        c = gp.cov(xo + 1e-9 * k)
        ts.append(time.perf_counter() - t0)
        del c
    t0 = time.perf_counter()
    d = gp._engine().cov(xo, host=False); torch.cuda.synchronize()
    td = time.perf_counter() - t0
    del d
    out["cov_m%d_ms" % m] = min(ts) * 1e3
    out["cov_m%d_device_ms" % m] = td * 1e3
    out["cov_m%d_pts_per_s" % m] = m / min(ts)
t0 = time.perf_counter(); gp.set_param("w", 0.51); k1 = gp.Kxx; out["Kxx_prop_ms"] = (time.perf_counter() - t0) * 1e3
t0 = time.perf_counter(); ki = gp.inv_Kxx; out["inv_Kxx_prop_ms"] = (time.perf_counter() - t0) * 1e3
# symmetry / correctness of the panelled cov against the single-panel path on a small case
xo = np.linspace(-6, 6, 3000)
c = gp.cov(xo)
out["cov_sym"] = float(np.abs(c - c.T).max())
kx = gaussian_c.K
Kxox = np.empty((3000, n)); kx(Kxox, xo, x, 1.0, 0.51)
Kxoxo = np.empty((3000, 3000)); kx(Kxoxo, xo, xo, 1.0, 0.51)
ref = Kxoxo - Kxox @ ki @ Kxox.T
out["cov_vs_numpy"] = float(np.abs(c - ref).max() / np.abs(ref).max())
print(json.dumps(out))
