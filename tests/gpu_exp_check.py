"""Accuracy of the functor exp against numpy (libm) over the Gaussian / periodic argument ranges, and
throughput of the exp-bound kernels (mean, K-only build)."""
import json, os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch
import gaussian_processes_b200 as gpb
from gaussian_processes_b200 import engine, device as D
from conftest import synth_xy
out = {}
# K(x1, x2) with h^2/(w sqrt(2 pi)) = 1 isolates exp(-d^2 / 2w^2)
w = 0.37
h = np.sqrt(w * np.sqrt(2 * np.pi))
k = gpb.GaussianKernel(h, w)
rng = np.random.RandomState(0)
x1 = np.zeros(1)
d = np.concatenate([rng.uniform(0, 14, 2000000), np.linspace(0, 13.9, 500000)])
got = k(x1, d)[0]
c1 = -0.5 / np.float64(w) ** 2
e = c1 * (0.0 - d) ** 2
ref = np.where(e < -705.6238298100243, 0.0, (0.5 * np.sqrt(2 / np.pi) * h * h / w) * np.exp(e))
nz = ref > 0
ulp = np.abs(got[nz] - ref[nz]) / np.spacing(ref[nz])
out["gauss_exp_max_ulp"] = float(ulp.max()); out["gauss_exp_mean_ulp"] = float(ulp.mean())
out["zeros_match"] = bool(((got == 0) == (ref == 0)).all())
# throughput
n = 4096
x, y = synth_xy(n, 0)
gp = gpb.GP(gpb.GaussianKernel(1.0, 0.5), x, y, s=1.0)
for m in (16384, 262144):
    xo = np.linspace(-6, 6, m)
    gp.mean(xo[:100]); torch.cuda.synchronize()
    t0 = time.perf_counter()
    for r in range(5): gp.mean(xo + 1e-9 * r)
    out["mean_pts_per_s_m%d" % m] = m / ((time.perf_counter() - t0) / 5)
e_ = gp._engine()
for nn in (4096, 8192):
    xx, _ = synth_xy(nn, 1)
    eng = engine.Engine(engine.GAUSSIAN, (1.0, 0.5), 1.0, xx, xx)
    buf = D.empty(1, nn, nn)
    for mask, nm in ((1, "K"),):
        eng.build(eng.dx, nn, eng.dx, nn, nn, nn, mask, out=buf); torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for r in range(5): eng.build(eng.dx, nn, eng.dx, nn, nn, nn, mask, out=buf)
        e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 5
        out["build_%s_n%d" % (nm, nn)] = dict(ms=ms, GBps=8.0 * nn * nn / ms / 1e6, Gelem_per_s=nn * nn / ms / 1e6)
print(json.dumps(out))
