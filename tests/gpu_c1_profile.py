"""cProfile of the C1 bundle (N=50): where the host-side time of a tiny GP goes."""
import cProfile, pstats, os, sys, io
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch
import gaussian_processes_b200 as gpb
x = np.linspace(-2 * np.pi, 2 * np.pi, 50); y = np.sin(x); xo = np.linspace(-2 * np.pi, 2 * np.pi, 100)
gp = gpb.GP(gpb.GaussianKernel(1.0, 0.2), x, y, s=0)
def bundle(k):
    gp.set_param("w", 0.2 + 1e-6 * (k + 1))
    return gp.log_lh, gp.dloglh_dtheta, gp.mean(xo), gp.cov(xo)
for k in range(20): bundle(k)
torch.cuda.synchronize()
pr = cProfile.Profile(); pr.enable()
for k in range(200): bundle(100 + k)
pr.disable()
s = io.StringIO(); pstats.Stats(pr, stream=s).sort_stats("cumulative").print_stats(45); print(s.getvalue()[:9000])
