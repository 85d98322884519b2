"""Small end-to-end workload for compute-sanitizer (memcheck / racecheck / initcheck / synccheck):
one-block and multi-block GPs of both kernels, batched evaluation, posterior calls."""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch
import gaussian_processes_b200 as gpb
from conftest import synth_xy
xo = np.linspace(-6, 6, 70)
for n in (50, 97, 300):
    x, y = synth_xy(n, n)
    for K in (gpb.GaussianKernel(1.1, 0.4), gpb.PeriodicKernel(0.9, 0.8, 1.4)):
        gp = gpb.GP(K, x, y, s=0.7)
        r = (gp.log_lh, gp.dloglh_dtheta, gp.mean(xo), gp.cov(xo), gp.var(xo), gp.dm_dtheta(xo), gp.Lxx, gp.inv_Kxx)
        if n != 300:
            gp.d2lh_dtheta2
        th = np.stack([gp.params, gp.params * 1.05, gp.params * 0.95])
        gp.batch_eval(th)
        print(n, type(K).__name__, float(r[0]), flush=True)
torch.cuda.synchronize()
print("done")
