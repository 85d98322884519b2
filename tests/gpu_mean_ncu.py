"""mean_kernel / K-only build alone (for ncu --set full)."""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch
import gaussian_processes_b200 as gpb
from gaussian_processes_b200 import engine, device as D
from conftest import synth_xy
n = 4096
x, y = synth_xy(n, 0)
gp = gpb.GP(gpb.GaussianKernel(1.0, 0.5), x, y, s=1.0)
gp.log_lh
for m in (16384, 262144):
    gp.mean(np.linspace(-6, 6, m))
eng = gp._engine()
buf = D.empty(1, n, n)
eng.build(eng.dx, n, eng.dx, n, n, n, 1, out=buf)
torch.cuda.synchronize()
