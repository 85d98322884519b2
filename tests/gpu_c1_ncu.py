"""A few C1 bundles (N=50) for an ncu launch list: which kernels a tiny GP launches and how long they run."""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import gaussian_processes_b200 as gpb
n = int(sys.argv[1]) if len(sys.argv) > 1 else 50
x = np.linspace(-2 * np.pi, 2 * np.pi, n); y = np.sin(x); xo = np.linspace(-2 * np.pi, 2 * np.pi, 100)
gp = gpb.GP(gpb.GaussianKernel(1.0, 0.2), x, y, s=0.1)
for k in range(4):
    gp.set_param("w", 0.2 + 1e-6 * (k + 1))
    gp.log_lh, gp.dloglh_dtheta, gp.mean(xo), gp.cov(xo)
torch.cuda.synchronize()
