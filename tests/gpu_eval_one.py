"""Two cold log_lh + dloglh_dtheta evaluations of ONE GP object at N = argv[1] (for an ncu launch list)."""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch
import gaussian_processes_b200 as gpb
from conftest import synth_xy
n = int(sys.argv[1]) if len(sys.argv) > 1 else 1000
x, y = synth_xy(n, 0)
gp = gpb.GP(gpb.GaussianKernel(1.0, 0.5), x, y, s=1.0)
for k in range(2):
    gp.set_param("w", 0.5 + 1e-7 * (k + 1)); gp.log_lh; gp.dloglh_dtheta
torch.cuda.synchronize()
