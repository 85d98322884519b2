"""One GP object at N (default 4096): cold log_lh + dloglh_dtheta twice -- for an ncu launch list of the single-object path."""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch
import gaussian_processes_b200 as gpb
from conftest import synth_xy
n = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
x, y = synth_xy(n, 0)
gp = gpb.GP(gpb.GaussianKernel(1.0, 0.5), x, y, s=1.0)
for k in range(2):
    gp.set_param("w", 0.5 - 1e-4 * (k + 1))
    g = gp.dloglh_dtheta; l = gp.log_lh
torch.cuda.synchronize()
print(l, g)
