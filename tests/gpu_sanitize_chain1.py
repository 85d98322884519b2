import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch
import gaussian_processes_b200 as gpb
from gaussian_processes_b200 import _lib
from conftest import synth_xy
n = int(sys.argv[1]) if len(sys.argv) > 1 else 700
for a in sys.argv[2:]:
    k_, v_ = a.split("=")
    _lib.set_option(k_, int(v_))
x, y = synth_xy(n, n)
gp = gpb.GP(gpb.GaussianKernel(1.1, 0.4), x, y, s=0.7)
print(gp.log_lh)
torch.cuda.synchronize()
