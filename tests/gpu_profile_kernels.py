"""Isolated launches of the hot kernels for `ncu --set full` (one GPU, short):
  lauum-shaped GEMM (long K, triangular), trailing SYRK (K=512, beta=1), TRSM-shaped GEMM,
  the diagonal-block kernel, the triangular solves and the fused gradient reduction."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch  # noqa: E402
import gaussian_processes_b200 as gpb  # noqa: E402
from gaussian_processes_b200 import _lib, device as D  # noqa: E402
from gaussian_processes_b200._lib import call  # noqa: E402
from conftest import synth_xy  # noqa: E402

n = int(os.environ.get("PROF_N", "4096"))
x, y = synth_xy(n, 0)
gp = gpb.GP(gpb.GaussianKernel(1.0, 0.5), x, y, s=1.0)
for rep in range(2):            # second pass is the one profiled (-s skips the first)
    gp.set_param("w", 0.5 + 0.01 * rep)
    gp.log_lh
    gp.dloglh_dtheta
torch.cuda.synchronize()
print("done", gp.log_lh)
