import os, sys, time
import numpy as np
sys.path.insert(0, "/root/repo"); sys.path.insert(0, "/root/repo/tests")
import torch
import gaussian_processes_b200 as gpb
from gaussian_processes_b200 import _lib
from conftest import synth_xy
for n in (300, 512, 1000, 1024, 2000, 2048, 4096):
    x, y = synth_xy(n, 0)
    for mode in (2, 0):
        _lib.set_option("potrf_dataflow", mode)
        gp = gpb.GP(gpb.GaussianKernel(1.0, 0.5), x, y, s=1.0)
        for k in range(5):
            gp.set_param("w", 0.5 + 1e-6 * (k + 1)); gp.dloglh_dtheta; gp.log_lh
        ts, tl = [], []
        for k in range(20):
            gp.set_param("w", 0.5 + 1e-5 * (k + 1))
            t0 = time.perf_counter(); gp.dloglh_dtheta; gp.log_lh; ts.append(time.perf_counter() - t0)
            gp.set_param("w", 0.5 - 1e-5 * (k + 1))
            t0 = time.perf_counter(); gp.log_lh; tl.append(time.perf_counter() - t0)
        print("n=%d dataflow=%s: log_lh+dloglh %.3f ms (median %.3f), log_lh only %.3f ms" % (n, mode == 0, min(ts) * 1e3, np.median(ts) * 1e3, min(tl) * 1e3), flush=True)
