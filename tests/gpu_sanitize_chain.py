"""compute-sanitizer workload for the persistent dataflow Cholesky (csrc/chain.cu): one GP object above one block,
every scheduling variant once (small N: the sanitizer slows the spin loops by orders of magnitude)."""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch
import gaussian_processes_b200 as gpb
from gaussian_processes_b200 import _lib
from conftest import synth_xy
variants = [{}, {"chain_sched": 1}, {"chain_mform": 2}, {"chain_mform": 4}, {"chain_fuse": 1}, {"chain_group": 8}]
for n in (300, 700, 1100):
    x, y = synth_xy(n, n)
    for opts in variants:
        for k, v in opts.items():
            _lib.set_option(k, v)
        gp = gpb.GP(gpb.GaussianKernel(1.1, 0.4), x, y, s=0.7)
        r = (gp.log_lh, gp.dloglh_dtheta)
        for k in opts:
            _lib.set_option(k, 0)
        print(n, opts, float(r[0]), flush=True)
torch.cuda.synchronize()
print("done")
