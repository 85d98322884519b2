import json, os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch
import gaussian_processes_b200 as gpb
from conftest import synth_xy
for n in [int(a) for a in sys.argv[1:]] or [768, 896, 1000, 1024, 1152, 1536]:
    x, y = synth_xy(n, 0)
    gp = gpb.GP(gpb.GaussianKernel(1.0, 0.5), x, y, s=1.0)
    for k in range(5):
        gp.set_param("w", 0.5 + 1e-7 * (k + 1)); gp.log_lh; gp.dloglh_dtheta
    torch.cuda.synchronize()
    ta = tb = 0.0
    reps = 50
    for k in range(reps):
        gp.set_param("w", 0.5 + 1e-7 * (k + 10))
        t0 = time.perf_counter(); gp.log_lh; t1 = time.perf_counter(); gp.dloglh_dtheta; t2 = time.perf_counter()
        ta += t1 - t0; tb += t2 - t1
    print(json.dumps(dict(n=n, log_lh_ms=round(ta / reps * 1e3, 3), dloglh_ms=round(tb / reps * 1e3, 3))), flush=True)
