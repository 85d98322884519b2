"""Cold log_lh + dloglh_dtheta latency of ONE GP object across sizes, with the reference's path
(oracle/_ref + scipy/numpy) on the host cores beside it."""
import json, os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch
import gaussian_processes_b200 as gpb
from conftest import synth_xy, load_oracle
oracle = load_oracle()
impl = "ref" if oracle.have_ref() else "c"
out = []
for n in (16, 50, 128, 129, 300, 512, 1000, 2000, 4096):
    x, y = synth_xy(n, 0)
    gp = gpb.GP(gpb.GaussianKernel(1.0, 0.5), x, y, s=1.0)
    reps = 200 if n <= 512 else 60
    for k in range(20):
        gp.set_param("w", 0.5 + 1e-7 * (k + 1)); gp.log_lh; gp.dloglh_dtheta
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for k in range(reps):
        gp.set_param("w", 0.5 + 1e-7 * (k + 10)); gp.log_lh; gp.dloglh_dtheta
    dt = (time.perf_counter() - t0) / reps
    creps = 50 if n <= 512 else (5 if n <= 2000 else 1)
    t0 = time.perf_counter()
    for k in range(creps):
        o = oracle.OracleGP(oracle.GAUSSIAN, (1.0, 0.5 + 1e-7 * (k + 1)), x, y, 1.0, impl)
        o.log_lh; o.dloglh_dtheta
    dtc = (time.perf_counter() - t0) / creps
    out.append(dict(n=n, ms=round(dt * 1e3, 4), cpu_reference_ms=round(dtc * 1e3, 3), speedup=round(dtc / dt, 1)))
    print(json.dumps(out[-1]), flush=True)
