"""One warm + one measured batched evaluation (B=8, N=4096) on a single stream, for an ncu launch list."""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch
from gaussian_processes_b200 import _lib, engine
from conftest import synth_xy
n = int(os.environ.get("STEP_N", "4096")); B = int(os.environ.get("STEP_B", "8"))
_lib.set_option("eval_streams", int(os.environ.get("STEP_STREAMS", "1")))
x, y = synth_xy(n, 0)
ev = engine.BatchEvaluator(engine.GAUSSIAN, x, y, max_batch=B)
th = np.tile([1.0, 0.5, 1.0], (B, 1)) * (1 + 0.002 * np.arange(B))[:, None]
for rep in range(2):
    ev.eval_device(th)
    torch.cuda.synchronize()
print("ok")
