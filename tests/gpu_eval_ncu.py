"""One batched evaluation (B candidates, N=4096) for targeted ncu captures of the non-GEMM kernels."""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch
from gaussian_processes_b200 import _lib, engine
from conftest import synth_xy
B = int(sys.argv[1]) if len(sys.argv) > 1 else 8
x, y = synth_xy(4096, 0)
_lib.set_option("eval_streams", 1)
ev = engine.BatchEvaluator(engine.GAUSSIAN, x, y, max_batch=B)
rng = np.random.RandomState(1)
th = np.stack([rng.uniform(0.9, 1.1, B), rng.uniform(0.45, 0.55, B), rng.uniform(0.9, 1.1, B)], axis=1)
ev.eval_device(th); torch.cuda.synchronize()
