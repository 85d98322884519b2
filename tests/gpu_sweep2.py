"""Second knob sweep (after the TMA GEMM): batch size x eval streams x Cholesky panel width.
usage: python tests/gpu_sweep2.py -> gpurun_out/sweep2.json"""
import itertools, json, os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch
from gaussian_processes_b200 import _lib, engine
from conftest import synth_xy

def timed(fn, reps=3):
    fn(); fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps

out = []
base = np.array([1.0, 0.5, 1.0])
for n, batches in ((4096, (8, 16, 32)), (1024, (256,))):
    x, y = synth_xy(n, 0)
    for B, streams, inner in itertools.product(batches, (2, 4, 8), (2, 4, 8)):
        _lib.set_option("eval_streams", streams)
        _lib.set_option("potrf_inner", inner)
        ev = engine.BatchEvaluator(engine.GAUSSIAN, x, y, max_batch=B)
        rng = np.random.RandomState(B)
        th = base * (1 + 0.05 * rng.uniform(-1, 1, (B, 3)))
        ms = timed(lambda: ev.eval_device(th))
        row = dict(n=n, B=B, streams=streams, inner=inner, ms=ms, evals_per_s=B / ms * 1e3)
        print(json.dumps(row), flush=True)
        out.append(row)
        del ev
json.dump(out, open(os.path.join(ROOT, "gpurun_out", "sweep2.json"), "w"), indent=1)
