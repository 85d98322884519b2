"""Kernel-class split (event-timed, one stream) of the batched evaluator at C4's shape (N=1024) and other sizes."""
import ctypes, json, os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch
from gaussian_processes_b200 import _lib, engine
from conftest import synth_xy
_lib.lib.gpb_profile_enable.argtypes = [ctypes.c_int]; _lib.lib.gpb_profile_enable.restype = None
_lib.lib.gpb_profile_read.argtypes = [ctypes.c_int, _lib.dp, ctypes.POINTER(ctypes.c_int64)]
def cand(B, seed):
    rng = np.random.RandomState(seed)
    return np.stack([rng.uniform(0.5, 2, B), rng.uniform(np.pi / 32, np.pi / 2, B), rng.uniform(0.75, 1.5, B)], axis=1)
for n, B in [(int(a.split("x")[0]), int(a.split("x")[1])) for a in (sys.argv[1:] or ["1024x512", "128x2048", "512x1024"])]:
    x, y = synth_xy(n, 0)
    out = {"n": n, "B": B}
    for streams in (0, 1):
        _lib.set_option("eval_streams", streams)
        ev = engine.BatchEvaluator(engine.GAUSSIAN, x, y, max_batch=B)
        ev.eval_device(cand(B, 1)); torch.cuda.synchronize()
        t0 = time.perf_counter()
        for k in range(3):
            ev.eval_device(cand(B, 2 + k))
        torch.cuda.synchronize()
        dt = (time.perf_counter() - t0) / 3
        out["evals_per_s_streams%d" % streams] = B / dt
        out["tflops_streams%d" % streams] = B * float(n) ** 3 / dt / 1e12
        if streams == 1:
            _lib.lib.gpb_profile_enable(1)
            ev.eval_device(cand(B, 9)); torch.cuda.synchronize()
            cl = {}
            for cid, nm in enumerate(["gemm", "diag", "build", "trsv", "reduce", "misc"]):
                m, c = ctypes.c_double(), ctypes.c_int64()
                _lib.lib.gpb_profile_read(cid, ctypes.byref(m), ctypes.byref(c))
                cl[nm] = [round(m.value, 3), c.value]
            _lib.lib.gpb_profile_enable(0)
            out["classes_ms_launches"] = cl
        del ev
    _lib.set_option("eval_streams", 0)
    print(json.dumps(out))
