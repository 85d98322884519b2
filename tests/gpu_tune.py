"""Sweep the tuning knobs (gemm tile, Cholesky panel width, eval streams) on the GPU:
throughput of the batched evaluator at N=4096 / N=1024 and the scalar-object latency,
with a parity check against the golden value under every configuration.
usage: python tests/gpu_tune.py  -> gpurun_out/tune.json"""
import itertools
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch  # noqa: E402
import gaussian_processes_b200 as gpb  # noqa: E402
from gaussian_processes_b200 import _lib, engine  # noqa: E402
from conftest import golden, synth_xy  # noqa: E402


def timed(fn, reps=3):
    fn()
    torch.cuda.synchronize()
    best = 1e9
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    return best


def main():
    out = []
    g4 = golden("gp_c2_n4096")
    x4, y4 = synth_xy(4096, 0)
    x1, y1 = synth_xy(1024, 0)
    g1 = golden("gp_c2_n1024")
    base = np.array([1.0, 0.5, 1.0])
    sweeps = list(itertools.product((64, 128), (1, 2, 4), (1, 2, 4)))
    if "--quick" in sys.argv:
        sweeps = [(64, 2, 4), (128, 1, 1)]
    for bm, inner, streams in sweeps:
        _lib.set_option("gemm_bm", bm)
        _lib.set_option("potrf_inner", inner)
        _lib.set_option("eval_streams", streams)
        row = dict(gemm_bm=bm, potrf_inner=inner, eval_streams=streams)
        for n, x, y, g, B in ((4096, x4, y4, g4, 8), (1024, x1, y1, g1, 64)):
            ev = engine.BatchEvaluator(engine.GAUSSIAN, x, y, max_batch=B)
            th = np.tile(base, (B, 1)) * (1 + 0.002 * np.arange(B))[:, None]
            th[0] = base
            llh, grad, info = ev.eval(th)
            row["err_llh_n%d" % n] = abs(llh[0] - float(g["log_lh"])) / abs(float(g["log_lh"]))
            row["err_grad_n%d" % n] = float(np.max(np.abs(grad[0] - g["dloglh_dtheta"])) / np.max(np.abs(g["dloglh_dtheta"])))
            ms = timed(lambda: ev.eval_device(th))
            row["batch_ms_n%d" % n] = ms
            row["evals_per_s_n%d" % n] = B / ms * 1e3
            del ev
        gp = gpb.GP(gpb.GaussianKernel(1.0, 0.5), x4, y4, s=1.0)
        ts = []
        for it in range(4):
            gp.set_param("w", 0.5 + 0.001 * it)
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            gp.log_lh
            gp.dloglh_dtheta
            torch.cuda.synchronize()
            ts.append((time.perf_counter() - t0) * 1e3)
        row["scalar_ms_n4096"] = min(ts)
        print(json.dumps(row), flush=True)
        out.append(row)
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    json.dump(out, open(os.path.join(ROOT, "gpurun_out", "tune.json"), "w"), indent=1)


if __name__ == "__main__":
    main()
