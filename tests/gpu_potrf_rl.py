"""gpb_potrf alone for one matrix: left-looking vs right-looking panel steps (potrf_panel_rl)."""
import json, os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch
from gaussian_processes_b200 import _lib, engine, device as D
from conftest import synth_xy
out = {}
for nn in (512, 1024, 2048, 4096, 8192, 16384):
    xx, yy = synth_xy(nn, 0)
    eng = engine.Engine(engine.GAUSSIAN, (1.0, 0.5), 1.0, xx, yy)
    W, V, info = D.empty(nn, nn), D.empty(nn, nn), D.izeros(1)
    for rl in (2, 1):
        for inner in ((4, 8) if rl == 1 else (4,)):
            _lib.set_option("potrf_panel_rl", rl)
            _lib.set_option("potrf_inner", inner)
            best = 1e9
            for k in range(4):
                L = eng.build(eng.dx, nn, eng.dx, nn, nn, nn, 1, add_diag=True, pad_identity=True)[0]
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                torch.cuda.synchronize()
                e0.record()
                _lib.call("gpb_potrf", D.ptr(L), nn, nn, 0, 1, D.ptr(W), nn, 0, D.ptr(V), nn, 0, D.ptr(info), D.stream_ptr())
                e1.record()
                torch.cuda.synchronize()
                best = min(best, e0.elapsed_time(e1))
                if k == 0:
                    chk = float(torch.log(torch.diagonal(L)).sum().item())
                del L
            out["n%d_%s_i%d" % (nn, "rl" if rl == 1 else "ll", inner)] = dict(ms=round(best, 3), tflops=round(nn ** 3 / 3 / best / 1e9, 2), logdet_half=chk, info=int(info.item()))
    del eng, W, V
for k, v in out.items():
    print(k, json.dumps(v))
