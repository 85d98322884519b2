"""One gpb_potrf of a single N x N matrix (for an ncu launch list of the serial chain)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch
from gaussian_processes_b200 import _lib, engine, device as D
from conftest import synth_xy
nn = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
xx, yy = synth_xy(nn, 0)
eng = engine.Engine(engine.GAUSSIAN, (1.0, 0.5), 1.0, xx, yy)
W, V, info = D.empty(nn, nn), D.empty(nn, nn), D.izeros(1)
for k in range(2):
    L = eng.build(eng.dx, nn, eng.dx, nn, nn, nn, 1, add_diag=True, pad_identity=True)[0]
    _lib.call("gpb_potrf", D.ptr(L), nn, nn, 0, 1, D.ptr(W), nn, 0, D.ptr(V), nn, 0, D.ptr(info), D.stream_ptr())
    torch.cuda.synchronize()
