"""Phase timing of potrf_diag_kernel (needs a library built with -DGPB_DIAG_CLK)."""
import ctypes, os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch
from gaussian_processes_b200 import _lib, engine, device as D
from conftest import synth_xy
x, y = synth_xy(1024, 0)
e = engine.Engine(engine.GAUSSIAN, (1.0, 0.5), 1.0, x, y)
for rep in range(3):
    e.reset(); e.factor(); torch.cuda.synchronize()
buf = (ctypes.c_longlong * 64)()
_lib.lib.gpb_debug_diag_clk(buf)
c = np.array(buf[:18], dtype=np.int64)
names = ["load"] + sum([["P1_%d" % b, "P2P3_%d" % b, "P4_%d" % b] for b in range(4)], []) + ["store", "P5_d1", "P5_d2", "P5_d3"]
d = np.diff(c)
for n_, v in zip(names, d):
    print("%-8s %7d cycles" % (n_, v))
print("total", c[17] - c[0])
