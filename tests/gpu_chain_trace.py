"""Critical-path trace of one dataflow factorisation (gpb_debug_chain_tiles): per helper iteration k, when the last
worker-side update of the helper's input tiles (k,k-2) and (k,k-1) became runnable, started and finished, relative to
the moment the chain CTA started diagonal block k-2.  Usage: python tests/gpu_chain_trace.py [N]"""
import ctypes, os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch
from gaussian_processes_b200 import _lib, engine, device as D
from conftest import synth_xy

n = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
for a in sys.argv[2:]:
    k_, v_ = a.split("=")
    _lib.set_option(k_, int(v_))
T = n // 128
x, y = synth_xy(n, 0)
eng = engine.Engine(engine.GAUSSIAN, (1.0, 0.5), 1.0, x, y)
W, V, info = D.zeros(n, n), D.zeros(n, n), D.izeros(1)
for rep in range(3):
    L = eng.build(eng.dx, n, eng.dx, n, n, n, 1, add_diag=True, pad_identity=True)[0]
    torch.cuda.synchronize()
    _lib.call("gpb_potrf", D.ptr(L), n, n, 0, 1, D.ptr(W), n, 0, D.ptr(V), n, 0, D.ptr(info), D.stream_ptr())
    torch.cuda.synchronize()
nl = 2 * T * T * 4 + T * 4
buf = (ctypes.c_longlong * nl)()
_lib.lib.gpb_debug_chain_tiles.argtypes = [ctypes.c_void_p, ctypes.POINTER(ctypes.c_longlong), ctypes.c_longlong]
_lib.lib.gpb_debug_chain_tiles(D.stream_ptr(), buf, nl)
raw = np.array(buf[:], dtype=np.int64)
tile = raw[:2 * T * T * 4].reshape(T, 2, T, 4)      # [i][h][j][s]
step = raw[2 * T * T * 4:].reshape(T, 4)            # DIAG | M1 | M2 | block start
t0 = step[0, 3]
us = lambda v: (v - t0) / 1000.0
print("diagonal block starts (us):", [round(us(v), 1) for v in step[:, 3]])
print("step lengths (us):", [round((step[k + 1, 3] - step[k, 3]) / 1000.0, 1) for k in range(T - 1)])
print("k : block(k-2) start | DIAG[k-2] | M1[k-3] | tile (k,k-3) complete h0,h1 | tile (k,k-2): runnable, start, done (h0 ; h1) | "
      "tile (k,k-1): runnable, start, done (h0 ; h1)      [us after block k-2 started]")
for k in range(4, T):
    b = step[k - 2, 3]
    r = lambda v: round((v - b) / 1000.0, 1)
    row = ["%2d" % k, round(us(b), 1), r(step[k - 2, 0]), r(step[k - 3, 1])]
    row.append((r(tile[k, 0, k - 3, 1]), r(tile[k, 1, k - 3, 1])))
    for jj, slot in ((k - 2, 1), (k - 1, 2)):
        cells = []
        for h in range(2):
            dep = max(step[k - 3, slot], tile[k, h, k - 3, 1])
            cells.append((r(dep), r(tile[k, h, jj, 0]), r(tile[k, h, jj, 1])))
        row.append(cells)
    print(*row)
