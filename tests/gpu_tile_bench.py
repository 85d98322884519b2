"""Cycles per 128x128x128 tile task of the dataflow factorisation's workers (no flags, all SMs busy)."""
import ctypes, os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from gaussian_processes_b200 import _lib, device as D
grid, reps = 148, 40
A = torch.rand(3 * grid * 128, 384, dtype=torch.float64, device="cuda") * 1e-3
W = torch.tril(torch.rand(128, 128, dtype=torch.float64, device="cuda")) * 1e-2
flags = D.izeros(1024)
out = torch.zeros(2 * grid, dtype=torch.int64, device="cuda")
_lib.lib.gpb_debug_tile_bench.argtypes = [ctypes.c_void_p, ctypes.c_longlong, ctypes.c_void_p, ctypes.c_longlong, ctypes.c_int,
                                          ctypes.c_int, ctypes.c_int, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p]
for mode, name in ((0, "upd"), (1, "trsm"), (2, "upd staggered"), (3, "trsm staggered")):
    for g in (148,):
        for k in range(2):
            _lib.lib.gpb_debug_tile_bench(A.data_ptr(), 384, W.data_ptr(), 128, g, reps, mode, flags.data_ptr(), out.data_ptr(), D.stream_ptr())
            torch.cuda.synchronize()
        c = out[:2 * g].cpu().numpy() / reps
        print("%s grid %d: cycles per full tile (two half-tile tasks, one per group, concurrently) mean %.0f max %.0f  (DMMA floor 32768)" % (name, g, c.mean(), c.max()))
