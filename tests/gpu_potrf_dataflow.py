"""gpb_potrf of one matrix: the persistent dataflow launch (chain.cu) against the launch-sequence
schedule (potrf.cu) -- same factor, inverted diagonal blocks and info; time of both."""
import json, os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch
from gaussian_processes_b200 import _lib, engine, device as D
from conftest import synth_xy

sizes = [int(a) for a in sys.argv[1:] if not a.startswith(("d", "g", "s", "e", "m", "f", "u", "h", "i", "b", "x"))] or [256, 384, 1024, 2048, 4096, 8192]
for a in sys.argv[1:]:
    if a.startswith("d"):
        _lib.set_option("chain_diag", int(a[1:]))
    if a.startswith("g"):
        _lib.set_option("chain_group", int(a[1:]))
    if a.startswith("s"):
        _lib.set_option("chain_sched", int(a[1:]))
    if a.startswith("m"):
        _lib.set_option("chain_mform", int(a[1:]))
    if a.startswith("f"):
        _lib.set_option("chain_fuse", int(a[1:]))
    if a.startswith("u"):
        _lib.set_option("chain_fuse_guard", int(a[1:]))
    if a.startswith("b"):
        _lib.set_option("chain_band", int(a[1:]))
    if a.startswith("x"):
        _lib.set_option("chain_band_x", int(a[1:]))
    if a.startswith("i"):
        _lib.set_option("chain_imminent", int(a[1:]))
    if a.startswith("h"):
        _lib.set_option("chain_horizon", int(a[1:]))
    if a.startswith("e"):
        _lib.set_option("chain_express", int(a[1:]))
for nn in sizes:
    xx, yy = synth_xy(nn, 0)
    eng = engine.Engine(engine.GAUSSIAN, (1.0, 0.5), 1.0, xx, yy)
    res = {}
    for mode in (2, 0):
        _lib.set_option("potrf_dataflow", mode)
        W, V, info = D.zeros(nn, nn), D.zeros(nn, nn), D.izeros(1)
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
        res[mode] = (torch.tril(L).clone(), W.clone(), V.clone(), int(info.item()), best)
        del L, W, V
    L0, W0, V0, i0, t0 = res[2]
    L1, W1, V1, i1, t1 = res[0]
    blk = torch.zeros(nn, nn, dtype=torch.bool, device=L0.device)
    for b in range(nn // 128):
        blk[b * 128:(b + 1) * 128, b * 128:(b + 1) * 128] = True
    out = dict(n=nn, info_old=i0, info_new=i1, ms_old=round(t0, 3), ms_new=round(t1, 3),
               tflops_old=round(nn ** 3 / 3 / t0 / 1e9, 2), tflops_new=round(nn ** 3 / 3 / t1 / 1e9, 2),
               dL=float((L0 - L1).abs().max() / L0.abs().max()),
               dW=float(((W0 - W1) * blk).abs().max() / (W0 * blk).abs().max()),
               dV=float(((V0 - V1) * blk).abs().max() / (V0 * blk).abs().max()),
               nan=bool(torch.isnan(L1).any().item()))
    print(json.dumps(out), flush=True)
    del eng, res
_lib.set_option("potrf_dataflow", 0)

# phase clocks of the chain CTA for the last size
import ctypes
nn = sizes[-1]
T = nn // 128
buf = (ctypes.c_longlong * (T * 8))()
_lib.lib.gpb_debug_chain_clocks.argtypes = [ctypes.c_void_p, ctypes.POINTER(ctypes.c_longlong), ctypes.c_int]
_lib.lib.gpb_debug_chain_clocks(D.stream_ptr(), buf, T)
c = np.array(buf[:], dtype=np.int64).reshape(T, 8)
d = np.diff(c, axis=1)[1:]
names = ["wait_sub", "trsm", "publish", "wait_diag", "syrk", "diag", "publish2"]
print("chain phases (cycles, mean over steps 1..):", {n: int(v) for n, v in zip(names, d.mean(axis=0))})
print("chain phases (cycles, step 1, mid, last):", d[0].tolist(), d[len(d) // 2].tolist(), d[-1].tolist())
print("step total mean cycles", int((c[1:, 7] - c[1:, 0]).mean()), "whole", int(c[-1, 7] - c[0, 0]),
      "| pipelined group: wait for the diagonal tile %d, diagonal block %d" % ((c[1:, 5] - c[1:, 0]).mean(), (c[1:, 6] - c[1:, 5]).mean()))
nc = (1280 + 256 + 12 * 64) // 4        # 2 * 148 worker-group rows + the diagonal block's phase clocks at [320, 324)
wb = (ctypes.c_longlong * (nc * 4))()
_lib.lib.gpb_debug_chain_workers.argtypes = [ctypes.c_void_p, ctypes.POINTER(ctypes.c_longlong), ctypes.c_int]
_lib.lib.gpb_debug_chain_workers(D.stream_ptr(), wb, nc)
w = np.array(wb[:], dtype=np.int64).reshape(nc, 4)
dc = w[320:323].reshape(-1)
arr = w[328:368].reshape(-1)[:160].reshape(10, 16) - dc[0]
names = ["S0", "C0", "S1", "C1", "S2", "C2", "S3", "T1", "T2", "T3"]
for ph in range(10):
    print("  %s warp arrivals (cycles since block start):" % names[ph], arr[ph].tolist())
w5 = w.reshape(-1)[1280 + 200:1280 + 248].reshape(8, 6) - dc[0]
for ph in (2, 4, 6, 7):
    print("  warp 5 in phase %d (cycles since block start): start %d | substitution done %d | products done %d | W block stored %d | barrier with the L-column warps %d | released %d" % ((ph,) + tuple(w5[ph].tolist())))
hc = w[(1280 + 256) // 4:(1280 + 256 + 8 * T) // 4].reshape(-1)[:8 * T].reshape(T, 8)
kk = min(T - 2, max(2, T // 2))
print("helper 0 waits per step (cycles): urgent inputs", (hc[2:, 1] - hc[2:, 0]).tolist())
print("helper 0 waits per step (cycles): upd inputs   ", (hc[2:, 3] - hc[2:, 2]).tolist())
print("helper 0 waits per step (cycles): syrk inputs  ", (hc[2:, 5] - hc[2:, 4]).tolist())
print("helper 0 iteration length per step (cycles)    ", (hc[2:, 7] - hc[2:, 0]).tolist())
print("chain CTA waits per step (cycles)              ", (c[1:, 5] - c[1:, 0]).tolist())
print("helper 0 at step %d (cycles rel. to the chain CTA's step start): start %d | urgent: inputs ready %d, trsm done %d, upd inputs ready %d, upd done %d, syrk inputs ready %d, syrk done %d | end %d || chain CTA: tile ready %d, diag done %d" % (
    (kk,) + tuple((hc[kk, :8] - c[kk - 1, 0]).tolist()) + (c[kk - 1, 5] - c[kk - 1, 0], c[kk - 1, 6] - c[kk - 1, 0])))
h2 = w[(1280 + 256 + 8 * 64) // 4:(1280 + 256 + 12 * 64) // 4].reshape(-1)[:4 * T].reshape(T, 4)
print("   (global timer, ns) inverter: DIAG[k] published after the chain CTA finished diagonal block k:", (h2[1:T - 1, 2] - h2[1:T - 1, 3]).tolist())
print("   (global timer, ns) helper 0 iteration k: DIAG[k-2] seen after the chain CTA finished block k-2:", (h2[3:T, 0] - h2[1:T - 2, 3]).tolist())
w = w[:296]
w = w[w[:, 3] > 0]
print("worker groups: n=%d tasks/group mean %.1f max %d | cycles mean: wait %d trsm %d upd %d | per half-tile task busy %d | busiest group total %d, least busy %d" % (
    len(w), w[:, 3].mean(), w[:, 3].max(), w[:, 0].mean(), w[:, 1].mean(), w[:, 2].mean(),
    (w[:, 1] + w[:, 2]).sum() / w[:, 3].sum(), (w[:, 0] + w[:, 1] + w[:, 2]).max(), (w[:, 1] + w[:, 2]).min()))
