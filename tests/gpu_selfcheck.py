"""Primitive-by-primitive GPU diagnostic (not a pytest module): checks every CUDA
primitive against numpy on the box and keeps going after a failure, so one gpurun
round trip reports everything.  Writes gpurun_out/selfcheck.json.

usage: python tests/gpu_selfcheck.py [--quick]
"""
import json
import os
import sys
import time
import traceback

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import torch  # noqa: E402
import gaussian_processes_b200 as gpb  # noqa: E402
from gaussian_processes_b200 import _lib, device as D, engine as E  # noqa: E402
from gaussian_processes_b200._lib import call  # noqa: E402
from conftest import load_oracle, golden, synth_xy  # noqa: E402

RES = {}
oracle = load_oracle()


def rel(got, ref):
    got, ref = np.asarray(got, float), np.asarray(ref, float)
    sc = np.max(np.abs(ref)) if ref.size else 1.0
    return float(np.max(np.abs(got - ref)) / max(sc, 1e-300)) if ref.size else 0.0


def record(name, err, tol=1e-9, extra=None):
    ok = bool(err <= tol)
    RES[name] = dict(err=err, tol=tol, ok=ok, extra=extra)
    print("%-46s err=%.3e  %s %s" % (name, err, "ok" if ok else "FAIL", extra or ""), flush=True)


def section(fn):
    try:
        fn()
    except Exception as exc:   # keep going: one run must report everything
        RES[fn.__name__] = dict(ok=False, err=None, exc=repr(exc))
        print("%-46s EXCEPTION %r" % (fn.__name__, exc), flush=True)
        traceback.print_exc()


def dev(a):
    return torch.from_numpy(np.ascontiguousarray(a, dtype=np.float64)).cuda()


def gemm(A, B, C, M, N, K, alpha=1.0, beta=0.0, a_tri=0, b_tri=0, lower_only=0, Ct=None):
    call("gpb_gemm_nt", D.ptr(A), A.stride(0), D.ptr(B), B.stride(0), D.ptr(C), C.stride(0),
         D.ptr(Ct), Ct.stride(0) if Ct is not None else 0, M, N, K, alpha, beta, a_tri, b_tri, lower_only,
         D.stream_ptr())
    torch.cuda.synchronize()


# ---------------------------------------------------------------------------
def t_builders():
    rng = np.random.RandomState(1)
    for (n1, n2) in ((10, 10), (23, 16), (1, 7), (130, 67), (64, 256)):
        x1, x2 = rng.uniform(-6, 6, n1), rng.uniform(-6, 6, n2)
        kg, kp = gpb.GaussianKernel(1.3, 0.4), gpb.PeriodicKernel(0.7, 0.9, 1.7)
        for tag, k, kind in (("g", kg, oracle.GAUSSIAN), ("p", kp, oracle.PERIODIC)):
            record("build/%s/K/%dx%d" % (tag, n1, n2), rel(k(x1, x2), oracle.K(kind, x1, x2, k.params)), 1e-13)
            record("build/%s/J/%dx%d" % (tag, n1, n2), rel(k.jacobian(x1, x2), oracle.jacobian(kind, x1, x2, k.params)), 1e-13)
            record("build/%s/H/%dx%d" % (tag, n1, n2), rel(k.hessian(x1, x2), oracle.hessian(kind, x1, x2, k.params)), 1e-12)
    g = golden("kernels")
    kz = gpb.GaussianKernel(*g["gz_params"])
    K = kz(g["gz_x"], g["gz_x"])
    record("build/g/min_rule_zero_pattern", float(((K == 0) != (g["gz_K"] == 0)).sum()), 0)
    names = oracle.slice_names(oracle.PERIODIC)
    x1 = rng.uniform(-6, 6, 33)
    kp = gpb.PeriodicKernel(0.7, 0.9, 1.7)
    worst = 0.0
    for sl, nm in enumerate(names[1:], 1):
        worst = max(worst, rel(getattr(kp, nm)(x1, x1), oracle.kernel_slice(oracle.PERIODIC, sl, x1, x1, kp.params)))
    record("build/p/per_slice_methods", worst, 1e-12)


def t_gemm():
    rng = np.random.RandomState(2)
    M, N, K = 256, 384, 272
    A, B, C0 = rng.randn(M, K), rng.randn(N, K), rng.randn(M, N)
    dA, dB = dev(A), dev(B)
    dC = dev(C0)
    gemm(dA, dB, dC, M, N, K)
    record("gemm/plain", rel(dC.cpu().numpy(), A @ B.T), 1e-13)
    dC = dev(C0)
    gemm(dA, dB, dC, M, N, K, alpha=-1.0, beta=1.0)
    record("gemm/alpha_beta", rel(dC.cpu().numpy(), C0 - A @ B.T), 1e-13)
    # triangular operands
    n = 384
    Lo, Up = np.tril(rng.randn(n, n)), np.triu(rng.randn(n, n))
    Dn = rng.randn(256, n)
    dC = torch.zeros(256, n, dtype=torch.float64, device="cuda")
    gemm(dev(Dn), dev(Lo), dC, 256, n, n, b_tri=1)
    record("gemm/b_lower", rel(dC.cpu().numpy(), Dn @ Lo.T), 1e-13)
    dC = torch.zeros(n, 256, dtype=torch.float64, device="cuda")
    gemm(dev(Up), dev(Dn), dC, n, 256, n, a_tri=2)
    record("gemm/a_upper", rel(dC.cpu().numpy(), Up @ Dn.T), 1e-13)
    dC = torch.zeros(n, 256, dtype=torch.float64, device="cuda")
    gemm(dev(Lo), dev(Dn), dC, n, 256, n, a_tri=1)
    record("gemm/a_lower", rel(dC.cpu().numpy(), Lo @ Dn.T), 1e-13)
    # lower_only + mirror (lauum shape)
    dC = torch.full((n, n), 7.0, dtype=torch.float64, device="cuda")
    dU = dev(Up)
    gemm(dU, dU, dC, n, n, n, a_tri=2, b_tri=2, lower_only=1, Ct=dC)
    record("gemm/lower_only_mirror", rel(dC.cpu().numpy(), Up @ Up.T), 1e-13)
    # syrk update, lower only without mirror
    P = rng.randn(n, 128)
    S0 = rng.randn(n, n)
    dC = dev(S0)
    dP = dev(P)
    gemm(dP, dP, dC, n, n, 128, alpha=-1.0, beta=1.0, lower_only=1)
    ref = S0 - P @ P.T
    got = dC.cpu().numpy()
    record("gemm/syrk_lower", rel(np.tril(got), np.tril(ref)), 1e-13)
    record("gemm/syrk_upper_untouched", rel(np.triu(got, 128), np.triu(S0, 128)), 0)


def spd(n, seed, cond_shift=1.0):
    rng = np.random.RandomState(seed)
    x = np.sort(rng.uniform(-6, 6, n))
    K = oracle.K(oracle.GAUSSIAN, x, x, (1.0, 0.5)) + cond_shift * np.eye(n)
    return K


def t_potrf_chain():
    import scipy.linalg
    for n in (128, 256, 384, 1024):
        K = spd(n, n)
        L = dev(K)
        W, V = torch.zeros(n, n, dtype=torch.float64, device="cuda"), torch.zeros(n, n, dtype=torch.float64, device="cuda")
        info = torch.zeros(1, dtype=torch.int32, device="cuda")
        call("gpb_potrf", D.ptr(L), n, n, 0, 1, D.ptr(W), n, 0, D.ptr(V), n, 0, D.ptr(info), D.stream_ptr())
        torch.cuda.synchronize()
        Lr = np.linalg.cholesky(K)
        record("potrf/L/n%d" % n, rel(np.tril(L.cpu().numpy()), Lr), 1e-12, "info=%d" % info.item())
        Wd = W.cpu().numpy()
        worst = 0.0
        for k in range(n // 128):
            sl = slice(k * 128, (k + 1) * 128)
            worst = max(worst, rel(Wd[sl, sl], np.linalg.inv(Lr[sl, sl])))
        record("potrf/Wdiag/n%d" % n, worst, 1e-11)
        # solves
        rng = np.random.RandomState(n + 1)
        y = rng.randn(n)
        z, a = torch.zeros(n, dtype=torch.float64, device="cuda"), torch.zeros(n, dtype=torch.float64, device="cuda")
        flags = torch.zeros(2 * (n // 128) + 2, dtype=torch.int32, device="cuda")
        dy = dev(y)
        call("gpb_potrs", D.ptr(L), D.ptr(W), n, n, n, 0, 0, 1, D.ptr(dy), 0, D.ptr(z), D.ptr(a), n, D.ptr(flags), D.stream_ptr())
        torch.cuda.synchronize()
        record("potrs/z/n%d" % n, rel(z.cpu().numpy(), scipy.linalg.solve_triangular(Lr, y, lower=True)), 1e-11)
        record("potrs/alpha/n%d" % n, rel(a.cpu().numpy(), np.linalg.solve(K, y)), 1e-11)
        # trtri + lauum
        T = torch.zeros(n, n, dtype=torch.float64, device="cuda")
        call("gpb_trtri", D.ptr(L), n, n, 0, 1, D.ptr(W), n, 0, D.ptr(V), n, 0, D.ptr(T), n, 0, D.stream_ptr())
        torch.cuda.synchronize()
        Wr = np.linalg.inv(Lr)
        record("trtri/W/n%d" % n, rel(np.tril(W.cpu().numpy()), Wr), 1e-11)
        record("trtri/V/n%d" % n, rel(np.triu(V.cpu().numpy()), Wr.T), 1e-11)
        Ki = torch.zeros(n, n, dtype=torch.float64, device="cuda")
        call("gpb_lauum", D.ptr(V), n, n, 0, 1, D.ptr(Ki), n, 0, D.stream_ptr())
        torch.cuda.synchronize()
        record("lauum/Ki/n%d" % n, rel(Ki.cpu().numpy(), Wr.T @ Wr), 1e-11)
        out = torch.zeros(3, dtype=torch.float64, device="cuda")
        call("gpb_loglh", D.ptr(L), n, n, D.ptr(dy), D.ptr(a), D.ptr(info), D.ptr(out), D.stream_ptr())
        torch.cuda.synchronize()
        ld = np.linalg.slogdet(K)[1]
        ref = -0.5 * y @ np.linalg.solve(K, y) - 0.5 * ld - 0.5 * n * np.log(2 * np.pi)
        record("loglh/n%d" % n, rel(out.cpu().numpy()[0], ref), 1e-11)
    # non positive definite -> info
    K = spd(256, 5)
    K[200, 200] = -1.0
    L = dev(K)
    W = torch.zeros(256, 256, dtype=torch.float64, device="cuda")
    info = torch.zeros(1, dtype=torch.int32, device="cuda")
    call("gpb_potrf", D.ptr(L), 256, 256, 0, 1, D.ptr(W), 256, 0, None, 0, 0, D.ptr(info), D.stream_ptr())
    torch.cuda.synchronize()
    record("potrf/info_nonpd", float(abs(info.item() - 201)), 0, "info=%d" % info.item())


def t_gp_golden():
    for name in ("gp_suite_g0", "gp_suite_g1", "gp_suite_g2", "gp_suite_p0", "gp_suite_p1", "gp_suite_p2", "gp_c1"):
        g = golden(name)
        gauss = g["params"].size == 3
        k = gpb.GaussianKernel(*g["params"][:-1]) if gauss else gpb.PeriodicKernel(*g["params"][:-1])
        gp = gpb.GP(k, g["x"], g["y"], s=g["params"][-1])
        cond = np.linalg.cond(g["Kxx"])
        tol = max(1e-9, 100 * cond * 2.2e-16)
        for key in ("Kxx", "Kxx_J", "Kxx_H", "Lxx", "inv_Kxx", "inv_Kxx_y", "log_lh", "dloglh_dtheta", "dlh_dtheta", "d2lh_dtheta2"):
            record("%s/%s" % (name, key), rel(getattr(gp, key), g[key]), tol, "cond=%.1e" % cond)
        for key in ("mean", "cov", "dm_dtheta"):
            record("%s/%s" % (name, key), rel(getattr(gp, key)(g["xo"]), g[key]), tol, "cond=%.1e" % cond)
        record("%s/d2lh_norm" % name, rel(gp.d2loglh_normalised(), g["d2lh_norm"]), tol)
    for name in ("gp_g300", "gp_p257"):
        g = golden(name)
        gauss = g["params"].size == 3
        k = gpb.GaussianKernel(*g["params"][:-1]) if gauss else gpb.PeriodicKernel(*g["params"][:-1])
        gp = gpb.GP(k, g["x"], g["y"], s=g["params"][-1])
        record("%s/log_lh" % name, rel(gp.log_lh, g["log_lh"]))
        record("%s/dloglh" % name, rel(gp.dloglh_dtheta, g["dloglh_dtheta"]))
        record("%s/inv_Kxx_y" % name, rel(gp.inv_Kxx_y, g["inv_Kxx_y"]))
        record("%s/mean" % name, rel(gp.mean(g["xo"]), g["mean"]))
        c = gp.cov(g["xo"])
        record("%s/cov_diag" % name, float(np.max(np.abs(np.diag(c) - g["cov_diag"])) / np.max(np.abs(c))))
        record("%s/dm" % name, rel(gp.dm_dtheta(g["xo"]), g["dm_dtheta"]))
        record("%s/d2lh_norm" % name, rel(gp.d2loglh_normalised(), g["d2lh_norm"]))
        record("%s/Lxx_diag" % name, rel(np.diag(gp.Lxx), g["Lxx_diag"]))
        record("%s/inv_Kxx_diag" % name, rel(np.diag(gp.inv_Kxx), g["inv_Kxx_diag"]))


def t_c2(n):
    g = golden("gp_c2_n%d" % n)
    x, y = synth_xy(n, 0)
    gp = gpb.GP(gpb.GaussianKernel(*g["params"][:-1]), x, y, s=g["params"][-1])
    t0 = time.time()
    llh = gp.log_lh
    gr = gp.dloglh_dtheta
    torch.cuda.synchronize()
    dt = time.time() - t0
    record("c2_n%d/log_lh" % n, rel(llh, g["log_lh"]), 1e-9, "first eval %.1f ms" % (dt * 1e3))
    record("c2_n%d/dloglh" % n, rel(gr, g["dloglh_dtheta"]))
    record("c2_n%d/lh_is_0" % n, float(gp.lh != 0), 0)
    record("c2_n%d/inv_Kxx_y" % n, rel(gp.inv_Kxx_y, g["inv_Kxx_y"]))
    record("c2_n%d/mean" % n, rel(gp.mean(g["xo"]), g["mean"]))
    record("c2_n%d/cov" % n, rel(gp.cov(g["xo"]), g["cov"]))
    record("c2_n%d/dm" % n, rel(gp.dm_dtheta(g["xo"]), g["dm_dtheta"]))
    record("c2_n%d/d2lh_norm" % n, rel(gp.d2loglh_normalised(), g["d2lh_norm"]))
    record("c2_n%d/Lxx_diag" % n, rel(np.diag(gp.Lxx), g["Lxx_diag"]))
    record("c2_n%d/inv_Kxx_diag" % n, rel(np.diag(gp.inv_Kxx), g["inv_Kxx_diag"]))
    # batched evaluator against the scalar path
    th = np.array([g["params"], g["params"] * [1.1, 0.9, 1.05], g["params"] * [0.8, 1.2, 0.9]])
    bl, bg = gp.batch_eval(th)
    record("c2_n%d/batch_llh0" % n, rel(bl[0], g["log_lh"]))
    record("c2_n%d/batch_grad0" % n, rel(bg[0], g["dloglh_dtheta"]))
    gp2 = gpb.GP(gpb.GaussianKernel(*th[1][:-1]), x, y, s=th[1][-1])
    record("c2_n%d/batch_llh1_vs_scalar" % n, rel(bl[1], gp2.log_lh), 1e-12)
    record("c2_n%d/batch_grad1_vs_scalar" % n, rel(bg[1], gp2.dloglh_dtheta), 1e-12)


def t_invalid():
    from suite_util import INVALID_X, INVALID_Y, INVALID_H, INVALID_W
    gp = gpb.GP(gpb.GaussianKernel(INVALID_H, INVALID_W), INVALID_X, INVALID_Y, s=0)
    ok = True
    for prop in ("Lxx", "inv_Kxx", "inv_Kxx_y"):
        try:
            getattr(gp, prop)
            ok = False
        except np.linalg.LinAlgError:
            pass
    ok = ok and gp.log_lh == -np.inf and gp.lh == 0 and np.isnan(gp.dloglh_dtheta).all() \
        and np.isnan(gp.dlh_dtheta).all() and np.isnan(gp.d2lh_dtheta2).all()
    record("invalid_params_known_answer", 0.0 if ok else 1.0, 0)


def t_ext_gp_c():
    g = golden("gp_suite_g0")
    from gaussian_processes_b200.ext import gp_c
    cond = np.linalg.cond(g["Kxx"])
    tol = max(1e-9, 100 * cond * 2.2e-16)
    y, s = g["y"], float(g["params"][-1])
    record("ext.gp_c/log_lh", rel(gp_c.log_lh(y, g["Kxx"], g["inv_Kxx_y"]), g["log_lh"]), tol)
    out = np.empty(3)
    gp_c.dloglh_dtheta(y, g["inv_Kxx"], g["Kxx_J"], g["inv_Kxx_y"], s, out)
    record("ext.gp_c/dloglh", rel(out, g["dloglh_dtheta"]), tol)
    gp_c.dlh_dtheta(y, g["inv_Kxx"], g["Kxx_J"], g["inv_Kxx_y"], s, float(g["lh"]), out)
    record("ext.gp_c/dlh", rel(out, g["dlh_dtheta"]), tol)
    o2 = np.empty((3, 3))
    gp_c.d2lh_dtheta2(y, g["inv_Kxx"], g["Kxx_J"], g["Kxx_H"], g["inv_Kxx_y"], s, float(g["lh"]), g["dlh_dtheta"], o2)
    record("ext.gp_c/d2lh", rel(o2, g["d2lh_dtheta2"]), tol)
    k = gpb.GaussianKernel(*g["params"][:-1])
    dm = np.empty((3, g["xo"].size))
    gp_c.dm_dtheta(y, g["inv_Kxx"], g["Kxx_J"], k.jacobian(g["xo"], g["x"]), k(g["xo"], g["x"]), s, dm)
    record("ext.gp_c/dm", rel(dm, g["dm_dtheta"]), tol)


def t_peaks():
    import ctypes
    for use_dmma, nm in ((1, "dmma"), (0, "dfma")):
        tf, ms = ctypes.c_double(), ctypes.c_double()
        call("gpb_microbench_fp64", use_dmma, 20000, ctypes.byref(tf), ctypes.byref(ms))
        RES["peak_" + nm] = dict(tflops=tf.value, ms=ms.value, ok=True)
        print("FP64 %s peak: %.2f TFLOP/s (%.2f ms)" % (nm, tf.value, ms.value), flush=True)


def t_timing():
    for n in (1024, 4096):
        x, y = synth_xy(n, 0)
        gp = gpb.GP(gpb.GaussianKernel(1.0, 0.5), x, y, s=1.0)
        times = []
        for it in range(6):
            gp.set_param("w", 0.5 + 0.001 * it)
            torch.cuda.synchronize()
            t0 = time.time()
            gp.log_lh
            gp.dloglh_dtheta
            torch.cuda.synchronize()
            times.append(time.time() - t0)
        RES["timing_scalar_n%d" % n] = dict(ms=[t * 1e3 for t in times], ok=True)
        print("scalar eval n=%d: %s ms" % (n, ["%.2f" % (t * 1e3) for t in times]), flush=True)
        B = 8 if n == 4096 else 64
        th = np.tile([1.0, 0.5, 1.0], (B, 1)) * (1 + 0.01 * np.arange(B))[:, None]
        for it in range(3):
            torch.cuda.synchronize()
            t0 = time.time()
            gp.batch_eval(th)
            torch.cuda.synchronize()
            dt = time.time() - t0
        RES["timing_batch_n%d" % n] = dict(ms=dt * 1e3, batch=B, evals_per_s=B / dt, ok=True)
        print("batch eval n=%d B=%d: %.2f ms -> %.1f evals/s" % (n, B, dt * 1e3, B / dt), flush=True)


if __name__ == "__main__":
    quick = "--quick" in sys.argv
    print("device:", torch.cuda.get_device_name(0), flush=True)
    section(t_peaks)
    section(t_builders)
    section(t_gemm)
    section(t_potrf_chain)
    section(t_gp_golden)
    section(t_invalid)
    section(lambda: t_c2(1024))
    if not quick:
        section(lambda: t_c2(4096))
    section(t_ext_gp_c)
    section(t_timing)
    nfail = sum(1 for v in RES.values() if not v.get("ok"))
    print("SELFCHECK: %d checks, %d failed" % (len(RES), nfail), flush=True)
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    with open(os.path.join(ROOT, "gpurun_out", "selfcheck.json"), "w") as f:
        json.dump(RES, f, indent=1, default=str)
    sys.exit(1 if nfail else 0)
