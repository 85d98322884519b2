"""BASELINE.json configs C1 - C5 on the GPU(s): timing + parity / size-independent checks.
Single process:  python tests/gpu_configs.py [c1 c3 c4 c5]
Multi GPU:       torchrun --nproc-per-node G tests/gpu_configs.py c3 c4      (test points / candidates shard)
Prints one JSON line per config (rank 0) and appends them to gpurun_out/configs.jsonl."""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

WORLD = int(os.environ.get("WORLD_SIZE", "1"))
RANK = int(os.environ.get("RANK", "0"))
LOCAL = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(LOCAL)
if WORLD > 1:
    dist.init_process_group("nccl", device_id=torch.device("cuda", LOCAL))
import gaussian_processes_b200 as gpb  # noqa: E402
from gaussian_processes_b200 import device as D  # noqa: E402
from gaussian_processes_b200.mlii import shard_bounds  # noqa: E402
from conftest import load_oracle, golden, synth_xy  # noqa: E402


def sync():
    if WORLD > 1:
        dist.barrier()
    torch.cuda.synchronize()


def maxtime(t):
    v = torch.tensor([t], dtype=torch.float64, device="cuda")
    if WORLD > 1:
        dist.all_reduce(v, op=dist.ReduceOp.MAX)
    return float(v.item())


def emit(d):
    d["n_gpus"] = WORLD
    if RANK == 0:
        print(json.dumps(d), flush=True)
        os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
        with open(os.path.join(ROOT, "gpurun_out", "configs.jsonl"), "a") as f:
            f.write(json.dumps(d) + "\n")


def rel(a, b):
    a, b = np.asarray(a, float), np.asarray(b, float)
    return float(np.max(np.abs(a - b)) / max(np.max(np.abs(b)), 1e-300))


def c1():
    """N=50, Gaussian(1, 0.2), s=0: log_lh + dloglh_dtheta + mean/cov at 100 test points."""
    g = golden("gp_c1")
    x, y, xo = g["x"], g["y"], g["xo"]
    gp = gpb.GP(gpb.GaussianKernel(1.0, 0.2), x, y, s=0)
    errs = dict(log_lh=rel(gp.log_lh, g["log_lh"]), dloglh=rel(gp.dloglh_dtheta, g["dloglh_dtheta"]),
                mean=rel(gp.mean(xo), g["mean"]), cov=rel(gp.cov(xo), g["cov"]))
    reps = 50
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for k in range(reps):
        gp.set_param("w", 0.2 + 1e-6 * (k + 1))
        gp.log_lh, gp.dloglh_dtheta, gp.mean(xo), gp.cov(xo)
    torch.cuda.synchronize()
    dt = (time.perf_counter() - t0) / reps
    # the reference's own path (oracle/_ref + scipy/numpy) on the host cores, same bundle
    oracle = load_oracle()
    impl = "ref" if oracle.have_ref() else "c"
    t0 = time.perf_counter()
    for k in range(reps):
        o = oracle.OracleGP(oracle.GAUSSIAN, (1.0, 0.2 + 1e-6 * (k + 1)), x, y, 0.0, impl)
        o.log_lh, o.dloglh_dtheta, o.mean(xo), o.cov(xo)
    dtc = (time.perf_counter() - t0) / reps
    emit(dict(config="C1", metric="bundles/s (log_lh+grad+mean+cov, N=50, M=100)", value=1 / dt,
              ms_per_bundle=dt * 1e3, parity_vs_reference=errs, cpu_reference_ms_per_bundle=dtc * 1e3,
              note="launch-latency regime: ~40 dependent launches and 6 host read-backs per bundle"))


def c2():
    """Gaussian N=4096, ONE GP through the public API (cold cache): Kxx build + Cholesky + log_lh /
    dloglh_dtheta / d2lh_dtheta2; parity against the golden vectors of the unmodified reference;
    plus the factorisation alone and the batched evaluator at several batch sizes."""
    if RANK != 0:
        return
    from gaussian_processes_b200 import engine
    g = golden("gp_c2_n4096")
    n = 4096
    x, y = synth_xy(n, 0)
    out = dict(config="C2", n=n)

    def fresh():
        return gpb.GP(gpb.GaussianKernel(1.0, 0.5), x, y, s=1.0)
    gp = fresh()
    gp.log_lh, gp.dloglh_dtheta, gp.d2loglh_normalised()          # warm (allocator, module load)
    out["parity_vs_reference"] = dict(
        log_lh=rel(gp.log_lh, g["log_lh"]), dloglh=rel(gp.dloglh_dtheta, g["dloglh_dtheta"]),
        d2lh_norm=rel(gp.d2loglh_normalised(), g["d2lh_norm"]), lh=float(gp.lh),
        d2lh_is_zero=bool(np.all(gp.d2lh_dtheta2 == 0)), alpha=rel(gp.inv_Kxx_y, g["inv_Kxx_y"]),
        inv_Kxx_diag=rel(np.diag(gp.inv_Kxx), g["inv_Kxx_diag"]), Lxx_diag=rel(np.diag(gp.Lxx), g["Lxx_diag"]),
        mean=rel(gp.mean(g["xo"]), g["mean"]), cov=rel(gp.cov(g["xo"]), g["cov"]),
        dm=rel(gp.dm_dtheta(g["xo"]), g["dm_dtheta"]))
    reps = 5
    lat = {}
    for name, fn in (("log_lh", lambda q: q.log_lh),
                     ("log_lh+dloglh", lambda q: (q.log_lh, q.dloglh_dtheta)),
                     ("log_lh+dloglh+d2lh", lambda q: (q.log_lh, q.dloglh_dtheta, q.d2loglh_normalised()))):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for k in range(reps):
            gp.set_param("w", 0.5 + 1e-7 * (k + 1))               # setters clear the memo (gp.py:231-240)
            fn(gp)
        torch.cuda.synchronize()
        lat[name + "_ms"] = (time.perf_counter() - t0) / reps * 1e3
    out["single_gp_cold_latency"] = lat
    # factorisation alone, device-timed
    e = gp._engine()
    for nm, nn in (("potrf_n4096", 4096), ("potrf_n8192", 8192), ("potrf_n16384", 16384)):
        xx, yy = synth_xy(nn, 0)
        eng = engine.Engine(engine.GAUSSIAN, (1.0, 0.5), 1.0, xx, yy)
        eng.factor()
        best = 1e9
        for k in range(3):
            eng.reset()
            L = eng.build(eng.dx, nn, eng.dx, nn, nn, nn, 1, add_diag=True, pad_identity=True)[0]
            W, V, info = D.empty(nn, nn), D.empty(nn, nn), D.izeros(1)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            torch.cuda.synchronize()
            e0.record()
            gpb._lib.call("gpb_potrf", D.ptr(L), nn, nn, 0, 1, D.ptr(W), nn, 0, D.ptr(V), nn, 0, D.ptr(info), D.stream_ptr())
            e1.record()
            torch.cuda.synchronize()
            best = min(best, e0.elapsed_time(e1))
        out[nm] = dict(ms=best, tflops=nn ** 3 / 3 / best / 1e9)
        del eng, L, W, V
    # batched evaluator vs batch size
    bs = {}
    for B in (1, 2, 4, 8, 16, 32):
        ev = engine.BatchEvaluator(engine.GAUSSIAN, x, y, max_batch=B)
        rng = np.random.RandomState(B)
        th = np.array([1.0, 0.5, 1.0]) * (1 + 0.05 * rng.uniform(-1, 1, (B, 3)))
        ev.eval_device(th); ev.eval_device(th)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for k in range(3):
            ev.eval_device(th + 1e-6 * k)
        e1.record()
        torch.cuda.synchronize()
        bs[str(B)] = 3 * B / (e0.elapsed_time(e1) * 1e-3)
        del ev
    out["batched_evals_per_s_by_batch"] = bs
    emit(out)


def c3():
    """Periodic N=8192: Kxx + jacobian/hessian builds, posterior mean/cov at M=16384, test points sharded."""
    n, m = 8192, 16384
    x, y = synth_xy(n, 0)
    xo = np.linspace(-2 * np.pi, 2 * np.pi, m)
    gp = gpb.GP(gpb.PeriodicKernel(1.0, 1.0, 1.0), x, y, s=1.0)
    e = gp._engine()
    out = dict(config="C3", n=n, m=m)
    # builders on the device (HBM-write roofline): K, jacobian (3 slices), hessian (9 slices), all 13 fused
    for name, mask, ns in (("K", 0x1, 1), ("jacobian", 0xE, 3), ("hessian", 0x1FF0, 9), ("all13", 0x1FFF, 13)):
        buf = D.empty(ns, e.npad, e.npad)
        e.build(e.dx, n, e.dx, n, e.npad, e.npad, mask, out=buf)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        e.build(e.dx, n, e.dx, n, e.npad, e.npad, mask, out=buf)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1)
        out["build_%s_ms" % name] = ms
        out["build_%s_GBps" % name] = ns * n * n * 8 / ms / 1e6
        del buf
    # factorisation + inverse (every rank, redundantly -- SURVEY 8e)
    sync()
    t0 = time.perf_counter()
    llh = gp.log_lh
    e.Ki()
    torch.cuda.synchronize()
    out["factor_inverse_ms"] = (time.perf_counter() - t0) * 1e3
    out["log_lh"] = float(llh)
    out["solve_residual"] = e.solve_residual()
    # sharded posterior: rank r owns test points [lo, hi)
    lo, hi = shard_bounds(m, WORLD, RANK)
    gp.mean(xo[lo:hi]); gp.cov_rows(xo[:256], 0, 128)      # warm
    sync()
    t0 = time.perf_counter()
    mean_r = gp.mean(xo[lo:hi])
    torch.cuda.synchronize()
    tm = maxtime(time.perf_counter() - t0)
    tcs = []
    cov_r = None
    for rep in range(3):             # rep 0 pays the one-time page-locking of the result buffer
        del cov_r
        sync()
        t0 = time.perf_counter()
        cov_r = gp.cov_rows(xo, lo, hi)
        torch.cuda.synchronize()
        tcs.append(maxtime(time.perf_counter() - t0))
    tc = min(tcs[1:])
    sync()
    t0 = time.perf_counter()
    dev = e.cov_rows(xo, lo, hi, host=False)
    torch.cuda.synchronize()
    tdev = maxtime(time.perf_counter() - t0)
    del dev
    out.update(mean_test_pts_per_s=m / tm, cov_test_pts_per_s=m / tc, mean_ms=tm * 1e3, cov_rows_ms=tc * 1e3,
               cov_rows_first_call_ms=tcs[0] * 1e3, cov_rows_device_only_ms=tdev * 1e3,
               cov_result_GB=cov_r.nbytes / 1e9,
               cov_includes="Kxox/Kxoxo builds + GEMMs + D2H of the [m_r, M] row block into a numpy array "
                            "(page-locked result pool, panel downloads overlapped with the GEMMs)")
    # checks: symmetric block, agreement with the Z Z^T path on a sub-block, oracle on a small sub-problem
    if RANK == 0:
        sub = gp.cov(xo[lo:lo + 300])
        out["cov_rows_vs_cov_block"] = rel(cov_r[:300, lo:lo + 300], sub)
        oracle = load_oracle()
        ns = 1024
        xs, ys = x[::8][:ns], y[::8][:ns]
        gps = gpb.GP(gpb.PeriodicKernel(1.0, 1.0, 1.0), xs, ys, s=1.0)
        o = oracle.OracleGP(oracle.PERIODIC, (1.0, 1.0, 1.0), xs, ys, 1.0)
        out["parity_subproblem_n1024"] = dict(mean=rel(gps.mean(xo[:200]), o.mean(xo[:200])),
                                              cov_rows=rel(gps.cov_rows(xo[:200], 50, 120), o.cov(xo[:200])[50:120]),
                                              log_lh=rel(gps.log_lh, o.log_lh), dloglh=rel(gps.dloglh_dtheta, o.dloglh_dtheta))
    emit(out)


def c4():
    """Batched fit_MLII: 4096 restarts x N=1024 Gaussian, candidates sharded, NCCL argmax gather."""
    n, B = 1024, 4096
    x, y = synth_xy(n, 0)
    rng = np.random.RandomState(4)
    cand = np.stack([rng.uniform(0.5, 2, B), rng.uniform(np.pi / 32, np.pi / 2, B), rng.uniform(0.75, 1.5, B)], axis=1)
    gp = gpb.GP(gpb.GaussianKernel(1.0, 1.0), x, y, s=1.0)
    gp.fit_MLII(cand[:64 * WORLD], set_params=False)          # warm (workspace, NCCL)
    sync()
    t0 = time.perf_counter()
    res = gp.fit_MLII(cand, set_params=False)
    torch.cuda.synchronize()
    dt_first = maxtime(time.perf_counter() - t0)              # includes growing the workspace to the full chunk size
    sync()
    t0 = time.perf_counter()
    res = gp.fit_MLII(cand, set_params=False)
    torch.cuda.synchronize()
    dt = maxtime(time.perf_counter() - t0)
    out = dict(config="C4", n=n, restarts=B, value=B / dt, metric="candidate evals/s (log_lh+grad), whole job",
               seconds=dt, first_call_seconds=dt_first, first_call_value=B / dt_first,
               best_index=res.best_index, best_log_lh=float(res.best_log_lh))
    if RANK == 0:
        oracle = load_oracle()
        idx = [res.best_index, 0, 1, B - 1]
        errs = []
        for i in idx:
            o = oracle.OracleGP(oracle.GAUSSIAN, cand[i, :2], x, y, cand[i, 2])
            errs.append((rel(res.log_lh[i], o.log_lh), rel(res.dloglh_dtheta[i], o.dloglh_dtheta)))
        out["parity_spot_checks"] = errs
        out["argmax_consistent"] = bool(res.best_index == int(np.argmax(np.where(np.isnan(res.log_lh), -np.inf, res.log_lh))))
    emit(out)


def c5():
    """N=32768 fp64 (8.6 GB Kxx): single-GPU Cholesky + cho_solve + full posterior cov at M=8192."""
    if RANK != 0:
        return
    n, m = 32768, 8192
    x, y = synth_xy(n, 0)
    xo = np.linspace(-2 * np.pi, 2 * np.pi, m)
    gp = gpb.GP(gpb.GaussianKernel(1.0, 0.5), x, y, s=1.0)
    e = gp._engine()
    out = dict(config="C5", n=n, m=m)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    info = e.factor()
    torch.cuda.synchronize()
    t1 = time.perf_counter()
    e.alpha()
    llh = gp.log_lh
    torch.cuda.synchronize()
    t2 = time.perf_counter()
    out.update(info=info, potrf_ms=(t1 - t0) * 1e3, potrf_tflops=n ** 3 / 3 / (t1 - t0) / 1e12,
               solve_loglh_ms=(t2 - t1) * 1e3, log_lh=float(llh), solve_residual=e.solve_residual())
    t3s = []
    cov = None
    for rep in range(2):
        del cov
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        cov = gp.cov(xo)
        torch.cuda.synchronize()
        t3s.append(time.perf_counter() - t0)
    t3 = t3s[1]
    t0 = time.perf_counter()
    dev = e.cov(xo, host=False)
    torch.cuda.synchronize()
    tdev = time.perf_counter() - t0
    del dev
    out.update(cov_ms=t3 * 1e3, cov_first_call_ms=t3s[0] * 1e3, cov_device_only_ms=tdev * 1e3, cov_test_pts_per_s=m / t3,
               cov_includes="first call: trtri + Z=Kxox L^-T + Kxoxo - Z Z^T + D2H of 0.5 GB; later calls reuse L^-1")
    t0 = time.perf_counter()
    mean = gp.mean(xo)
    torch.cuda.synchronize()
    out["mean_ms"] = (time.perf_counter() - t0) * 1e3
    # independent check of a few covariance entries: cho_solve path (substitution) instead of L^-1
    idx = [0, 1234, m - 1]
    kv = gpb.GaussianKernel(1.0, 0.5)(x, xo[idx])                    # [n, 3]
    worst = 0.0
    for c, j in enumerate(idx):
        sol = D.to_host(gp._engine().solve(kv[:, c])[:n]).copy()
        ref_col = gpb.GaussianKernel(1.0, 0.5)(xo, xo[j:j + 1])[:, 0] - kv.T[c] @ sol * 0 - (gpb.GaussianKernel(1.0, 0.5)(xo, x) @ sol)
        worst = max(worst, rel(cov[:, j], ref_col))
    out["cov_columns_vs_cho_solve_path"] = worst
    out["cov_symmetry"] = float(np.max(np.abs(cov - cov.T)))
    out["cov_diag_min"] = float(np.min(np.diag(cov)))
    out["peak_mem_GB"] = torch.cuda.max_memory_allocated() / 1e9
    emit(out)


if __name__ == "__main__":
    which = [a for a in sys.argv[1:] if a in ("c1", "c2", "c3", "c4", "c5")] or ["c1", "c2", "c3", "c4", "c5"]
    for w in which:
        try:
            globals()[w]()
        except Exception as exc:
            import traceback
            traceback.print_exc()
            emit(dict(config=w.upper(), error=repr(exc)))
    if WORLD > 1:
        dist.barrier()
        dist.destroy_process_group()
