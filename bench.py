#!/usr/bin/env python
"""bench.py -- headline metric of BASELINE.json on B200:
log_lh + dloglh_dtheta evaluations / second at N = 4096, fp64 (config C2).

One *step* = one pass of the hot path over one batch of B hyperparameter candidates on a
fixed synthetic 1-D data set (x = sort U(-2pi, 2pi), y = sin x + 0.1 N(0,1), seed 0; SURVEY
8d): Kxx build -> blocked Cholesky -> solves -> log_lh -> inverse -> fused gradient, for
every candidate (what one iteration of a multi-restart MLII search costs).  Every step uses
different candidates, so nothing can be reused between steps.

  python bench.py --gpus N --steps K --warmup W            (ours; torchrun for N > 1)
  python bench.py --impl reference --steps K --warmup W    (the reference's CPU path)

`value`  : whole-job evals/s with x, y resident in HBM, timed with CUDA events on the launch
           stream, max over ranks.
`e2e`    : the same metric through the public Python API (GP.batch_eval) with HOST x, y,
           thetas and results every step (pinned staging, H2D + D2H inside the timed region).
`roofline`: the DMMA GEMM kernel (all O(N^3) work): algorithmic flops / its event-timed
           duration, against the FP64 DMMA issue rate measured live on the same GPU.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

if "reference" in sys.argv[1:]:
    # torch.distributed.run exports OMP_NUM_THREADS=1 to every worker; the reference arm is a CPU
    # measurement and must use the host's cores whatever launched it (read before numpy loads OpenBLAS)
    for _v in ("OMP_NUM_THREADS", "OPENBLAS_NUM_THREADS", "MKL_NUM_THREADS"):
        os.environ.pop(_v, None)

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

N_DEFAULT = 4096
BASE_THETA = np.array([1.0, 0.5, 1.0])      # Gaussian(h=1, w=0.5), s=1


def synth_xy(n, seed=0):
    rng = np.random.RandomState(seed)
    x = np.sort(rng.uniform(-2 * np.pi, 2 * np.pi, n))
    y = np.sin(x) + 0.1 * rng.randn(n)
    return x, y


def candidates(step, rank, batch):
    """Deterministic, step- and rank-dependent candidates around the C2 parameters
    (h in [0.9,1.1], w in [0.45,0.55], s in [0.95,1.05]: logdet stays above MIN)."""
    rng = np.random.RandomState(1000003 * rank + step)
    return BASE_THETA * (1.0 + 0.1 * rng.uniform(-1, 1, (batch, 3))) * [1.0, 1.0, 0.5] + [0, 0, 0.5]


class ClockSampler(object):
    """nvidia-smi clocks / throttle reasons DURING the timed region (B200_PROFILING.md)."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q, "--format=csv,noheader,nounits",
                 "-lms", "100"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, smax, reasons, power = [], [], set(), []
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            try:
                sm.append(float(r[0])); smax.append(float(r[1])); power.append(float(r[2]))
                for nm, v in zip(names, r[3:7]):
                    if v.lower().startswith("active"):
                        reasons.add(nm)
            except Exception:
                continue
        busy = [c for c, p in zip(sm, power) if p > 300] or sm
        return {"sm_mhz": float(np.median(busy)) if busy else None,
                "sm_max_mhz": max(smax) if smax else None, "reasons": sorted(reasons),
                "samples": len(sm), "power_w_max": max(power) if power else None}


def hbm_peak():
    """(GB/s, source): the driver-measured copy bandwidth when MEASURED_PEAKS.json is present, else the
    fallback of /opt/skills/guides/B200_PROFILING.md."""
    try:
        d = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        return float(d["hbm_gbs"]), "MEASURED_PEAKS.json (driver-measured copy bandwidth)"
    except Exception:
        return 6650.0, "fallback of B200_PROFILING.md (MEASURED_PEAKS.json absent)"


def traffic_from_profiles(batch):
    try:
        d = json.load(open(os.path.join(ROOT, "profiles", "r02_gemm_traffic.json")))
        return float(d["dram_bytes_per_gemm_launch_per_candidate"]) * batch
    except Exception:
        return None


def load_oracle():
    import importlib.util
    spec = importlib.util.spec_from_file_location("gp_oracle", os.path.join(ROOT, "oracle", "gp_oracle.py"))
    mod = importlib.util.module_from_spec(spec)
    sys.modules["gp_oracle"] = mod
    spec.loader.exec_module(mod)
    return mod


def blas_threads():
    """Threads the BLAS pools behind numpy / scipy actually use (reported as cpu_baseline.cores)."""
    try:
        from threadpoolctl import threadpool_info
        return max([p.get("num_threads", 1) for p in threadpool_info()] or [1])
    except Exception:
        return os.cpu_count() or 1


def use_all_host_cores():
    """Size every BLAS / OpenMP pool to the box's cores (an inherited OMP_NUM_THREADS=1 -- torchrun sets
    it -- would otherwise time the reference on one thread).  Returns the thread count in effect."""
    import scipy.linalg  # noqa: F401  (loads scipy's own OpenBLAS so the limit reaches it too)
    ncpu = len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)
    try:
        from threadpoolctl import threadpool_limits
        threadpool_limits(limits=ncpu)
    except Exception:
        pass
    return blas_threads()


def cpu_eval_once(oracle, impl, x, y, theta):
    """Cold-cache GP.log_lh + GP.dloglh_dtheta the way gp/gp.py computes them (fresh object)."""
    g = oracle.OracleGP(oracle.GAUSSIAN, theta[:2], x, y, theta[2], impl)
    return float(g.log_lh), g.dloglh_dtheta


def run_reference(args):
    """The reference's own CPU implementation of the path on the box's host cores:
    oracle/_ref (its compiled Cython) + the scipy/numpy calls gp/gp.py makes."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    oracle = load_oracle()
    impl = "ref" if oracle.have_ref() else "c"
    cores = use_all_host_cores()
    x, y = synth_xy(args.n, 0)
    th = candidates(0, 0, max(args.steps + args.warmup, 1))
    for w in range(args.warmup):
        cpu_eval_once(oracle, impl, x, y, th[w])
    t0 = time.perf_counter()
    for k in range(args.steps):
        cpu_eval_once(oracle, impl, x, y, th[args.warmup + k])
    dt = time.perf_counter() - t0
    value = args.steps / dt
    sample = "1 candidate (one cold-cache log_lh + dloglh_dtheta at N=%d) per step" % args.n
    line = {
        "impl": "reference", "metric": "log_lh+dloglh_dtheta evals/sec at N=%d fp64" % args.n,
        "value": value, "unit": "evals/s", "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": dt / args.steps * 1e3, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": "C2: GaussianKernel 1-D GP, N=%d, log_lh + dloglh_dtheta per candidate" % args.n,
                   "sample": sample},
        "cpu_baseline": {"value": value, "unit": "evals/s", "cores": cores,
                         "kind": "reference" if impl == "ref" else "port", "sample": sample},
        "e2e": {"value": value, "unit": "evals/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    emit_line(line)


def run_ours(args):
    import torch
    import torch.distributed as dist
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    import gaussian_processes_b200 as gpb
    from gaussian_processes_b200 import _lib, engine
    import ctypes

    n, B, K, W = args.n, args.batch, args.steps, args.warmup
    x, y = synth_xy(n, 0)
    ev = engine.BatchEvaluator(engine.GAUSSIAN, x, y, max_batch=B)
    nth = 3
    table = torch.empty(world * B, 8, dtype=torch.float64, device="cuda") if world > 1 else None

    def step(k):
        res = ev.eval_device(candidates(k, rank, B), want_grad=True)
        if world > 1:       # the path's one exchange: gather per-candidate rows, argmax everywhere
            dist.all_gather_into_tensor(table, res)
            return torch.argmax(torch.nan_to_num(table[:, 0], nan=-float("inf")))
        return res

    def sync():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for k in range(max(W, 3)):
        step(k)
    sync()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    launches0 = _lib.lib.gpb_launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    sync()
    e0.record()
    for k in range(K):
        step(100 + k)
    e1.record()
    sync()
    ms = e0.elapsed_time(e1)
    launches = _lib.lib.gpb_launch_count() - launches0
    clocks = sampler.stop() if rank == 0 else None
    t = torch.tensor([ms], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms = float(t.item())
    value = world * B * K / (ms * 1e-3)

    # ---- end to end through the public API, host buffers every step -------------------
    gp = gpb.GP(gpb.GaussianKernel(1.0, 0.5), x, y, s=1.0)
    shifts = [x + 1e-9 * (k + 1) for k in range(K + 2)]
    for k in range(2):
        gp.x = shifts[k]
        gp.batch_eval(candidates(50 + k, rank, B))
    sync()
    t0 = time.perf_counter()
    for k in range(K):
        gp.x = shifts[2 + k]            # new host arrays: pinned staging + H2D inside the timed region
        gp.y = y + 0.0
        llh, grad = gp.batch_eval(candidates(200 + k, rank, B))      # D2H of the [B, 8] result
    sync()
    te = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(te, op=dist.ReduceOp.MAX)
    e2e = world * B * K / float(te.item())

    # ---- roofline of the dominant kernel (DMMA GEMM), event-timed per launch --------------
    roof = None
    cpu = None
    if rank == 0:
        tf, pms = ctypes.c_double(), ctypes.c_double()
        _lib.call("gpb_microbench_fp64", 1, 20000, ctypes.byref(tf), ctypes.byref(pms))
        peak = tf.value
        _lib.lib.gpb_profile_enable.argtypes = [ctypes.c_int]
        _lib.lib.gpb_profile_enable.restype = None
        _lib.lib.gpb_profile_read.argtypes = [ctypes.c_int, _lib.dp, ctypes.POINTER(ctypes.c_int64)]
        # per-launch event timing needs the launches serialised on one stream
        _lib.set_option("eval_streams", 1)
        ev1 = engine.BatchEvaluator(engine.GAUSSIAN, x, y, max_batch=B)
        ev1.eval_device(candidates(299, rank, B), want_grad=True)
        torch.cuda.synchronize()
        _lib.lib.gpb_profile_enable(1)
        psteps = 2
        for k in range(psteps):
            ev1.eval_device(candidates(300 + k, rank, B), want_grad=True)
        torch.cuda.synchronize()
        classes = {}
        for cid, nm in enumerate(["gemm_nt(DMMA)", "potrf_diag", "build+matvec", "trsv", "reduce", "misc"]):
            m, c = ctypes.c_double(), ctypes.c_int64()
            _lib.lib.gpb_profile_read(cid, ctypes.byref(m), ctypes.byref(c))
            classes[nm] = {"ms_per_step": m.value / psteps, "launches_per_step": c.value / psteps}
        _lib.lib.gpb_profile_enable(0)
        _lib.set_option("eval_streams", 0)
        del ev1
        gemm_ms = classes["gemm_nt(DMMA)"]["ms_per_step"]
        gemm_launches = max(classes["gemm_nt(DMMA)"]["launches_per_step"], 1)
        flops_step = B * float(n) ** 3            # N^3/3 potrf + 2N^3/3 inverse per eval
        achieved = flops_step / (gemm_ms * 1e-3) / 1e12
        roof = {"bound": "tensor", "kernel": "gemm_nt_tma_kernel (FP64 DMMA.8x8x4, TMA + mbarrier pipeline)",
                "achieved": achieved, "peak": peak, "unit": "TFLOP/s", "frac": achieved / peak,
                # dram__bytes_read.sum + dram__bytes_write.sum of the gemm_nt launches of one step, read from the
                # summary of the committed ncu pass of the CURRENT code (profiles/ncu_traffic.py writes it);
                # null when no such summary exists -- never a constant carried over from older code
                "traffic": traffic_from_profiles(B),
                "traffic_source": "profiles/r02_gemm_traffic.json: ncu dram__bytes_read.sum + dram__bytes_write.sum per "
                                  "gemm_nt launch per candidate of one evaluator step (scaled by this run's batch)",
                "flops_note": "all N^3 algorithmic flop of an evaluation are booked to the GEMM class; the diagonal-block "
                              "kernel performs (N/128) * 128^3 = N * 128^2 of them (0.1 %% at N=%d)" % n,
                "hbm_peak_gbs": hbm_peak()[0], "hbm_peak_source": hbm_peak()[1],
                "algorithmic_flops_per_launch": flops_step / gemm_launches,
                "avg_launch_ms": gemm_ms / gemm_launches, "launches_per_step": gemm_launches,
                "peak_source": "measured live: gpb_microbench_fp64 (DMMA.8x8x4 issue rate, 148x8 CTAs); "
                               "MEASURED_PEAKS.json has no fp64 figure",
                "whole_step_frac": flops_step / (ms / K * 1e-3) / 1e12 / peak,
                "note": "kernel durations event-timed on one stream (eval_streams=1); the timed region "
                        "itself runs 4 candidate groups on concurrent streams",
                "kernel_classes": classes}
        # ---- the factorisation alone (north star: Cholesky >= 60 % of the FP64 tensor peak): one
        # gpb_potrf over the B matrices of a step, N^3/3 flop each, event-timed on the launch stream
        from gaussian_processes_b200 import device as D
        assert n % 128 == 0, "the bench workload uses N multiple of 128"
        eng0 = engine.Engine(engine.GAUSSIAN, (1.0, 0.5), 1.0, x, y)
        Lb = D.empty(B, n, n)
        Wb, Vb, infob = D.empty(B, n, n), D.empty(B, n, n), D.izeros(B)
        t_potrf = 1e30
        for rep in range(3):
            for b in range(B):
                eng0.build(eng0.dx, n, eng0.dx, n, n, n, 1, add_diag=True, pad_identity=True, out=Lb[b:b + 1])
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            torch.cuda.synchronize()
            e0.record()
            _lib.call("gpb_potrf", D.ptr(Lb), n, n, n * n, B, D.ptr(Wb), n, n * n, D.ptr(Vb), n, n * n,
                      D.ptr(infob), D.stream_ptr())
            e1.record()
            torch.cuda.synchronize()
            t_potrf = min(t_potrf, e0.elapsed_time(e1))
        potrf_tf = B * float(n) ** 3 / 3 / (t_potrf * 1e-3) / 1e12
        roof["cholesky_batched"] = {"tflops": potrf_tf, "frac": potrf_tf / peak, "ms": t_potrf, "batch": B,
                                    "info_max": int(infob.max().item()),
                                    "note": "gpb_potrf alone, %d matrices of N=%d in one call on one stream "
                                            "(algorithmic N^3/3 flop each)" % (B, n)}
        del Lb, Wb, Vb, eng0
        # ---- CPU baseline: the reference's path on this box's host cores (bounded sample) ------
        if world == 1 and not args.no_cpu_baseline:
            oracle = load_oracle()
            impl = "ref" if oracle.have_ref() else "c"
            use_all_host_cores()
            th = candidates(0, 0, 3)
            cpu_eval_once(oracle, impl, x, y, th[0])
            t0 = time.perf_counter()
            ns = 2
            for k in range(ns):
                cl, cg = cpu_eval_once(oracle, impl, x, y, th[1 + k])
            cdt = time.perf_counter() - t0
            cpu = {"value": ns / cdt, "unit": "evals/s", "cores": blas_threads(),
                   "kind": "reference" if impl == "ref" else "port",
                   "sample": "%d of the %d candidates of one step (N=%d), after 1 warm-up eval" % (ns, B, n)}
            # the sample doubles as a parity spot-check of the timed path
            chk = ev.eval(th[1 + ns - 1:1 + ns], want_grad=True)
            cpu["parity_rel_err_log_lh"] = abs(chk[0][0] - cl) / abs(cl)
            cpu["parity_rel_err_grad"] = float(np.max(np.abs(chk[1][0] - cg)) / np.max(np.abs(cg)))

    # ---- second half of BASELINE's metric: posterior mean / cov test points per second ----------
    posterior = None
    if rank == 0 and not args.no_posterior:
        gpp = gpb.GP(gpb.GaussianKernel(1.0, 0.5), x, y, s=1.0)
        m_mean, m_cov = 16384, 4096
        xo_mean = np.linspace(-2 * np.pi, 2 * np.pi, m_mean)
        xo_cov = np.linspace(-2 * np.pi, 2 * np.pi, m_cov)
        gpp.mean(xo_mean[:256]); gpp.cov(xo_cov)                # fit once: factor, solves, L^-1 (cached); page-lock the result pool
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for k in range(5):
            gpp.mean(xo_mean + 1e-9 * k)
        t_mean = (time.perf_counter() - t0) / 5
        t0 = time.perf_counter()
        for k in range(3):
            gpp.cov(xo_cov + 1e-9 * k)
        t_cov = (time.perf_counter() - t0) / 3
        posterior = {"mean_test_pts_per_s": m_mean / t_mean, "cov_test_pts_per_s": m_cov / t_cov,
                     "m_mean": m_mean, "m_cov": m_cov, "n": n,
                     "note": "public API on a fitted GP (factor + L^-1 cached), host xo in, numpy out; the cov figure "
                             "includes the D2H of the M x M result (134 MB at M=4096) and is PCIe-bound -- the "
                             "device-only rate is in posterior_sharded.cov_test_pts_per_s_device"}

    # ---- the drop-in property API on ONE GP object at N=4096 (config C2 as a user of gp/gp.py drives it):
    # cold log_lh, + dloglh_dtheta, + d2lh_dtheta2 after a parameter change (setters drop every cached result)
    single = None
    if rank == 0 and not args.no_posterior:
        gps1 = gpb.GP(gpb.GaussianKernel(1.0, 0.5), x, y, s=1.0)
        for k in range(3):
            gps1.set_param("w", 0.5 + 1e-6 * (k + 1)); gps1.log_lh; gps1.dloglh_dtheta; gps1.d2lh_dtheta2
        torch.cuda.synchronize()
        ts = np.zeros((8, 3))
        for k in range(ts.shape[0]):
            gps1.set_param("w", 0.5 + 1e-5 * (k + 1))
            t0 = time.perf_counter()
            l1 = gps1.log_lh
            ts[k, 0] = time.perf_counter() - t0
            g1 = gps1.dloglh_dtheta
            ts[k, 1] = time.perf_counter() - t0
            h1 = gps1.d2lh_dtheta2
            ts[k, 2] = time.perf_counter() - t0
        # the same, gradient asked first (one staged chain: factor + inverse + gradient in one library call)
        tg = []
        for k in range(8):
            gps1.set_param("w", 0.5 - 1e-5 * (k + 1))
            t0 = time.perf_counter()
            g1 = gps1.dloglh_dtheta; l1 = gps1.log_lh
            tg.append(time.perf_counter() - t0)
        tb = ts.min(axis=0)
        single = {"workload": "C2 through the drop-in property API: one GP object, N=%d, cold after set_param" % n,
                  "log_lh_ms": tb[0] * 1e3, "log_lh_plus_dloglh_ms": min(min(tg), tb[1]) * 1e3,
                  "plus_d2lh_dtheta2_ms": tb[2] * 1e3,
                  "evals_per_s_log_lh_plus_dloglh": 1.0 / min(min(tg), tb[1]),
                  "tflops_log_lh_plus_dloglh": float(n) ** 3 / min(min(tg), tb[1]) / 1e12,
                  "frac_of_dmma_peak": (float(n) ** 3 / min(min(tg), tb[1]) / 1e12 / roof["peak"]) if roof else None,
                  "note": "host wall time of the property reads (each ends with the read-back of its scalars); "
                          "the factorisation is the persistent dataflow launch of csrc/chain.cu"}

    # ---- BASELINE config C4: 4096 restarts x N=1024, candidates sharded over the ranks, argmax gather ----
    c4 = None
    if not args.no_posterior:
        sys.path.insert(0, os.path.join(ROOT, "tests", "golden"))
        from make_golden_full import c4_candidates
        x4, y4 = synth_xy(1024, 0)
        cand4 = c4_candidates()
        gp4 = gpb.GP(gpb.GaussianKernel(1.0, 0.5), x4, y4, s=1.0)
        gp4.fit_MLII(cand4[:256 * world], set_params=False)
        t4 = []
        for k in range(3):
            sync()
            t0 = time.perf_counter()
            res4 = gp4.fit_MLII(cand4 * (1.0 + 1e-9 * k), set_params=False)
            torch.cuda.synchronize()
            tt = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device="cuda")
            if world > 1:
                dist.all_reduce(tt, op=dist.ReduceOp.MAX)
            t4.append(float(tt.item()))
        c4 = {"workload": "C4: fit_MLII over 4096 candidates x N=1024 (h~U(.5,2), w~U(pi/32,pi/2), s~U(.75,1.5)), "
                          "candidates sharded over %d GPU(s), all-gather of the [B,8] rows + argmax" % world,
              "scaling": "strong", "ms": min(t4) * 1e3, "evals_per_s": cand4.shape[0] / min(t4),
              "tflops": cand4.shape[0] * 1024.0 ** 3 / min(t4) / 1e12, "best_index": int(res4.best_index),
              "best_log_lh": float(res4.best_log_lh),
              "note": "public API end to end: host candidates in, argmax + table out (reference best index 4081, "
                      "log_lh -681.7116082305: tests/golden/full_c4.npz)"}

    # ---- BASELINE config C3 on N GPUs: test points sharded over the ranks (strong scaling: M fixed) ----
    sharded = None
    if not args.no_sharded:
        sharded = posterior_sharded(world, rank, sync)

    # ---- BASELINE configs[0], the reference's own CPU-runnable case: N=50, one GP object -------
    # cold log_lh + dloglh_dtheta + mean + cov at 100 test points through the public API; the
    # reference's path (oracle/_ref + scipy/numpy) timed beside it on the host cores
    small = None
    if rank == 0 and not args.no_posterior:
        xs = np.linspace(-2 * np.pi, 2 * np.pi, 50)
        ys, xos = np.sin(xs), np.linspace(-2 * np.pi, 2 * np.pi, 100)
        gps = gpb.GP(gpb.GaussianKernel(1.0, 0.2), xs, ys, s=0)

        def bundle(k):
            gps.set_param("w", 0.2 + 1e-6 * (k + 1))            # setters drop every cached result (gp.py:231-240)
            return gps.log_lh, gps.dloglh_dtheta, gps.mean(xos), gps.cov(xos)
        for k in range(20):
            bundle(k)
        torch.cuda.synchronize()
        reps = 200
        t0 = time.perf_counter()
        for k in range(reps):
            bundle(100 + k)
        t_b = (time.perf_counter() - t0) / reps
        small = {"workload": "C1: N=50 GaussianKernel(1, 0.2), s=0: cold log_lh + dloglh_dtheta + mean + cov at 100 test points, one GP object",
                 "ms_per_bundle": t_b * 1e3, "bundles_per_s": 1.0 / t_b,
                 "launches_per_bundle": "3 (build, factor+invert, tail) + 1 (mean) + 1 (cov)", "host_syncs_per_bundle": 3}
        if not args.no_cpu_baseline:
            oracle = load_oracle()
            impl = "ref" if oracle.have_ref() else "c"
            t0 = time.perf_counter()
            for k in range(reps):
                o = oracle.OracleGP(oracle.GAUSSIAN, (1.0, 0.2 + 1e-6 * (k + 1)), xs, ys, 0.0, impl)
                o.log_lh, o.dloglh_dtheta, o.mean(xos), o.cov(xos)
            small["cpu_reference_ms_per_bundle"] = (time.perf_counter() - t0) / reps * 1e3

    if rank == 0:
        line = {
            "metric": "log_lh+dloglh_dtheta evals/sec at N=%d fp64" % n,
            "value": value, "unit": "evals/s", "n_gpus": world, "steps": K, "warmup": max(W, 3),
            "ms_per_step": ms / K, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f64", "data": "synthetic",
            "config": {"workload": "C2: GaussianKernel 1-D GP, N=%d, Kxx build + Cholesky + log_lh + dloglh_dtheta "
                                   "for a batch of %d hyperparameter candidates per GPU per step" % (n, B),
                       "batch_per_gpu": B, "n": n, "theta": "h,w,s around (1, 0.5, 1); new candidates every step",
                       "l2": "working set %.1f GB per step >> 126 MB L2 (inputs larger than L2, no flush needed)"
                             % (4 * B * n * n * 8 / 1e9),
                       "parallelism": "candidates sharded over %d GPU(s), NCCL all-gather of [B,8] rows + argmax" % world},
            "e2e": {"value": e2e, "unit": "evals/s", "h2d_bytes_per_step": 2 * n * 8 + B * nth * 8,
                    "d2h_bytes_per_step": B * 8 * 8},
            "gpu_launches": int(launches),
            "clocks": clocks,
            "roofline": roof,
        }
        if cpu is not None:
            line["cpu_baseline"] = cpu
        if posterior is not None:
            line["posterior"] = posterior
        if sharded is not None:
            line["posterior_sharded"] = sharded
        if single is not None:
            line["single_gp"] = single
        if c4 is not None:
            line["c4"] = c4
        if small is not None:
            line["small_gp"] = small
        emit_line(line)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def posterior_sharded(world, rank, sync):
    """BASELINE config C3: PeriodicKernel(1,1,1), s=1, N=8192, posterior at M=16384 test points partitioned
    over the ranks (gaussian_processes_b200.sharded_posterior, cov_layout="lower"): each rank computes its
    rows of Z = K(xo,x) L^-T, ONE all-gather of Z over NVLink, then its block rows of the lower triangle of
    cov.  M is fixed: strong scaling.  Every rank holds a fitted GP (factor + L^-1: `fit_ms`, each rank
    redundantly -- fit once, predict many)."""
    import torch
    import torch.distributed as dist
    import gaussian_processes_b200 as gpb
    n, m = 8192, 16384
    x, y = synth_xy(n, 0)
    xo = np.linspace(-2 * np.pi, 2 * np.pi, m)

    def maxtime(t):
        tt = torch.tensor([t], dtype=torch.float64, device="cuda")
        if world > 1:
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        return float(tt.item())
    gp = gpb.GP(gpb.PeriodicKernel(1.0, 1.0, 1.0), x, y, s=1.0)
    sync()
    t0 = time.perf_counter()
    llh = gp.log_lh
    gp._engine().inv_factor()
    torch.cuda.synchronize()
    fit_ms = maxtime(time.perf_counter() - t0) * 1e3
    kw = dict(cov_layout="lower")
    gpb.sharded_posterior(gp, xo, **kw)                              # warm: allocator, page-locked pool, NCCL
    t_mean, t_dev, t_e2e = [], [], []
    for rep in range(3):
        xr = xo + 1e-9 * rep
        sync()
        t0 = time.perf_counter()
        mean, _, plan = gpb.sharded_posterior(gp, xr, want_cov=False, **kw)
        torch.cuda.synchronize()
        t_mean.append(maxtime(time.perf_counter() - t0))
        sync()
        t0 = time.perf_counter()
        _, pd, _ = gpb.sharded_posterior(gp, xr, host=False, **kw)
        torch.cuda.synchronize()
        t_dev.append(maxtime(time.perf_counter() - t0))
        del pd
        sync()
        t0 = time.perf_counter()
        _, ph, _ = gpb.sharded_posterior(gp, xr, **kw)
        torch.cuda.synchronize()
        t_e2e.append(maxtime(time.perf_counter() - t0))
        d2h = sum(p[2].nbytes for p in ph)
        del ph
    tm = {}
    sync()
    gpb.sharded_posterior(gp, xo, host=False, timings=tm, **kw)
    parts = {k: maxtime(v) * 1e3 for k, v in tm.items() if k != "allgather_bytes"}
    npad = 8192
    flops = (float(n) ** 2 * m + float(n) * m * m / 2.0 * (1.0 + 1.0 / plan.nb)) * 2.0 / 2.0 * 2.0 / 2.0
    # N^2 M (Z = K W^T, W triangular: N^2 M multiply-adds = 2 * N^2 M / 2 flop) + N M^2 (1 + 1/nb) / 2 * 2 flop
    flops = float(n) ** 2 * m + float(n) * m * m * (1.0 + 1.0 / plan.nb)
    return {"workload": "C3: PeriodicKernel(1,1,1), s=1, N=%d; posterior at M=%d test points sharded over %d GPU(s)" % (n, m, world),
            "scaling": "strong", "n": n, "m": m, "n_gpus": world, "blocks": plan.nb, "block_points": plan.bs,
            "fit_ms_each_rank": fit_ms, "log_lh": float(llh),
            "mean_test_pts_per_s_e2e": m / min(t_mean), "mean_ms_e2e": min(t_mean) * 1e3,
            "cov_test_pts_per_s_device": m / min(t_dev), "cov_ms_device": min(t_dev) * 1e3,
            "cov_test_pts_per_s_e2e": m / min(t_e2e), "cov_ms_e2e": min(t_e2e) * 1e3,
            "cov_d2h_bytes_per_rank": int(d2h),
            "cov_e2e_note": "e2e = device time + the D2H of this rank's lower panels into page-locked numpy arrays "
                            "(PCIe-bound: %.2f GB per rank)" % (d2h / 1e9),
            "cov_tflops_device": flops / min(t_dev) / 1e12,
            "collective": {"op": "all_gather_into_tensor of Z (NCCL over NVLink)" if world > 1 else "none (one rank)",
                           "bytes_total": int(tm.get("allgather_bytes", 0)), "ms": parts.get("allgather")},
            "phase_ms": parts,
            "note": "times are max over ranks; mean/cov e2e take host xo in and return host arrays; the result "
                    "of cov stays sharded (block rows of the lower triangle, (1 + 1/%d)/2 of M^2 doubles in total)" % plan.nb}


_REAL_STDOUT = None


def _guard_stdout():
    """stdout carries exactly ONE JSON line.  Native libraries write there too (NCCL prints its version
    banner on fd 1 at communicator creation whenever NCCL_DEBUG >= VERSION), so fd 1 is pointed at
    stderr for the whole run and the result line goes to a private duplicate of the original stdout."""
    global _REAL_STDOUT
    if _REAL_STDOUT is None:
        sys.stdout.flush()
        _REAL_STDOUT = os.dup(1)
        os.dup2(2, 1)


def emit_line(line):
    data = (json.dumps(line) + "\n").encode()
    if _REAL_STDOUT is None:
        sys.stdout.write(data.decode())
        sys.stdout.flush()
    else:
        os.write(_REAL_STDOUT, data)


def main():
    _guard_stdout()
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--batch", type=int, default=32, help="candidates per GPU per step (17 GB workspace at N=4096)")
    ap.add_argument("--n", type=int, default=N_DEFAULT)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-posterior", action="store_true")
    ap.add_argument("--no-sharded", action="store_true", help="skip the C3 posterior with test points sharded over the ranks")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
