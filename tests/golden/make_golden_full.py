"""Full-size digests of BASELINE configs C3, C4 and C5 from the UNMODIFIED reference
(build container only; minutes of CPU).  Same loader as make_golden.py: the reference's
own gp/gp.py + gp/kernels/*.py on top of its own Cython compiled in place (oracle/_ref).

  python tests/golden/make_golden_full.py [c3] [c4] [c5]

Fixtures (inputs are regenerated from the seeds by the tests):
  full_c3.npz  PeriodicKernel(1,1,1), s=1, N=8192 (x law of SURVEY 8d, seed 0), M=16384 test points:
               log_lh, dloglh_dtheta, inv_Kxx_y, mean at all 16384 points, rows {0, 8191, 16383} of cov,
               diag(cov), and per-slice (sum, sum of squares, 64 sampled entries) of Kxx, Kxx_J, Kxx_H.
  full_c4.npz  GaussianKernel, N=1024, the 4096 BASELINE candidates (h~U(.5,2), w~U(pi/32,pi/2),
               s~U(.75,1.5), RandomState(4)): reference log_lh of ALL candidates, dloglh_dtheta of a
               fixed 64-row subset, plus 16 clamp candidates s~U(0,0.5) (log_lh = -inf rule).
  full_c5.npz  GaussianKernel(1,0.5), s=1 at N=8192 and N=16384 (the reference cannot hold N=32768 in
               host RAM): log_lh, dloglh_dtheta, inv_Kxx_y digest -- the "extrapolated" pin of C5.
"""
import os
import sys
import time

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
from make_golden import load_reference, synth_xy  # noqa: E402

SAMPLE_SEED = 7


def sample_idx(n1, n2, k=64):
    rng = np.random.RandomState(SAMPLE_SEED)
    return rng.randint(0, n1, k), rng.randint(0, n2, k)


def digest(a, ri, ci):
    """(sum, sum of squares, sampled entries) of one 2-D slice."""
    return np.float64(a.sum()), np.float64((a * a).sum()), a[ri, ci].copy()


def c4_candidates(B=4096, seed=4):
    rng = np.random.RandomState(seed)
    h = rng.uniform(0.5, 2.0, B)
    w = rng.uniform(np.pi / 32., np.pi / 2., B)
    s = rng.uniform(0.75, 1.5, B)
    return np.stack([h, w, s], axis=1)


def c4_clamp_candidates(B=16, seed=5):
    rng = np.random.RandomState(seed)
    h = rng.uniform(0.5, 2.0, B)
    w = rng.uniform(np.pi / 32., np.pi / 2., B)
    s = rng.uniform(0.0, 0.5, B)           # tests/util.py:24-25 -- logdet < MIN at N=1024
    return np.stack([h, w, s], axis=1)


def make_c3(gp):
    n, m = 8192, 16384
    x, y = synth_xy(n, 0)
    xo = np.linspace(-2 * np.pi, 2 * np.pi, m)
    t0 = time.time()
    g = gp.GP(gp.PeriodicKernel(1.0, 1.0, 1.0), x, y, s=1.0)
    out = dict(n=n, m=m, seed=0, params=g.params)
    out["log_lh"] = np.float64(g.log_lh)
    print("c3 log_lh %.12f (%.0f s)" % (out["log_lh"], time.time() - t0), flush=True)
    out["inv_Kxx_y"] = g.inv_Kxx_y
    out["dloglh_dtheta"] = g.dloglh_dtheta
    print("c3 dloglh", out["dloglh_dtheta"], "(%.0f s)" % (time.time() - t0), flush=True)
    ri, ci = sample_idx(n, n)
    out["sample_rows"], out["sample_cols"] = ri, ci
    K = g.Kxx
    out["Kxx_sum"], out["Kxx_sumsq"], out["Kxx_samples"] = digest(K, ri, ci)
    J = g.Kxx_J
    for i in range(3):
        out["J%d_sum" % i], out["J%d_sumsq" % i], out["J%d_samples" % i] = digest(J[i], ri, ci)
    del g._memoized["Kxx_J"]
    del J
    H = g.Kxx_H
    print("c3 hessian built (%.0f s)" % (time.time() - t0), flush=True)
    for i in range(3):
        for j in range(3):
            k = "H%d%d" % (i, j)
            out[k + "_sum"], out[k + "_sumsq"], out[k + "_samples"] = digest(H[i, j], ri, ci)
    del g._memoized["Kxx_H"]
    del H
    out["mean"] = g.mean(xo)
    print("c3 mean (%.0f s)" % (time.time() - t0), flush=True)
    c = g.cov(xo)
    print("c3 cov (%.0f s)" % (time.time() - t0), flush=True)
    out["cov_rows"] = np.array([0, 8191, 16383])
    out["cov_row_values"] = c[[0, 8191, 16383]].copy()
    out["cov_diag"] = np.diag(c).copy()
    out["cov_fro"] = np.float64(np.linalg.norm(c))
    out["cov_max"] = np.float64(np.abs(c).max())
    np.savez_compressed(os.path.join(HERE, "full_c3.npz"), **out)
    print("full_c3.npz written (%.0f s)" % (time.time() - t0), flush=True)


def make_c4(gp):
    n = 1024
    x, y = synth_xy(n, 0)
    cand = c4_candidates()
    clamp = c4_clamp_candidates()
    t0 = time.time()
    g = gp.GP(gp.GaussianKernel(1.0, 0.5), x, y, s=1.0)
    llh = np.empty(cand.shape[0])
    for b, th in enumerate(cand):
        g.params = th
        llh[b] = g.log_lh
        if b % 512 == 0:
            print("c4 %d/%d (%.0f s)" % (b, cand.shape[0], time.time() - t0), flush=True)
    rng = np.random.RandomState(SAMPLE_SEED)
    subset = np.sort(rng.choice(cand.shape[0], 64, replace=False))
    best = int(np.argmax(llh))
    if best not in subset:
        subset[0] = best
        subset = np.sort(subset)
    grad = np.empty((subset.size, 3))
    for q, b in enumerate(subset):
        g.params = cand[b]
        grad[q] = g.dloglh_dtheta
    cl_llh = np.empty(clamp.shape[0])
    cl_grad = np.empty((clamp.shape[0], 3))
    for b, th in enumerate(clamp):
        g.params = th
        cl_llh[b] = g.log_lh
        cl_grad[b] = g.dloglh_dtheta
    np.savez_compressed(os.path.join(HERE, "full_c4.npz"), n=n, seed=0, cand_seed=4, clamp_seed=5,
                        log_lh=llh, best_index=best, subset=subset, subset_dloglh=grad,
                        clamp_log_lh=cl_llh, clamp_dloglh=cl_grad)
    print("full_c4.npz written: best %d log_lh %.10f, clamp -inf count %d (%.0f s)"
          % (best, llh[best], int(np.isneginf(cl_llh).sum()), time.time() - t0), flush=True)


def make_c5(gp):
    out = {}
    for n in (8192, 16384):
        t0 = time.time()
        x, y = synth_xy(n, 0)
        g = gp.GP(gp.GaussianKernel(1.0, 0.5), x, y, s=1.0)
        out["log_lh_%d" % n] = np.float64(g.log_lh)
        a = g.inv_Kxx_y
        out["inv_Kxx_y_%d" % n] = a
        print("c5 N=%d log_lh %.12f (%.0f s)" % (n, out["log_lh_%d" % n], time.time() - t0), flush=True)
        out["dloglh_%d" % n] = g.dloglh_dtheta
        out["inv_Kxx_diag_%d" % n] = np.diag(g.inv_Kxx).copy()
        print("c5 N=%d dloglh %s (%.0f s)" % (n, out["dloglh_%d" % n], time.time() - t0), flush=True)
        del g
    np.savez_compressed(os.path.join(HERE, "full_c5.npz"), **out)
    print("full_c5.npz written", flush=True)


def main():
    which = [a for a in sys.argv[1:]] or ["c4", "c3", "c5"]
    gp = load_reference()
    for w in which:
        {"c3": make_c3, "c4": make_c4, "c5": make_c5}[w](gp)


if __name__ == "__main__":
    main()
