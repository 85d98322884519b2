"""sharded_posterior over NCCL: run with torchrun (--nproc-per-node G).  Every rank checks the gathered
mean / covariance against its own single-GPU computation."""
import json, os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch
import torch.distributed as dist
import gaussian_processes_b200 as gpb
from conftest import synth_xy
rank, world = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", 0)))
if world > 1:
    dist.init_process_group("nccl")
n, m = 2048, 4099
x, y = synth_xy(n, 0)
xo = np.linspace(-2 * np.pi, 2 * np.pi, m)
gp = gpb.GP(gpb.PeriodicKernel(1.0, 1.0, 1.0), x, y, s=1.0)
ref_mean, ref_cov = gp.mean(xo), gp.cov(xo)
for rep in range(2):
    torch.cuda.synchronize()
    if world > 1: dist.barrier()
    t0 = time.perf_counter()
    mean, rows, (lo, hi) = gpb.sharded_posterior(gp, xo)
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
_, full, _ = gpb.sharded_posterior(gp, xo, gather_cov=True)
scale = np.abs(ref_cov).max()
out = dict(rank=rank, world=world, lo=lo, hi=hi, ms=dt * 1e3,
           mean_err=float(np.abs(mean - ref_mean).max() / np.abs(ref_mean).max()),
           rows_err=float(np.abs(rows - ref_cov[lo:hi]).max() / scale),
           full_err=float(np.abs(full - ref_cov).max() / scale))
print(json.dumps(out), flush=True)
assert out["mean_err"] < 1e-9 and out["rows_err"] < 1e-9 and out["full_err"] < 1e-9      # the 1e-9 parity bar (cov_rows takes the K^-1 form)
if world > 1:
    dist.barrier(); dist.destroy_process_group()
