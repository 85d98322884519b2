"""torchrun check of sharded_posterior(cov_layout="lower") over NCCL: every rank's panels and the gathered
matrix against the single-GPU GP.cov / GP.mean of the same problem."""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch
import torch.distributed as dist
local = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
import gaussian_processes_b200 as gpb
from conftest import synth_xy
rank, world = dist.get_rank(), dist.get_world_size()
for n, m, k in ((1000, 3000, gpb.PeriodicKernel(1.0, 1.0, 1.0)), (2048, 5000, gpb.GaussianKernel(1.0, 0.5)), (300, 40, gpb.GaussianKernel(1.0, 0.5))):
    x, y = synth_xy(n, 0)
    xo = np.linspace(-2 * np.pi, 2 * np.pi, m)
    gp = gpb.GP(k, x, y, s=1.0)
    ref_c, ref_m = gp.cov(xo), gp.mean(xo)
    scale = np.abs(ref_c).max()
    tm = {}
    mean, panels, plan = gpb.sharded_posterior(gp, xo, cov_layout="lower", timings=tm)
    err = max([np.abs(P - ref_c[lo:hi, :hi]).max() for lo, hi, P in panels] + [0.0]) / scale
    em = np.abs(mean - ref_m).max() / np.abs(ref_m).max()
    _, full, _ = gpb.sharded_posterior(gp, xo, cov_layout="lower", gather_cov=True)
    ef = np.abs(full - ref_c).max() / scale
    rows = sum(hi - lo for lo, hi, _ in panels)
    print("rank %d/%d n=%d m=%d: blocks %s rows %d panel err %.2e mean err %.2e gathered err %.2e allgather %.2f ms (%d bytes)" % (
        rank, world, n, m, plan.mine(rank), rows, err, em, ef, tm["allgather"] * 1e3, tm["allgather_bytes"]), flush=True)
    assert err <= 1e-11 and em <= 1e-12 and ef <= 1e-11
dist.barrier()
if rank == 0:
    print("sharded lower-panel posterior over NCCL: OK")
dist.destroy_process_group()
