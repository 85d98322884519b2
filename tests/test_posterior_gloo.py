"""Test points sharded over ranks: world_size 2 over gloo on CPU.  The per-block
computations are injected (CPU oracle); what runs is the product's own partition and ragged
all-gather -- the logic that runs over NCCL on the GPUs."""
import os
import sys

import numpy as np
import pytest
import torch.distributed as dist
import torch.multiprocessing as mp

from conftest import ROOT, load_oracle, synth_xy


def _worker(rank, world, port, m, outdir):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import gaussian_processes_b200 as gpb
    from conftest import load_oracle, synth_xy
    oracle = load_oracle()
    x, y = synth_xy(40, 1)
    o = oracle.OracleGP(oracle.PERIODIC, (1.0, 0.9, 1.5), x, y, 0.7)
    xo = np.linspace(-6, 6, m)
    seen = []

    def mean_fn(xb):
        seen.append(len(xb))
        return o.mean(xb) if len(xb) else np.empty(0)

    def cov_rows_fn(xall, lo, hi):
        return o.cov(xall)[lo:hi]
    gp = gpb.GP(gpb.PeriodicKernel(1.0, 0.9, 1.5), x, y, s=0.7)
    mean, rows, (lo, hi) = gpb.sharded_posterior(gp, xo, mean_fn=mean_fn, cov_rows_fn=cov_rows_fn)
    mean2, full, _ = gpb.sharded_posterior(gp, xo, gather_cov=True, mean_fn=mean_fn, cov_rows_fn=cov_rows_fn)
    mean3, none, _ = gpb.sharded_posterior(gp, xo, want_cov=False, mean_fn=mean_fn, cov_rows_fn=cov_rows_fn)
    assert none is None and np.array_equal(mean, mean2) and np.array_equal(mean, mean3)
    np.savez(os.path.join(outdir, "r%d.npz" % rank), mean=mean, rows=rows, full=full, lo=lo, hi=hi, seen=np.array(seen))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("m", [1, 9, 64, 300])
def test_sharded_posterior_world2_gloo(tmp_path, m):
    port = 31500 + (os.getpid() % 2000) + m
    mp.spawn(_worker, args=(2, port, m, str(tmp_path)), nprocs=2, join=True)
    oracle = load_oracle()
    x, y = synth_xy(40, 1)
    o = oracle.OracleGP(oracle.PERIODIC, (1.0, 0.9, 1.5), x, y, 0.7)
    xo = np.linspace(-6, 6, m)
    ref_mean, ref_cov = o.mean(xo), o.cov(xo)
    r0, r1 = dict(np.load(tmp_path / "r0.npz")), dict(np.load(tmp_path / "r1.npz"))
    assert int(r0["lo"]) == 0 and int(r0["hi"]) == int(r1["lo"]) and int(r1["hi"]) == m      # a partition
    if m < 256:
        assert abs((int(r0["hi"]) - int(r0["lo"])) - (int(r1["hi"]) - int(r1["lo"]))) <= 1
    else:                                        # shards start on tile boundaries
        assert int(r1["lo"]) % 128 == 0
    assert np.array_equal(r0["mean"], r1["mean"])
    for r in (r0, r1):
        assert np.allclose(r["mean"], ref_mean, rtol=1e-13, atol=1e-14)   # every rank holds the full mean (BLAS blocks differ in the last bit)
        assert np.array_equal(r["full"], ref_cov)                     # ... and the full matrix when asked
        assert np.array_equal(r["rows"], ref_cov[int(r["lo"]):int(r["hi"])])
        assert int(r["seen"][0]) == int(r["hi"]) - int(r["lo"])       # each rank evaluated only its block
