"""Multi-rank path of fit_MLII on CPU: world_size 2 over gloo.  The candidate
evaluation is injected (CPU oracle) so that what is exercised is the product's own
sharding, ragged all-gather and argmax -- the logic that runs over NCCL on the GPUs."""
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from conftest import ROOT, load_oracle, synth_xy


def _worker(rank, world, port, B, outdir):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import gaussian_processes_b200 as gpb
    from conftest import load_oracle, synth_xy
    oracle = load_oracle()
    x, y = synth_xy(48, 3)
    rng = np.random.RandomState(5)
    cand = np.stack([rng.uniform(0.5, 2, B), rng.uniform(np.pi / 32, np.pi / 2, B), rng.uniform(0.75, 1.5, B)], axis=1)
    calls = []

    def evaluate(th):
        calls.append(len(th))
        rows = []
        for t in th:
            o = oracle.OracleGP(oracle.GAUSSIAN, t[:-1], x, y, t[-1])
            rows.append([float(o.log_lh)] + list(o.dloglh_dtheta) + [0.0])
        return torch.tensor(rows, dtype=torch.float64)
    gp = gpb.GP(gpb.GaussianKernel(1.0, 1.0), x, y, s=1.0)
    res = gpb.fit_MLII(gp, cand, evaluate=evaluate)
    np.savez(os.path.join(outdir, "r%d.npz" % rank), best=res.best_index, llh=res.log_lh,
             grad=res.dloglh_dtheta, params=gp.params, ncalls=np.array(calls), cand=cand)
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("B", [7, 8])
def test_fit_mlii_world2_gloo(tmp_path, B):
    port = 29500 + (os.getpid() % 2000) + B
    mp.spawn(_worker, args=(2, port, B, str(tmp_path)), nprocs=2, join=True)
    oracle = load_oracle()
    x, y = synth_xy(48, 3)
    r0 = dict(np.load(tmp_path / "r0.npz"))
    r1 = dict(np.load(tmp_path / "r1.npz"))
    best, llh, grad = oracle.oracle_fit_mlii(oracle.GAUSSIAN, x, y, r0["cand"])
    for r in (r0, r1):                       # every rank holds the same full table and winner
        assert int(r["best"]) == best
        assert np.array_equal(r["llh"], llh) and np.array_equal(r["grad"], grad)
        assert np.array_equal(r["params"], r0["cand"][best])
    # each rank evaluated only its shard
    assert int(r0["ncalls"].sum()) + int(r1["ncalls"].sum()) == B
    assert abs(int(r0["ncalls"].sum()) - int(r1["ncalls"].sum())) <= 1
