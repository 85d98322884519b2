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
    params_search = gp.params
    nsearch = list(calls)
    # the same search followed by the batched BFGS polish of the 3 best candidates (starts shard over ranks)
    res2 = gpb.fit_MLII(gp, cand, evaluate=evaluate, refine_top=3, refine_steps=25, refine_gtol=1e-7)
    np.savez(os.path.join(outdir, "r%d.npz" % rank), best=res.best_index, llh=res.log_lh,
             grad=res.dloglh_dtheta, params=params_search, ncalls=np.array(nsearch), cand=cand,
             ref_params=res2.refined["params"], ref_llh=res2.refined["log_lh"], ref_grad=res2.refined["dloglh_dtheta"],
             ref_start=res2.refined["start_index"], ref_best=res2.refined["best"], params2=gp.params)
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
    # refinement: both ranks hold the same refined table; starts are the 3 best of the search; every
    # refined point is at least as good as its start and (nearly) stationary in log-parameters
    for k in ("ref_params", "ref_llh", "ref_grad", "ref_start", "params2"):
        assert np.array_equal(r0[k], r1[k]), k
    top3 = np.argsort(-llh, kind="stable")[:3]
    assert sorted(r0["ref_start"].tolist()) == sorted(top3.tolist())
    assert np.all(r0["ref_llh"] >= llh[r0["ref_start"]] - 1e-12)
    assert np.max(np.abs(r0["ref_grad"] * r0["ref_params"])) < 1e-4
    assert np.array_equal(r0["params2"], r0["ref_params"][int(r0["ref_best"])])
    o = oracle.OracleGP(oracle.GAUSSIAN, r0["params2"][:-1], x, y, r0["params2"][-1])
    assert abs(float(o.log_lh) - r0["ref_llh"][int(r0["ref_best"])]) <= 1e-9 * abs(float(o.log_lh))
