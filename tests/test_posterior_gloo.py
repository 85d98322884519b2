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


# ------------------------------------------------------------------ lower-panel layout (round 2)
def _worker_lower(rank, world, port, m, outdir):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import torch
    import scipy.linalg
    import gaussian_processes_b200 as gpb
    from gaussian_processes_b200.posterior import PanelPlan
    from conftest import load_oracle, synth_xy
    oracle = load_oracle()
    x, y = synth_xy(40, 1)
    kp = (1.0, 0.9, 1.5)
    o = oracle.OracleGP(oracle.PERIODIC, kp, x, y, 0.7)
    Linv = scipy.linalg.solve_triangular(o.Lxx, np.eye(x.size), lower=True)
    xo = np.linspace(-6, 6, m)
    plan = PanelPlan(m, world)
    calls = dict(z=0, panels=[])

    def z_fn(lo, hi, bs):                       # Z_b = K(xo_b, x) L^-T, zero rows beyond the block
        calls["z"] += 1
        Z = np.zeros((bs, x.size))
        if hi > lo:
            Z[:hi - lo] = oracle.K(oracle.PERIODIC, xo[lo:hi], x, kp) @ Linv.T
        return torch.from_numpy(Z)

    def panel_fn(b, bs, zget):
        lo, hi = plan.bounds(b)
        calls["panels"].append(b)
        Zall = np.concatenate([zget(c).numpy()[:plan.bounds(c)[1] - plan.bounds(c)[0]] for c in range(b + 1)], axis=0)
        C = oracle.K(oracle.PERIODIC, xo[lo:hi], xo[:hi], kp) - zget(b).numpy()[:hi - lo] @ Zall.T
        out = np.zeros((bs, (b + 1) * bs))
        out[:hi - lo, :hi] = C
        return torch.from_numpy(out)

    gp = gpb.GP(gpb.PeriodicKernel(*kp), x, y, s=0.7)
    kw = dict(cov_layout="lower", mean_fn=lambda xb: o.mean(xb) if len(xb) else np.empty(0), z_fn=z_fn, panel_fn=panel_fn)
    tm = {}
    mean, panels, pl = gpb.sharded_posterior(gp, xo, timings=tm, **kw)
    first_blocks, first_z = list(calls["panels"]), calls["z"]
    mean2, full, _ = gpb.sharded_posterior(gp, xo, gather_cov=True, **kw)
    mean3, none, _ = gpb.sharded_posterior(gp, xo, want_cov=False, **kw)
    assert none is None and np.array_equal(mean, mean2) and np.array_equal(mean, mean3)
    assert set(tm) >= {"z", "allgather", "panels"} and pl.nb == plan.nb
    np.savez(os.path.join(outdir, "l%d.npz" % rank), mean=mean, full=full, nz=first_z, blocks=np.array(first_blocks),
             los=np.array([p[0] for p in panels]), his=np.array([p[1] for p in panels]),
             **{"p%d" % i: np.asarray(p[2]) for i, p in enumerate(panels)})
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("m", [1, 9, 300, 2500])
def test_sharded_posterior_lower_panels_world2_gloo(tmp_path, m):
    from gaussian_processes_b200.posterior import PanelPlan
    port = 33500 + (os.getpid() % 2000) + (m % 97)
    mp.spawn(_worker_lower, args=(2, port, m, str(tmp_path)), nprocs=2, join=True)
    oracle = load_oracle()
    x, y = synth_xy(40, 1)
    o = oracle.OracleGP(oracle.PERIODIC, (1.0, 0.9, 1.5), x, y, 0.7)
    xo = np.linspace(-6, 6, m)
    ref_mean, ref_cov = o.mean(xo), o.cov(xo)
    scale = np.max(np.abs(ref_cov))
    plan = PanelPlan(m, 2)
    covered = np.zeros((m, m), dtype=bool)
    cost = []
    for r in range(2):
        d = dict(np.load(tmp_path / ("l%d.npz" % r)))
        assert np.allclose(d["mean"], ref_mean, rtol=1e-13, atol=1e-14)
        assert np.max(np.abs(d["full"] - ref_cov)) <= 1e-12 * scale              # Z Z^T form vs explicit inverse
        assert sorted(d["blocks"].tolist()) == [b for b in plan.mine(r) if plan.bounds(b)[1] > plan.bounds(b)[0]]
        for i, (lo, hi) in enumerate(zip(d["los"], d["his"])):
            P = d["p%d" % i]
            assert P.shape == (hi - lo, hi)
            assert np.max(np.abs(P - ref_cov[lo:hi, :hi])) <= 1e-12 * scale
            covered[lo:hi, :hi] = True
        cost.append(sum(b + 1 for b in plan.mine(r)))
    assert covered[np.tril_indices(m)].all()                                     # the panels tile the lower triangle
    assert cost[0] == cost[1]                                                    # equal block products per rank
