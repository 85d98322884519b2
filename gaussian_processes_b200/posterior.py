"""Posterior mean / covariance with the test points sharded over GPUs.

Two layouts of the sharded covariance:

``cov_layout="lower"`` (the scalable one).  cov(xo) = K(xo,xo) - Z Z^T with Z = K(xo,x) L^-T
(gp/gp.py:599-625 through the Cholesky factor) is symmetric, so only block rows of its LOWER triangle
are computed.  The test points are cut into ``nb`` equal blocks; block row b costs (b + 1) block
products, so blocks are dealt out in pairs (b, nb-1-b) of equal cost, round-robin over the ranks:
every rank does 1/G of N^2 M (its rows of Z) + N M^2 / 2 (its panels) flop -- half of what
full-width row shards cost.  The one exchange is an all-gather of Z over NVLink (8 M N bytes in
total, 1.07 GB at N = 8192, M = 16384): rank r then holds every Z_c its panels multiply by.  Each rank
factors its own copy of Kxx (fit once, predict many: the factorisation is not part of a prediction).

``cov_layout="rows"`` (round 1): rank r owns the full-width row block cov(xo)[lo:hi, :]
(``GP.cov_rows``); needs no exchange at all, costs 2 N^2 m_r + 2 N M m_r flop per rank.

The mean needs no peer data in either layout: rank r evaluates its own test points and one all-gather
of M doubles assembles the vector.  NCCL over NVLink on GPUs; gloo in the CPU tests of this logic.
"""
import numpy as np

from .mlii import shard_bounds, _gather_rows

__all__ = ["sharded_posterior", "panel_plan"]


class PanelPlan(object):
    """nb blocks of bs test points; ``owner[b]`` = rank of block row b; ``mine(rank)`` ascending."""

    def __init__(self, m, world, unit=128):
        self.m, self.world = int(m), int(world)
        per = 2 * max(1, -(-8 // self.world))                    # blocks per rank (>= 2, 16 in total up to 8 ranks)
        self.nb = per * self.world
        self.per = per
        bs = -(-max(self.m, 1) // self.nb)
        self.bs = -(-bs // unit) * unit
        self.owner = [min(b, self.nb - 1 - b) % self.world for b in range(self.nb)]

    def bounds(self, b):
        lo = min(self.m, b * self.bs)
        return lo, min(self.m, lo + self.bs)

    def mine(self, rank):
        return [b for b in range(self.nb) if self.owner[b] == rank]

    def slot(self, b):
        """Position of block b among its owner's blocks (its place in the owner's all-gather chunk)."""
        return self.mine(self.owner[b]).index(b)


def panel_plan(m, world, unit=128):
    return PanelPlan(m, world, unit)


def _cuda_fns(gp, xo):
    """Device implementations of the two per-block computations (everything stays in HBM)."""
    from . import device as D
    e = gp._engine()
    dxo = D.to_device(xo)
    m = int(xo.size)

    def z_fn(lo, hi, bs):
        if hi <= lo:
            return D.zeros(bs, e.npad)
        return e.z_block(dxo[lo:hi], hi - lo, bs)

    def panel_fn(b, bs, zget):
        return e.cov_panel(dxo, m, b, bs, zget)
    return z_fn, panel_fn


def sharded_posterior(gp, xo, want_cov=True, gather_cov=False, group=None, distributed=None,
                      mean_fn=None, cov_rows_fn=None, cov_layout="rows", z_fn=None, panel_fn=None,
                      host=True, timings=None):
    """Posterior at ``xo`` (identical on every rank) with the test points partitioned over ranks.

    ``cov_layout="rows"``: returns ``(mean, cov, (lo, hi))``: ``mean`` is the full ``[M]`` vector on every
    rank; ``cov`` is ``None`` (``want_cov=False``), this rank's row block ``cov(xo)[lo:hi, :]`` or the
    full matrix on every rank (``gather_cov=True``).

    ``cov_layout="lower"``: returns ``(mean, panels, plan)``: ``panels`` is the list of this rank's block
    rows of the lower triangle, ``(lo, hi, C[lo:hi, 0:hi])`` (numpy arrays; device tensors with
    ``host=False``), ``plan`` the :class:`PanelPlan`; with ``gather_cov=True`` ``panels`` is the full
    symmetric ``[M, M]`` matrix on every rank instead.

    ``mean_fn / cov_rows_fn / z_fn / panel_fn`` default to the GP's CUDA path; they exist so the partition
    and exchange logic can run on CPU (gloo) without a GPU.  ``timings`` (a dict) receives the seconds
    spent in ``z``, ``allgather`` and ``panels`` when given (device-synchronised).
    """
    import torch
    xo = np.ascontiguousarray(xo, dtype=np.float64).reshape(-1)
    m = int(xo.size)
    if distributed is None:
        import torch.distributed as dist
        distributed = dist.is_available() and dist.is_initialized()
    if cov_layout == "lower":
        return _sharded_lower(gp, xo, want_cov, gather_cov, group, distributed, mean_fn, z_fn, panel_fn, host, timings)
    if cov_layout != "rows":
        raise ValueError("cov_layout must be 'rows' or 'lower'")
    mean_fn = mean_fn or gp.mean
    cov_rows_fn = cov_rows_fn or gp.cov_rows
    if not distributed:
        mean = np.asarray(mean_fn(xo))
        cov = np.asarray(cov_rows_fn(xo, 0, m)) if want_cov else None
        return mean, cov, (0, m)
    import torch.distributed as dist
    world, rank = dist.get_world_size(group), dist.get_rank(group)
    # shards start on multiples of 128 test points (the tile edge) whenever there are enough of them,
    # so that a shard's own diagonal block of the covariance takes the symmetric tile path of cov_rows
    unit = 128 if m >= 128 * world else 1
    nblk = (m + unit - 1) // unit
    bounds = []
    for r in range(world):
        b0, b1 = shard_bounds(nblk, world, r)
        bounds.append((min(m, b0 * unit), min(m, b1 * unit)))
    counts = [b[1] - b[0] for b in bounds]
    lo, hi = bounds[rank]
    dev = "cuda" if dist.get_backend(group) == "nccl" else "cpu"
    local_mean = np.asarray(mean_fn(xo[lo:hi]), dtype=np.float64).reshape(-1, 1)
    mean = _gather_rows(torch.from_numpy(np.ascontiguousarray(local_mean)).to(dev), counts, group)
    mean = mean.cpu().numpy().reshape(-1)
    cov = None
    if want_cov:
        cov = np.asarray(cov_rows_fn(xo, lo, hi), dtype=np.float64).reshape(hi - lo, m)
        if gather_cov:
            cov = _gather_rows(torch.from_numpy(np.ascontiguousarray(cov)).to(dev), counts, group).cpu().numpy()
    return mean, cov, (lo, hi)


def _sync(t):
    import torch
    if t.is_cuda:
        torch.cuda.synchronize()


def _sharded_lower(gp, xo, want_cov, gather_cov, group, distributed, mean_fn, z_fn, panel_fn, host, timings):
    import time
    import torch
    m = int(xo.size)
    if distributed:
        import torch.distributed as dist
        world, rank = dist.get_world_size(group), dist.get_rank(group)
    else:
        world, rank = 1, 0
    plan = PanelPlan(m, world)
    mine = plan.mine(rank)
    bs = plan.bs
    mean_fn = mean_fn or gp.mean
    if z_fn is None or panel_fn is None:
        z_fn, panel_fn = _cuda_fns(gp, xo)

    # ---- mean: own test points, one all-gather of equal chunks (per * bs doubles per rank) -------------
    own_idx = np.concatenate([np.arange(*plan.bounds(b)) for b in mine]) if mine else np.empty(0, dtype=np.int64)
    local = np.zeros(plan.per * bs)
    if own_idx.size:
        vals = np.asarray(mean_fn(xo[own_idx]), dtype=np.float64).reshape(-1)
        pos = 0
        for q, b in enumerate(mine):
            lo, hi = plan.bounds(b)
            local[q * bs:q * bs + (hi - lo)] = vals[pos:pos + (hi - lo)]
            pos += hi - lo
    if distributed:
        import torch.distributed as dist
        dev = "cuda" if dist.get_backend(group) == "nccl" else "cpu"
        lt = torch.from_numpy(local).to(dev)
        allm = torch.empty(world * plan.per * bs, dtype=torch.float64, device=dev)
        dist.all_gather_into_tensor(allm, lt, group=group)
        allm = allm.cpu().numpy().reshape(world, plan.per, bs)
    else:
        allm = local.reshape(1, plan.per, bs)
    mean = np.empty(m)
    for b in range(plan.nb):
        lo, hi = plan.bounds(b)
        mean[lo:hi] = allm[plan.owner[b], plan.slot(b), :hi - lo]
    if not want_cov:
        return mean, None, plan

    # ---- Z rows of the own blocks, then the one exchange: all-gather of Z ------------------------------
    t0 = time.perf_counter()
    zl = [z_fn(plan.bounds(b)[0], plan.bounds(b)[1], bs) for b in mine]
    zloc = torch.stack(zl, dim=0) if len(zl) > 1 else zl[0].unsqueeze(0)          # [per, bs, k]
    del zl
    if timings is not None:
        _sync(zloc)
        timings["z"] = time.perf_counter() - t0
        t0 = time.perf_counter()
    if distributed:
        import torch.distributed as dist
        zall = torch.empty((world,) + tuple(zloc.shape), dtype=zloc.dtype, device=zloc.device)
        dist.all_gather_into_tensor(zall.view(world * zloc.shape[0] * bs, -1), zloc.view(zloc.shape[0] * bs, -1), group=group)
    else:
        zall = zloc.unsqueeze(0)
    if timings is not None:
        _sync(zall)
        timings["allgather"] = time.perf_counter() - t0
        timings["allgather_bytes"] = int(zall.numel() * 8)
        t0 = time.perf_counter()
    slots = [plan.slot(b) for b in range(plan.nb)]

    def zget(c):
        return zall[plan.owner[c], slots[c]]

    # ---- this rank's block rows of the lower triangle -------------------------------------------------
    panels = []
    if host and zall.is_cuda:
        from . import device as D
        cs = D.copy_stream()
        for b in mine:
            lo, hi = plan.bounds(b)
            if hi <= lo:
                continue
            C = panel_fn(b, bs, zget)
            out, pinned = D.host_array(hi - lo, hi)
            if pinned:                       # the panel's DMA overlaps the next panel's products
                cs.wait_stream(torch.cuda.current_stream())
                D.download_2d(C, hi - lo, hi, out=out, pinned=True, stream=cs, sync=False)
                C.record_stream(cs)
            else:
                D.download_2d(C, hi - lo, hi, out=out, pinned=False)
            panels.append((lo, hi, out))
        cs.synchronize()
        torch.cuda.current_stream().wait_stream(cs)
    else:
        for b in mine:
            lo, hi = plan.bounds(b)
            if hi <= lo:
                continue
            C = panel_fn(b, bs, zget)
            C = C[:hi - lo, :hi]
            panels.append((lo, hi, C.cpu().numpy() if (host and isinstance(C, torch.Tensor)) else C))
    if timings is not None:
        _sync(zall)
        timings["panels"] = time.perf_counter() - t0
    if not gather_cov:
        return mean, panels, plan
    # full symmetric matrix on every rank: own panels (+ their mirror images left of the diagonal block), summed
    full = np.zeros((m, m))
    for lo, hi, C in panels:
        C = np.asarray(C.cpu().numpy() if isinstance(C, torch.Tensor) else C)
        full[lo:hi, :hi] = C
        full[:lo, lo:hi] = C[:, :lo].T
    if distributed:
        import torch.distributed as dist
        dev = "cuda" if dist.get_backend(group) == "nccl" else "cpu"
        ft = torch.from_numpy(full).to(dev)
        dist.all_reduce(ft, group=group)
        full = ft.cpu().numpy()
    return mean, full, plan
