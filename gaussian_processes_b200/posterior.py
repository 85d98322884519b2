"""Posterior mean / covariance with the test points sharded over GPUs.

A block of test points needs no peer data: rank r owns ``xo[lo:hi)`` (contiguous block partition,
``mlii.shard_bounds``), computes ``mean(xo[lo:hi])`` (gp/gp.py:574-597) and the row block
``cov(xo)[lo:hi, :]`` (gp.py:599-625, ``GP.cov_rows``) on its own GPU from its own copy of the
factorisation, and the only exchange is the assembly of the results: one ragged all-gather of
``M/G`` doubles per rank for the mean and -- only when the caller asks for the full matrix on every
rank -- one of the ``[m_r, M]`` row blocks (NCCL over NVLink; gloo in the CPU tests of this logic).
"""
import numpy as np

from .mlii import shard_bounds, _gather_rows

__all__ = ["sharded_posterior"]


def sharded_posterior(gp, xo, want_cov=True, gather_cov=False, group=None, distributed=None,
                      mean_fn=None, cov_rows_fn=None):
    """Posterior at ``xo`` (identical on every rank) with the test points partitioned over ranks.

    Returns ``(mean, cov, (lo, hi))``: ``mean`` is the full ``[M]`` vector on every rank; ``cov`` is
    ``None`` (``want_cov=False``), this rank's row block ``cov(xo)[lo:hi, :]`` (default: the
    covariance stays sharded, 8 M^2 / G bytes per rank) or the full ``[M, M]`` matrix on every rank
    (``gather_cov=True``).  ``mean_fn(xo_block)`` / ``cov_rows_fn(xo, lo, hi)`` default to the GP's
    CUDA path; they exist so the partition / gather logic can run on CPU (gloo) without a GPU.
    """
    import torch
    xo = np.ascontiguousarray(xo, dtype=np.float64).reshape(-1)
    m = int(xo.size)
    if distributed is None:
        import torch.distributed as dist
        distributed = dist.is_available() and dist.is_initialized()
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
