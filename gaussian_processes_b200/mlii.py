"""Batched type-II maximum-likelihood search (``fit_MLII``).

The reference removed its ``fit_MLII`` in v1.0.3 (CHANGELOG.md:16-19) and the v1.0.5
tree holds no implementation, so the *search* is specified here (parity of the search
is unpinned; parity of every per-candidate ``log_lh`` / ``dloglh_dtheta`` is pinned
against the reference's GP properties):

    given candidates theta_b (b < B), evaluate log p(y | x, theta_b) and its gradient for
    all b, and select  argmax_b log_lh[b]  with ties -> lowest b and NaN ordered last.

Candidates are independent, so they shard across GPUs with no data-path collective:
rank r evaluates rows [r*B/G, (r+1)*B/G) and one all-gather of B/G x (n_theta + 2)
doubles per rank assembles the table on every rank (torch.distributed: NCCL on GPUs,
gloo in the CPU tests of the sharding logic).
"""
import numpy as np

from . import engine as _engine

__all__ = ["fit_MLII", "batch_eval", "shard_bounds", "select_best", "MLIIResult"]


class MLIIResult(object):
    """Outcome of a batched search."""

    def __init__(self, best_index, candidates, log_lh, dloglh):
        self.best_index = int(best_index)
        self.candidates = candidates
        self.log_lh = log_lh
        self.dloglh_dtheta = dloglh

    @property
    def best_params(self):
        return self.candidates[self.best_index].copy()

    @property
    def best_log_lh(self):
        return self.log_lh[self.best_index]


def shard_bounds(total, world_size, rank):
    """Contiguous block partition of ``total`` units over ``world_size`` ranks."""
    base, rem = divmod(int(total), int(world_size))
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def select_best(log_lh):
    """argmax with ties -> lowest index, NaN ordered last, -inf allowed."""
    key = np.where(np.isnan(log_lh), -np.inf, log_lh)
    return int(np.argmax(key))


def _evaluator(gp):
    """One evaluator (device buffers + workspace) per GP object; new x / y arrays of the
    same length are re-uploaded into it instead of rebuilding it."""
    ev = getattr(gp, "_batch_ev", None)
    shape_key = (type(gp.K).KIND, gp._x.size)
    data_key = (id(gp._x), id(gp._y))
    if ev is None or ev[0] != shape_key:
        ev = [shape_key, data_key, _engine.BatchEvaluator(shape_key[0], gp._x, gp._y)]
        gp._batch_ev = ev
    elif ev[1] != data_key:
        ev[2].set_data(gp._x, gp._y)
        ev[1] = data_key
    return ev[2]


def batch_eval(gp, thetas, grad=True):
    thetas = np.ascontiguousarray(thetas, dtype=np.float64)
    if thetas.ndim != 2 or thetas.shape[1] != gp.params.size:
        raise ValueError("thetas must have shape [B, %d]" % gp.params.size)
    llh, g, info = _evaluator(gp).eval(thetas, want_grad=grad)
    return llh, (g if grad else None)


def _gather_rows(local, counts, group=None):
    """all-gather ragged row blocks [c_r, w] -> [sum c_r, w] on every rank."""
    import torch
    import torch.distributed as dist
    world = dist.get_world_size(group)
    width = local.shape[1]
    cmax = max(counts)
    dev = local.device
    buf = torch.zeros(cmax, width, dtype=local.dtype, device=dev)
    buf[:local.shape[0]] = local
    out = [torch.empty_like(buf) for _ in range(world)]
    dist.all_gather(out, buf, group=group)
    return torch.cat([o[:c] for o, c in zip(out, counts)], dim=0)


def fit_MLII(gp, candidates, distributed=None, group=None, set_params=True, evaluate=None):
    """Pick the candidate parameter vector with the largest marginal log likelihood.

    candidates : [B, n_theta] array, rows ordered like ``gp.params`` (identical on every
        rank when distributed).
    distributed : None -> use torch.distributed when it is initialised.
    evaluate : optional ``f(thetas) -> torch tensor [b, 2 + n_theta]`` (log_lh, grad...,
        info); defaults to the CUDA evaluator.  Exists so the sharding logic can be
        exercised on CPU (gloo) without a GPU.
    Returns an :class:`MLIIResult`; with ``set_params`` the GP is moved to the winner.
    """
    import torch
    cand = np.ascontiguousarray(candidates, dtype=np.float64)
    B, nth = cand.shape
    if distributed is None:
        import torch.distributed as dist
        distributed = dist.is_available() and dist.is_initialized()
    if evaluate is None:
        ev = _evaluator(gp)

        def evaluate(th):
            r = ev.eval_device(th, want_grad=True)
            return torch.cat([r[:, :1 + nth], r[:, 7:8]], dim=1)
    if distributed:
        import torch.distributed as dist
        world, rank = dist.get_world_size(group), dist.get_rank(group)
        bounds = [shard_bounds(B, world, r) for r in range(world)]
        lo, hi = bounds[rank]
        if hi > lo:
            local = evaluate(cand[lo:hi])
        else:
            local = None
        counts = [b[1] - b[0] for b in bounds]
        if local is None:
            import torch.distributed as dist  # noqa: F811
            dev = "cuda" if dist.get_backend(group) == "nccl" else "cpu"
            local = torch.zeros(0, 2 + nth, dtype=torch.float64, device=dev)
        table = _gather_rows(local, counts, group)
    else:
        table = evaluate(cand)
    table = table.detach().cpu().numpy()
    llh, grad = table[:, 0].copy(), table[:, 1:1 + nth].copy()
    best = select_best(llh)
    res = MLIIResult(best, cand, llh, grad)
    if set_params:
        gp.params = cand[best]
    return res
