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

__all__ = ["fit_MLII", "batch_eval", "shard_bounds", "select_best", "MLIIResult", "refine"]


class MLIIResult(object):
    """Outcome of a batched search."""

    def __init__(self, best_index, candidates, log_lh, dloglh, refined=None):
        #: index of the winning candidate; -1 when NO candidate has a finite log likelihood (every
        #: Kxx non-PD or below the MIN clamp): the search found nothing and the GP is left untouched
        self.best_index = int(best_index)
        self.candidates = candidates
        self.log_lh = log_lh
        self.dloglh_dtheta = dloglh
        #: None, or dict(params [R, n_theta], log_lh [R], dloglh_dtheta [R, n_theta], start_index [R],
        #: best (row of the winner), evals (candidate evaluations spent)) from the gradient refinement
        self.refined = refined

    @property
    def found(self):
        """False when every candidate evaluated to -inf / NaN."""
        return self.best_index >= 0

    @property
    def best_params(self):
        if not self.found:
            return None
        if self.refined is not None:
            return self.refined["params"][self.refined["best"]].copy()
        return self.candidates[self.best_index].copy()

    @property
    def best_log_lh(self):
        if not self.found:
            return -np.inf
        if self.refined is not None:
            return self.refined["log_lh"][self.refined["best"]]
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


def validate_candidates(cand, names=None):
    """The checks ``Kernel.set_param`` / the ``s`` setter apply to one parameter vector
    (gaussian.py:63-73, periodic.py:67-83, gp.py:192-193), applied to every row up front: a search must
    not end on a winner the GP then refuses.  Kernel parameters must be >= EPS, s >= 0, all finite."""
    eps = np.finfo(np.float64).eps
    bad = ~np.isfinite(cand).all(axis=1) | (cand[:, :-1] < eps).any(axis=1) | (cand[:, -1] < 0)
    if bad.any():
        b = int(np.nonzero(bad)[0][0])
        raise ValueError("invalid candidate %d: %s (kernel parameters must be >= EPS, s >= 0)" % (b, cand[b]))


def _evaluator(gp):
    """One evaluator (device buffers + workspace) per GP object; new x / y arrays of the
    same length are re-uploaded into it instead of rebuilding it."""
    from .gp import fused_kind
    ev = getattr(gp, "_batch_ev", None)
    shape_key = (fused_kind(gp.K), gp._x.size)
    # generation counter bumped by every x / y assignment (GP._reset(data=True)); id() of the arrays
    # is NOT a key: CPython hands a freed array's id to the next allocation
    data_key = gp._data_gen
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
    from .gp import fused_kind
    if fused_kind(gp.K) is None:
        # user-defined kernel: its matrices come from Python methods, so candidates are evaluated one
        # by one through the GP properties (the linear algebra of each still runs on the device)
        trial = gp.copy()
        llh = np.empty(thetas.shape[0])
        g = np.empty(thetas.shape) if grad else None
        for b, th in enumerate(thetas):
            trial.params = th
            llh[b] = trial.log_lh
            if grad:
                g[b] = trial.dloglh_dtheta
        return llh, g
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


def refine(starts, evaluate_np, steps=20, gtol=1e-6, max_halvings=8):
    """Batched BFGS ascent on log p(y | x, theta) from several starting points at once.

    Every round evaluates ONE trial point per restart in a single batched device call, so R
    restarts cost the same number of launches as one.  The search runs in u = log(theta)
    (all parameters must stay positive: h, w, p >= EPS, s > 0), with an Armijo backtracking
    line search per restart; a restart whose trial is not PD (log_lh = -inf / NaN) simply
    backtracks.  Restarts stop individually when |d log_lh / du|_inf <= gtol.

    starts      : [R, n_theta] strictly positive parameter rows
    evaluate_np : f(thetas [b, n_theta]) -> (log_lh [b], dloglh_dtheta [b, n_theta]) numpy
    Returns (theta [R, n_theta], log_lh [R], dloglh_dtheta [R, n_theta], n_evals).
    """
    th = np.array(starts, dtype=np.float64, copy=True)
    R, nth = th.shape
    if R == 0:
        return th, np.empty(0), np.empty((0, nth)), 0
    if np.any(th <= 0):
        raise ValueError("refine: parameters must be strictly positive (search runs in log space)")
    u = np.log(th)
    f, g = evaluate_np(th)
    f = np.where(np.isnan(f), -np.inf, f)
    n_evals = R
    gu = g * th                                     # d f / d u = theta * d f / d theta
    H = np.tile(np.eye(nth), (R, 1, 1))             # inverse-Hessian estimates of -f
    active = np.isfinite(f) & np.all(np.isfinite(gu), axis=1)
    for _ in range(int(steps)):
        active &= np.max(np.abs(gu), axis=1) > gtol
        if not active.any():
            break
        d = np.einsum("rij,rj->ri", H, gu)          # ascent direction
        slope = np.einsum("ri,ri->r", d, gu)
        bad = ~(slope > 0)                          # lost positive-definiteness: restart from steepest ascent
        H[bad] = np.eye(nth)
        d[bad] = gu[bad]
        slope[bad] = np.einsum("ri,ri->r", gu[bad], gu[bad])
        # first trial: unit quasi-Newton step, clipped to a factor e^2 per parameter
        alpha = np.minimum(1.0, 2.0 / np.maximum(np.max(np.abs(d), axis=1), 1e-300))
        pending = active.copy()
        u_new, f_new, g_new = u.copy(), f.copy(), g.copy()
        for _h in range(int(max_halvings)):
            idx = np.nonzero(pending)[0]
            if idx.size == 0:
                break
            ut = u[idx] + alpha[idx, None] * d[idx]
            ft, gt = evaluate_np(np.exp(ut))
            n_evals += idx.size
            ft = np.where(np.isnan(ft), -np.inf, ft)
            ok = ft >= f[idx] + 1e-4 * alpha[idx] * slope[idx]
            ok &= np.all(np.isfinite(gt), axis=1)
            acc = idx[ok]
            u_new[acc], f_new[acc], g_new[acc] = ut[ok], ft[ok], gt[ok]
            pending[acc] = False
            alpha[idx[~ok]] *= 0.5
        moved = active & ~pending
        active &= ~pending                          # line search failed: this restart is done
        if moved.any():
            th_new = np.exp(u_new)
            gu_new = g_new * th_new
            sk = u_new - u
            yk = -(gu_new - gu)                     # gradient difference of -f
            for r in np.nonzero(moved)[0]:
                sy = float(sk[r] @ yk[r])
                if sy > 1e-12 * float(np.linalg.norm(sk[r]) * np.linalg.norm(yk[r])):
                    rho = 1.0 / sy
                    I = np.eye(nth)
                    V = I - rho * np.outer(sk[r], yk[r])
                    H[r] = V @ H[r] @ V.T + rho * np.outer(sk[r], sk[r])
            u[moved], f[moved], g[moved], gu[moved] = u_new[moved], f_new[moved], g_new[moved], gu_new[moved]
    return np.exp(u), f, g, n_evals


def fit_MLII(gp, candidates, distributed=None, group=None, set_params=True, evaluate=None,
             refine_top=0, refine_steps=20, refine_gtol=1e-6):
    """Pick the candidate parameter vector with the largest marginal log likelihood.

    candidates : [B, n_theta] array, rows ordered like ``gp.params`` (identical on every
        rank when distributed).
    distributed : None -> use torch.distributed when it is initialised.
    evaluate : optional ``f(thetas) -> torch tensor [b, 2 + n_theta]`` (log_lh, grad...,
        info); defaults to the CUDA evaluator.  Exists so the sharding logic can be
        exercised on CPU (gloo) without a GPU.
    refine_top : R > 0 -> after the search, the R best candidates are polished by batched BFGS
        ascent on their gradients (:func:`refine`; at most ``refine_steps`` iterations, stopping at
        ``refine_gtol``).  Distributed: rank r refines starts r, r+G, ... and one more all-gather
        assembles the refined rows.  The winner is the best refined point.
    Returns an :class:`MLIIResult`; with ``set_params`` the GP is moved to the winner.
    """
    import torch
    cand = np.ascontiguousarray(candidates, dtype=np.float64)
    if cand.ndim != 2 or cand.shape[1] != gp.params.size:
        raise ValueError("candidates must have shape [B, %d]" % gp.params.size)
    B, nth = cand.shape
    if B == 0:
        raise ValueError("fit_MLII needs at least one candidate")
    validate_candidates(cand)
    if distributed is None:
        import torch.distributed as dist
        distributed = dist.is_available() and dist.is_initialized()
    from .gp import fused_kind
    if evaluate is None and fused_kind(gp.K) is None:
        def evaluate(th):                       # user-defined kernel: per-candidate GP properties
            l, g = batch_eval(gp, th)
            bad = (~np.isfinite(l)).astype(np.float64)
            t = torch.from_numpy(np.concatenate([l[:, None], g, bad[:, None]], axis=1))
            return t.cuda() if torch.cuda.is_available() else t
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
    if not np.isfinite(llh[best]):
        # nothing to pick: every Kxx was non-PD or under the MIN clamp.  The GP keeps its parameters.
        return MLIIResult(-1, cand, llh, grad, None)
    refined = None
    if refine_top and refine_top > 0:
        order = np.argsort(-np.where(np.isnan(llh), -np.inf, llh), kind="stable")
        order = order[np.isfinite(llh[order])][:int(refine_top)]

        def evaluate_np(th):
            t = evaluate(np.ascontiguousarray(th)).detach().cpu().numpy()
            return t[:, 0].copy(), t[:, 1:1 + nth].copy()
        if distributed:
            import torch.distributed as dist
            world, rank = dist.get_world_size(group), dist.get_rank(group)
            mine = order[rank::world]
        else:
            world, rank, mine = 1, 0, order
        th_r, f_r, g_r, n_ev = refine(cand[mine], evaluate_np, steps=refine_steps, gtol=refine_gtol)
        rows = np.concatenate([th_r, f_r[:, None], g_r, mine[:, None].astype(np.float64),
                               np.full((mine.size, 1), float(n_ev))], axis=1).reshape(mine.size, 2 * nth + 3)
        if distributed:
            import torch.distributed as dist
            dev = "cuda" if dist.get_backend(group) == "nccl" else "cpu"
            counts = [len(order[r::world]) for r in range(world)]
            rows = _gather_rows(torch.from_numpy(rows).to(dev), counts, group).cpu().numpy()
        if rows.shape[0]:
            refined = dict(params=rows[:, :nth].copy(), log_lh=rows[:, nth].copy(),
                           dloglh_dtheta=rows[:, nth + 1:2 * nth + 1].copy(),
                           start_index=rows[:, 2 * nth + 1].astype(np.int64),
                           evals=int(rows[:, 2 * nth + 2].max() if distributed else n_ev))
            refined["best"] = select_best(refined["log_lh"])
            if not refined["log_lh"][refined["best"]] >= llh[best]:
                refined = None                       # never return something worse than the search
    res = MLIIResult(best, cand, llh, grad, refined)
    if set_params:
        gp.params = res.best_params
    return res
