"""Ensembles across GPUs: one process per GPU, contiguous trajectory shards, no collective during
the solve; the only communication is the final gather of the per-trajectory results (the
reference has no cross-trajectory state at all, so shards are independent)."""
import numpy as np

from .api import solve_ensemble
from .large import shard_bounds


def solve_ensemble_distributed(alg, F, y0, tspan, *, params=None, group=None, gather=True, device=None,
                               _local_solve=None, **kw):
    """Every rank passes the FULL ensemble y0 [n_traj, dof] (and params [n_traj, n_params] if per
    trajectory); rank r integrates rows shard_bounds(n_traj, world, r) on its own GPU.

    gather=True : every rank returns the full result dict (all_gather of final states and counters);
    gather=False: every rank returns only its shard, plus "rows" = (lo, hi).
    `_local_solve` lets the CPU tests substitute the compute (same signature as solve_ensemble)."""
    import torch
    import torch.distributed as dist

    world = dist.get_world_size(group) if dist.is_available() and dist.is_initialized() else 1
    rank = dist.get_rank(group) if world > 1 else 0
    y0 = np.asarray(y0, dtype=np.float64)
    n = y0.shape[0]
    lo, hi = shard_bounds(n, world, rank)
    if device is None:
        device = (torch.cuda.current_device() if torch.cuda.is_available() else 0)
    solve = _local_solve or (lambda *a, **k: solve_ensemble(*a, device=device, **k))
    p_loc = None if params is None else np.asarray(params, dtype=np.float64)[lo:hi]
    loc = solve(alg, F, y0[lo:hi], tspan, params=p_loc, **kw)
    keys = ("t_final", "y_final", "n_accept", "n_reject", "n_rhs", "status")
    if not gather or world == 1:
        out = {k: loc[k] for k in keys}
        out["rows"] = (lo, hi)
        return out
    # final gather: ragged shards are padded to the largest shard, then trimmed
    m = max(shard_bounds(n, world, r)[1] - shard_bounds(n, world, r)[0] for r in range(world))
    use_cuda = dist.get_backend(group) == "nccl"
    dev = torch.device("cuda", device) if use_cuda else torch.device("cpu")
    out = {}
    for k in keys:
        a = np.asarray(loc[k])
        pad = np.zeros((m,) + a.shape[1:], dtype=a.dtype)
        pad[:hi - lo] = a
        t = torch.from_numpy(pad).to(dev)
        parts = [torch.empty_like(t) for _ in range(world)]
        dist.all_gather(parts, t, group=group)
        rows = []
        for r, part in enumerate(parts):
            l, h = shard_bounds(n, world, r)
            rows.append(part[:h - l].cpu().numpy())
        out[k] = np.concatenate(rows, 0)
    out["rows"] = (0, n)
    return out
