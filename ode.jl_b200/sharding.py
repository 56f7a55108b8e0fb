"""Ensembles across GPUs: one process per GPU, contiguous trajectory shards, no collective during
the solve; the only communication is the final gather of the per-trajectory results (the
reference has no cross-trajectory state at all, so shards are independent).

The final gather ("per-device D2H into disjoint slices of the caller's buffers"): the ranks
of one node share ONE host buffer (a file in /dev/shm mapped by every rank, page-locked for its own slice); every
rank's library call copies its results device -> host straight into its slice, and after a barrier every rank
sees the whole ensemble.  No device all-gather, no padding, nothing crosses NVLink or the network.  Ranks on
different hosts fall back to an all_gather of the shards."""
import os
import shutil
import socket
import tempfile

import numpy as np

from .api import solve_ensemble
from .large import shard_bounds

_KEYS = (("t_final", np.float64, 0), ("y_final", np.float64, 1), ("n_accept", np.int32, 0), ("n_reject", np.int32, 0),
         ("n_rhs", np.int32, 0), ("status", np.int32, 0))
_counter = [0]


class SharedHostBlock:
    """`nbytes` of host memory shared by the ranks of one node (a file in /dev/shm mapped by every rank).
    Collective constructor and `close()`; `pin(array)` page-locks a slice for this rank's device copies."""

    def __init__(self, nbytes, group=None):
        import torch.distributed as dist
        self.group = group
        self.world = dist.get_world_size(group)
        self.rank = dist.get_rank(group)
        hosts = [None] * self.world
        dist.all_gather_object(hosts, socket.gethostname(), group=group)
        if len(set(hosts)) != 1:
            raise RuntimeError("ranks on different hosts cannot share a host buffer")
        nbytes = max(int(nbytes), 4096)
        self._bar_off = (nbytes + 4095) & ~4095  # one cache line per rank behind the payload: host_barrier()
        nbytes = self._bar_off + 64 * self.world
        self._gen = 0
        name = [None]
        if self.rank == 0:
            _counter[0] += 1
            for d in ("/dev/shm", tempfile.gettempdir()):  # memory-backed if it fits (containers may keep /dev/shm tiny)
                try:
                    if shutil.disk_usage(d).free < nbytes + (64 << 20):
                        continue
                    name[0] = os.path.join(d, "ode_b200_%d_%d" % (os.getpid(), _counter[0]))
                    with open(name[0], "wb") as f:
                        f.truncate(nbytes)
                    break
                except OSError:
                    name[0] = None
        dist.broadcast_object_list(name, src=dist.get_global_rank(group, 0) if group is not None else 0, group=group)
        self.path = name[0]
        if self.path is None:
            raise RuntimeError("no room for a %d-byte shared host buffer" % nbytes)
        self.map = np.memmap(self.path, dtype=np.uint8, mode="r+", shape=(nbytes,))
        self._pinned = []
        dist.barrier(group)

    def host_barrier(self, timeout_s=60.0):
        """Barrier of the node's ranks through the shared block itself (a generation counter per rank, spun on by the
        CPU): the "every slice has landed" point of the final gather without a device collective -- a NCCL barrier
        costs a kernel launch and a stream synchronise per rank."""
        import time
        self._gen += 1
        flags = self.map[self._bar_off:self._bar_off + 64 * self.world].view(np.int64).reshape(self.world, 8)
        flags[self.rank, 0] = self._gen
        t0 = time.perf_counter()
        while int(flags[:, 0].min()) < self._gen:
            if time.perf_counter() - t0 > timeout_s:
                raise RuntimeError("host barrier timed out (a rank left?)")

    def view(self, dtype, shape, offset=0):
        cnt = int(np.prod(shape)) * np.dtype(dtype).itemsize
        return self.map[offset:offset + cnt].view(dtype).reshape(shape)

    def pin(self, a):
        """Page-lock the memory of array `a` (a view of this block) so device copies run at full PCIe rate."""
        try:
            import torch
            if a.nbytes and torch.cuda.is_available():
                if int(torch.cuda.cudart().cudaHostRegister(a.ctypes.data, a.nbytes, 0)) == 0:
                    self._pinned.append(a.ctypes.data)
                    return True
        except Exception:  # noqa: BLE001 -- pageable memory works too, staged
            pass
        return False

    def close(self):
        import torch.distributed as dist
        if self.map is None:
            return
        if self._pinned:
            import torch
            for p in self._pinned:
                torch.cuda.cudart().cudaHostUnregister(p)
            self._pinned = []
        dist.barrier(self.group)
        self.map = None
        if self.rank == 0:
            try:
                os.unlink(self.path)
            except OSError:
                pass


class SharedEnsembleResult:
    """Result arrays of a whole ensemble [n_traj] in host memory shared by the ranks of one node.
    Collective constructor and `close()`.  `arrays[k]` is the full array, `mine[k]` this rank's slice (what the
    rank's solve writes); reusable across solves of the same shape."""

    def __init__(self, n_traj, dof, group=None, pin=True):
        self.n, self.dof = int(n_traj), int(dof)
        offs, total = {}, 0
        for k, dt, vec in _KEYS:
            offs[k] = total
            total += self.n * (self.dof if vec else 1) * np.dtype(dt).itemsize
            total = (total + 4095) & ~4095
        self.block = SharedHostBlock(total, group)
        self.group, self.world, self.rank = group, self.block.world, self.block.rank
        lo, hi = shard_bounds(self.n, self.world, self.rank)
        self.rows = (lo, hi)
        self.arrays, self.mine = {}, {}
        for k, dt, vec in _KEYS:
            a = self.block.view(dt, (self.n, self.dof) if vec else (self.n,), offs[k])
            self.arrays[k] = a
            self.mine[k] = a[lo:hi]
            if pin:  # this rank's slices: device -> host copies at full PCIe rate, asynchronously
                self.block.pin(self.mine[k])

    def close(self):
        self.arrays, self.mine = {}, {}
        self.block.close()


def _gather_collective(loc, n, lo, hi, world, group, device):
    """Fallback for ranks on different hosts: all_gather of the (padded) shards."""
    import torch
    import torch.distributed as dist
    m = max(shard_bounds(n, world, r)[1] - shard_bounds(n, world, r)[0] for r in range(world))
    use_cuda = dist.get_backend(group) == "nccl"
    dev = torch.device("cuda", device) if use_cuda else torch.device("cpu")
    out = {}
    for k, _, _ in _KEYS:
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
    return out


def solve_ensemble_distributed(alg, F, y0, tspan, *, params=None, group=None, gather=True, device=None,
                               shared=None, _local_solve=None, **kw):
    """Every rank passes the FULL ensemble y0 [n_traj, dof] (and params [n_traj, n_params] if per
    trajectory); rank r integrates rows shard_bounds(n_traj, world, r) on its own GPU.

    gather=True : every rank returns the full result dict.  The ranks of one node write their shards into one
                  shared host buffer (`shared`: a SharedEnsembleResult to reuse across calls -- the returned
                  arrays then are views of it -- or None to make a temporary one and return copies);
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
    p_loc = None if params is None else np.asarray(params, dtype=np.float64)[lo:hi]
    keys = [k for k, _, _ in _KEYS]
    if not gather or world == 1:
        solve = _local_solve or (lambda *a, **k: solve_ensemble(*a, device=device, **k))
        loc = solve(alg, F, y0[lo:hi], tspan, params=p_loc, **kw)
        out = {k: loc[k] for k in keys}
        out["rows"] = (lo, hi)
        return out
    own = None
    if shared is None:
        try:
            own = shared = SharedEnsembleResult(n, y0.shape[1], group, pin=_local_solve is None)
        except RuntimeError:
            shared = None
    if shared is None:  # different hosts
        solve = _local_solve or (lambda *a, **k: solve_ensemble(*a, device=device, **k))
        loc = solve(alg, F, y0[lo:hi], tspan, params=p_loc, **kw)
        out = _gather_collective(loc, n, lo, hi, world, group, device)
        out["rows"] = (0, n)
        return out
    if _local_solve is None:  # the library copies device -> host straight into this rank's slice
        solve_ensemble(alg, F, y0[lo:hi], tspan, params=p_loc, device=device, out=shared.mine, **kw)
    else:
        loc = _local_solve(alg, F, y0[lo:hi], tspan, params=p_loc, **kw)
        for k in keys:
            shared.mine[k][...] = loc[k]
    shared.block.host_barrier()  # every slice is complete: this is the whole "gather" (no device collective)
    if own is None:
        out = dict(shared.arrays)
    else:
        out = {k: np.array(shared.arrays[k]) for k in keys}
        own.close()
    out["rows"] = (0, n)
    return out
