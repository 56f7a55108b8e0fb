"""The large-system regime: one method-of-lines ODE system of N unknowns.

`solve_large` is the call with host buffers (ode_b200_solve_large); with
`n_gpus=N` the library itself splits the grid over N devices of this process
(ode_b200_solve_large_multi).
`solve_large_distributed` is the domain-decomposed path with one process per GPU:
every rank owns a contiguous slab of the periodic grid plus ghost cells; per step
ATTEMPT there is exactly one small all-gather (ODE_B200_XCHG_DOUBLES doubles per
rank) that carries the double-double error partials, the NaN flag and the
slab-edge values used to refresh the ghosts when the step is accepted.  Every rank
then runs the same controller arithmetic on the same gathered data, so the
accept/reject decision and the new dt are identical everywhere without a broadcast
(reference semantics: oderk_adapt, /root/reference src/runge_kutta.jl:232-369,
where the whole state is one Vector in one process).

Two ways to do that all-gather:
  exchange="peer"  the attempt kernel itself writes the record into every peer's
                   HBM (CUDA IPC mappings over NVLink) and its last thread block
                   runs the controller; the whole integration is ONE C call per
                   rank (ode_b200_slab_run), no Python in the loop.  Ranks of one
                   node only.
  exchange="nccl"  torch.distributed.all_gather_into_tensor between the attempt
                   and the control kernel, driven by `integrate_slab` below.
"auto" tries "peer" and falls back to "nccl".

`integrate_slab` is engine-agnostic: the CUDA engine below is the product; the
CPU tests inject a stand-in engine to exercise the collective plumbing with gloo.
"""
import ctypes as C

import numpy as np

from . import _lib
from ._lib import LargeStats, check, lib, XCHG_DOUBLES
from .api import make_opts, _alg_name, _raise_like_reference
from .rhs import DeviceRHS


def shard_bounds(n_global, world, rank):
    """Contiguous slab [lo, hi) of rank `rank`; the first n_global % world ranks get one extra point."""
    base, rem = divmod(int(n_global), int(world))
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


_FIXED_RK = ("ode1", "ode2_midpoint", "ode2_heun", "ode4", "ode4ms", "ode5ms", "ode_ms1", "ode_ms2", "ode_ms3", "ode_ms4",
             "ode_ms5")  # fixed-step methods with a dof = N kernel


def _solve_large_fixed(name, F, y0, ts, points, device):
    """ode1 / ode2_midpoint / ode2_heun / ode4 on a dof = N state (src/runge_kutta.jl:176-212)."""
    ts = np.ascontiguousarray(ts)
    o = make_opts(name, F, ts, dof=min(y0.size, 2 ** 31 - 1), device=device, points="all")
    p = F.param_array()
    if p is None:
        p = np.zeros(1)
    st = LargeStats()
    dp = _lib.dp
    rows = None if points == "final" else np.empty((ts.size, y0.size))
    yf = np.empty_like(y0)
    check(lib().ode_b200_solve_large_fixed(C.byref(o), y0.size, y0.ctypes.data_as(dp), p.ctypes.data_as(dp),
                                           ts.ctypes.data_as(dp), ts.size,
                                           None if rows is None else rows.ctypes.data_as(dp), yf.ctypes.data_as(dp),
                                           C.byref(st)))
    r = dict(y=yf if rows is None else rows, t=ts.copy(), t_final=st.t_final, n_accept=st.n_accept, n_reject=0,
             n_rhs=st.n_rhs, status=st.status, kernel_ms=st.kernel_ms)
    return r


def solve_large(alg, F, y0, tspan, *, trace=False, trace_cap=65536, out=None, **kw):
    """One GPU, host buffers in and out.  Returns dict(y, t_final, n_accept, n_reject, n_rhs, status, kernel_ms[, trace]).
    `out` (optional) is a caller-owned float64 array for the final state, e.g. pinned memory."""
    if not isinstance(F, DeviceRHS) or not getattr(F, "is_stencil", False):
        raise TypeError("the large-system path takes DeviceRHS('reaction_diffusion', kappa, r) or a register_stencil token")
    y0 = np.ascontiguousarray(np.asarray(y0, dtype=np.float64))
    ts = np.asarray(tspan, dtype=np.float64)
    kw.setdefault("points", "final")
    if _alg_name(alg) in _FIXED_RK:  # oderk_fixed on the tspan grid: one state per tspan entry, no keywords
        return _solve_large_fixed(_alg_name(alg), F, y0, ts, kw.get("points"), kw.get("device", 0))
    o = make_opts(_alg_name(alg), F, ts, dof=min(y0.size, 2 ** 31 - 1), **kw)
    p = F.param_array()
    if p is None:
        p = np.zeros(1)  # a stencil without parameters
    if out is None:
        out = np.empty_like(y0)
    elif out.dtype != np.float64 or out.size != y0.size or not out.flags.c_contiguous:
        raise ValueError("out must be a C-contiguous float64 array of y0's size")
    st = LargeStats()
    tr = np.zeros((trace_cap, 4)) if trace else None
    tn = C.c_int64(0)
    dp = _lib.dp
    specified = o.points == _lib.POINTS["specified"]
    all_points = o.points == _lib.POINTS["all"]
    try:
        if all_points:  # every accepted step (+ interior tspan points), src/runge_kutta.jl:327-340
            ts = np.ascontiguousarray(ts)
            cap = ts.size + 64
            rows = C.c_int64(0)
            while True:
                pt = np.empty(cap)
                pts = np.empty((cap, y0.size))
                check(lib().ode_b200_solve_large_all(C.byref(o), y0.size, y0.ctypes.data_as(dp), p.ctypes.data_as(dp),
                                                     ts.ctypes.data_as(dp), ts.size, cap, pt.ctypes.data_as(dp),
                                                     pts.ctypes.data_as(dp), C.byref(rows), C.byref(st),
                                                     tr.ctypes.data_as(dp) if trace else None,
                                                     trace_cap if trace else 0, C.byref(tn)))
                if rows.value <= cap:
                    break
                cap = rows.value  # deterministic: rerun with the reported size
            out = pts[:rows.value]
            ts = pt[:rows.value]
            specified = True
        elif specified:  # tout = tspan, yout[i] = Hermite value at tspan[i] (src/runge_kutta.jl:321-326)
            ts = np.ascontiguousarray(ts)
            pts = np.empty((ts.size, y0.size))
            check(lib().ode_b200_solve_large_points(C.byref(o), y0.size, y0.ctypes.data_as(dp), p.ctypes.data_as(dp),
                                                    ts.ctypes.data_as(dp), ts.size, pts.ctypes.data_as(dp),
                                                    C.byref(st), tr.ctypes.data_as(dp) if trace else None,
                                                    trace_cap if trace else 0, C.byref(tn)))
            out = pts
        else:  # (opts.n_gpus > 1: the library splits the grid over that many devices)
            check(lib().ode_b200_solve_large(C.byref(o), y0.size, y0.ctypes.data_as(dp), p.ctypes.data_as(dp),
                                             float(ts[0]), float(ts[-1]), out.ctypes.data_as(dp), C.byref(st),
                                             tr.ctypes.data_as(dp) if trace else None, trace_cap if trace else 0,
                                             C.byref(tn)))
    except _lib.OdeB200Error as e:
        _raise_like_reference(e)
    res = dict(y=out, t=ts if specified else None, t_final=st.t_final, n_accept=st.n_accept, n_reject=st.n_reject, n_rhs=st.n_rhs,
               status=st.status, kernel_ms=st.kernel_ms)
    if trace:
        res["trace"] = tr[:min(tn.value, trace_cap)]
    return res


class CudaSlabEngine:
    """A slab session of libode_b200.so (device pointers, caller's CUDA stream)."""

    def __init__(self, alg, F, n_local, n_global, rank, world, tspan, stream_ptr, device=0, **kw):
        ts = np.asarray(tspan, dtype=np.float64)
        kw.setdefault("points", "final")
        self.opts = make_opts(_alg_name(alg), F, ts, dof=min(int(n_global), 2 ** 31 - 1), device=device, **kw)
        p = F.param_array()
        if p is None:
            p = np.zeros(1)
        self.h = C.c_void_p()
        self.L = lib()
        try:
            check(self.L.ode_b200_slab_create(C.byref(self.opts), int(n_local), int(n_global), int(rank), int(world),
                                              p.ctypes.data_as(_lib.dp), float(ts[0]), float(ts[-1]),
                                              C.c_void_p(stream_ptr), C.byref(self.h)))
        except _lib.OdeB200Error as e:
            _raise_like_reference(e)

    def reset(self, t0, tend):
        check(self.L.ode_b200_slab_reset(self.h, float(t0), float(tend)))

    def set_fused(self, on=True):
        """Single stencil slab: the attempt kernel's last thread block runs the step controller (no exchange needed
        between attempt() and control())."""
        check(self.L.ode_b200_slab_set_fused(self.h, 1 if on else 0))
        self.fused = bool(on)

    def set_state(self, dev_ptr):
        check(self.L.ode_b200_slab_set_state(self.h, C.c_void_p(dev_ptr)))

    def get_state(self, dev_ptr):
        check(self.L.ode_b200_slab_get_state(self.h, C.c_void_p(dev_ptr)))

    def set_output(self, tspan, out_dev_ptr):
        """points=:specified rows [len(tspan), n_local] into a device buffer; call before set_state."""
        ts = np.ascontiguousarray(np.asarray(tspan, dtype=np.float64))
        check(self.L.ode_b200_slab_set_output(self.h, ts.ctypes.data_as(_lib.dp), ts.size, C.c_void_p(out_dev_ptr)))

    def set_trace(self, dev_ptr, cap):
        check(self.L.ode_b200_slab_set_trace(self.h, C.c_void_p(dev_ptr), int(cap)))

    def init_phase(self, phase, send_ptr, recv_ptr):
        check(self.L.ode_b200_slab_init_phase(self.h, phase, C.c_void_p(send_ptr), C.c_void_p(recv_ptr)))

    def attempt(self, send_ptr):
        check(self.L.ode_b200_slab_attempt(self.h, C.c_void_p(send_ptr)))

    def control(self, recv_ptr):
        check(self.L.ode_b200_slab_control(self.h, C.c_void_p(recv_ptr)))

    @staticmethod
    def _ctl(a):
        return dict(done=bool(a[0]), status=int(a[1]), n_accept=int(a[2]), n_reject=int(a[3]), t=a[4], dt=a[5],
                    err=a[6], t_final=a[7], n_rhs=int(a[8]), attempts=int(a[9]), trace_n=int(a[10]), ghost=int(a[11]),
                    out_rows=int(a[12]))

    def poll(self):
        a = np.zeros(16)
        check(self.L.ode_b200_slab_poll(self.h, a.ctypes.data_as(_lib.dp)))
        return self._ctl(a)

    def peer_alloc(self):
        """This rank's receive block for the exchange over peer memory: returns the CUDA IPC handle (bytes) the
        other processes of the node map it with."""
        h = C.create_string_buffer(_lib.PEER_HANDLE_BYTES)
        check(self.L.ode_b200_slab_peer_alloc(self.h, h, None))
        return h.raw

    def peer_connect(self, handles):
        """handles: the IPC handles of all ranks in rank order (this rank's own entry is ignored)."""
        blob = b"".join(handles)
        check(self.L.ode_b200_slab_peer_connect(self.h, blob, None))
        self.fused = True
        self.peer = True

    def run(self):
        """The whole integration in one C call (hinit, graph-replayed attempt batches, device-side controller):
        for a single slab or slabs connected over peer memory.  Returns the control block like poll()."""
        a = np.zeros(16)
        ms = C.c_double(0.0)
        check(self.L.ode_b200_slab_run(self.h, a.ctypes.data_as(_lib.dp), C.byref(ms)))
        ctl = self._ctl(a)
        ctl["kernel_ms"] = ms.value
        return ctl

    def close(self):
        if self.h:
            self.L.ode_b200_slab_free(self.h)
            self.h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def integrate_slab(engine, send, recv, exchange, poll_every=4, max_polls=10 ** 9):
    """The host loop shared by every rank.

    engine   : slab engine (init_phase / attempt / control / poll)
    send     : this rank's exchange buffer  [XCHG_DOUBLES]   (object with .data_ptr(), or an int pointer)
    recv     : gathered buffer              [world, XCHG_DOUBLES]
    exchange : callable() performing recv <- all_gather(send) in stream order
    The done flag lives on the device; it is polled every `poll_every` attempts
    (attempts issued after the end are no-ops on the device).
    """
    sp = send.data_ptr() if hasattr(send, "data_ptr") else send
    rp = recv.data_ptr() if hasattr(recv, "data_ptr") else recv
    for phase in range(4):
        engine.init_phase(phase, sp, rp)
        if phase < 3:
            exchange()
    polls = 0
    fused = getattr(engine, "fused", False)  # single slab: the controller already ran inside attempt()
    while True:
        for _ in range(poll_every):
            engine.attempt(sp)
            if not fused:
                exchange()
            engine.control(rp)
        ctl = engine.poll()
        polls += 1
        if ctl["done"] or polls >= max_polls:
            return ctl


def connect_peers(eng, group=None):
    """Collective over `group`: map every rank's receive block into every other rank (CUDA IPC) so the slabs
    exchange their records from inside the attempt kernel.  Raises on every rank if any rank failed."""
    import torch.distributed as dist
    world = dist.get_world_size(group)
    err, handle = None, b""
    try:
        handle = eng.peer_alloc()
    except Exception as e:  # noqa: BLE001 -- reported collectively below
        err = repr(e)
    got = [None] * world
    dist.all_gather_object(got, (handle, err), group=group)
    if any(g[1] for g in got):
        raise RuntimeError("peer exchange unavailable: %s" % [g[1] for g in got if g[1]][0])
    try:
        eng.peer_connect([g[0] for g in got])
    except Exception as e:  # noqa: BLE001
        err = repr(e)
    got2 = [None] * world
    dist.all_gather_object(got2, err, group=group)  # also the barrier before the first run
    if any(got2):
        eng.peer = False
        raise RuntimeError("peer exchange unavailable: %s" % [g for g in got2 if g][0])


def solve_large_distributed(alg, F, y_local, n_global, tspan, *, group=None, poll_every=4, trace=False,
                            trace_cap=65536, engine=None, keep_engine=False, exchange="auto", **kw):
    """Domain-decomposed solve.  y_local: CUDA float64 torch tensor holding this rank's slab
    (shard_bounds(n_global, world, rank)); returns (y_local_final tensor, ctl dict[, trace tensor]).
    exchange: "peer" | "nccl" | "auto" (see the module docstring); ctl["exchange"] says which one ran."""
    import torch
    import torch.distributed as dist

    world = dist.get_world_size(group) if dist.is_available() and dist.is_initialized() else 1
    rank = dist.get_rank(group) if world > 1 else 0
    dev = y_local.device
    stream = torch.cuda.current_stream(dev)
    if engine is None:
        eng = CudaSlabEngine(alg, F, y_local.numel(), n_global, rank, world, tspan, stream.cuda_stream,
                             device=dev.index or 0, **kw)
        eng.peer = False
        if world > 1 and exchange in ("peer", "auto") and world <= _lib.MAX_PEERS:
            try:
                connect_peers(eng, group)
            except RuntimeError:
                if exchange == "peer":
                    raise
    else:  # reuse the device buffers of an earlier call (same slab size, options and stream)
        eng = engine
        eng.reset(tspan[0], tspan[-1])
    if world == 1 or getattr(eng, "peer", False):
        # one slab, or peers: one C call per rank runs the whole integration (CUDA-graph batches of attempts, the
        # controller on the device); peers exchange their records from inside the kernels
        tr = None
        if trace:
            tr = torch.zeros(trace_cap, 4, dtype=torch.float64, device=dev)
            eng.set_trace(tr.data_ptr(), trace_cap)
        y_local = y_local.contiguous()
        eng.set_state(y_local.data_ptr())
        ctl = eng.run()
        ctl["exchange"] = "peer" if world > 1 else "none"
        out = torch.empty_like(y_local)
        eng.get_state(out.data_ptr())
        torch.cuda.current_stream(dev).synchronize()
        if keep_engine:
            ctl["engine"] = eng
        else:
            if world > 1:
                dist.barrier(group)  # no rank unmaps its block while a peer may still look at it
            eng.close()
        if trace:
            return out, ctl, tr[:min(ctl["trace_n"], trace_cap)]
        return out, ctl
    send = torch.zeros(XCHG_DOUBLES, dtype=torch.float64, device=dev)
    recv = torch.zeros(world, XCHG_DOUBLES, dtype=torch.float64, device=dev)
    tr = None
    if trace:
        tr = torch.zeros(trace_cap, 4, dtype=torch.float64, device=dev)
        eng.set_trace(tr.data_ptr(), trace_cap)

    def do_exchange():
        dist.all_gather_into_tensor(recv.view(-1), send, group=group)

    y_local = y_local.contiguous()
    eng.set_state(y_local.data_ptr())
    ctl = integrate_slab(eng, send, recv, do_exchange, poll_every=poll_every)
    ctl["exchange"] = "nccl"
    out = torch.empty_like(y_local)
    eng.get_state(out.data_ptr())
    torch.cuda.current_stream(dev).synchronize()
    if keep_engine:
        ctl["engine"] = eng
    else:
        eng.close()
    if trace:
        return out, ctl, tr[:min(ctl["trace_n"], trace_cap)]
    return out, ctl
