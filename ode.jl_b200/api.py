"""The reference's `tout, yout = odeXX(F, y0, tspan; keywords...)` surface.

Host mirror of /root/reference src/runge_kutta.jl:165-168,218-224 (front-ends),
:232-240 (keywords and defaults) and src/ODE.jl:252-259 (ode23s), over the C ABI
of libode_b200.so.  `F` must be a DeviceRHS; anything else raises -- there is no
CPU fallback.  Error messages and the minstep warning are the reference's.
"""
import ctypes as C

import numpy as np

from . import _lib
from ._lib import Opts, EnsembleIO, METHODS, POINTS, check, lib
from .rhs import DeviceRHS

_FIXED = ("ode1", "ode2_midpoint", "ode2_heun", "ode4")
_ROSENBROCK_FIXED = ("ode4s", "ode4s_s", "ode4s_kr")  # oderosenbrock(...; jacobian=nothing, kwargs...), src/ODE.jl:366
_MULTISTEP = ("ode4ms", "ode5ms", "ode_ms1", "ode_ms2", "ode_ms3", "ode_ms4", "ode_ms5")  # ode_ms, src/ODE.jl:175-213


def _vp(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


def _points(points):
    if isinstance(points, str):
        points = points.lstrip(":")
    if points not in POINTS:
        raise ValueError("Unrecognized option points==%s" % (points,))  # src/runge_kutta.jl:289
    return POINTS[points]


def make_opts(method, F, tspan, *, reltol=1e-5, abstol=1e-8, minstep=None, maxstep=None, initstep=0.0,
              points="all", norm=None, jacobian=None, device=0, max_steps=0, params_stride=0, dof=None,
              direct_output=False, n_gpus=1, strict_fp=True):
    """ode_b200_opts with the reference's defaults (src/runge_kutta.jl:233-239, src/ODE.jl:253-259)."""
    if not isinstance(F, DeviceRHS):
        raise TypeError("F must be a registered DeviceRHS (a host closure cannot run on the device; "
                        "libode_b200 has no CPU fallback), got %r" % type(F).__name__)
    if norm is not None:
        raise NotImplementedError("only the default norm (LinearAlgebra.norm) is compiled into the kernels")
    # `jacobian = G(t,y)` (src/ODE.jl:254,262-267): a host closure cannot run on the device, so the
    # keyword selects the REGISTERED analytic Jacobian of F (jacobian=True); None/False keeps fdjacobian.
    if jacobian not in (None, False, True, 0, 1):
        raise TypeError("jacobian must be None/False (fdjacobian) or True (the registered analytic Jacobian of F)")
    if jacobian and method not in ("ode23s",) + _ROSENBROCK_FIXED:
        raise TypeError("%s does not take a jacobian keyword" % method)
    span = abs(float(tspan[-1]) - float(tspan[0]))
    o = Opts()
    o.method = METHODS[method]
    o.rhs = F.id
    o.dof = int(F.dof if dof is None else dof)
    o.n_params = F.n_params
    o.params_stride = int(params_stride)
    o.points = _points(points)
    o.jacobian = 1 if jacobian else 0
    o.device = int(device)
    o.max_steps = int(max_steps)
    o.reltol = float(reltol)
    o.abstol = float(abstol)
    o.minstep = span / 1e18 if minstep is None else float(minstep)  # src/runge_kutta.jl:235
    o.maxstep = span / 2.5 if maxstep is None else float(maxstep)   # :236
    o.initstep = float(initstep)
    o.mathmode = 0
    o.host_output = 1 if direct_output else 0  # kernel writes results into pinned `out=` buffers itself
    o.n_gpus = int(n_gpus)  # fan-out inside the library over devices device .. device + n_gpus - 1
    o.fp_mode = _lib.FP_STRICT if strict_fp else _lib.FP_FMA  # strict_fp=False: opt-in FMA contraction, not bit-exact
    return o


def _raise_like_reference(err):
    """Map C-ABI error codes to the exceptions/messages of the reference."""
    msg = {_lib.E_ZERO_SPAN: "Zero time span", _lib.E_INITSTEP: "initstep has wrong sign.",
           _lib.E_NOT_ADAPTIVE: "Can only use this solver with an adaptive RK Butcher table"}.get(err.code)
    if msg is not None:
        raise ValueError(msg) from err
    if err.code == _lib.E_POINTS:
        raise ValueError(str(err).split(": ", 1)[-1]) from err
    raise err


class Solver:
    """One odeXX front-end.  Calling it with arguments solves; calling it with none
    returns the algorithm marker used by `solve(prob, alg)` (src/algorithm_types.jl)."""

    def __init__(self, name):
        self.name = name

    def __repr__(self):
        return "%s()" % self.name

    def __call__(self, F=None, y0=None, tspan=None, **kw):
        if F is None and y0 is None and tspan is None and not kw:
            return self
        return _solve_single(self.name, F, y0, tspan, **kw)

    def stats(self, F, y0, tspan, **kw):
        """Like the call, but also returns dict(n_accept, n_reject, n_rhs, status)."""
        return _solve_single(self.name, F, y0, tspan, _stats=True, **kw)


def _solve_single(name, F, y0, tspan, _stats=False, **kw):
    if isinstance(F, DeviceRHS) and getattr(F, "is_stencil", False):
        # a dof = N method-of-lines state goes to the fused large-system kernels; same (tout, yout) contract
        from .large import solve_large
        kw.setdefault("points", "all")
        r = solve_large(name, F, y0, tspan, **kw)
        if r["status"] == 1:
            print("Warning: dt < minstep.  Stopping.")  # src/runge_kutta.jl:359
        st = dict(n_accept=r["n_accept"], n_reject=r["n_reject"], n_rhs=r["n_rhs"], status=r["status"])
        return (r["t"], r["y"], st) if _stats else (r["t"], r["y"])
    scalar = np.ndim(y0) == 0
    y = np.ascontiguousarray(np.atleast_1d(np.asarray(y0, dtype=np.float64)))
    ts = np.ascontiguousarray(np.asarray(tspan, dtype=np.float64))
    if name in _FIXED:
        # oderk_fixed(fn, y0::AbstractVector, tspan, btab) takes no keywords (src/runge_kutta.jl:176-177);
        # the scalar method swallows them (:170)
        if kw and not scalar:
            raise TypeError("%s on a vector y0 takes no keyword arguments (reference: oderk_fixed)" % name)
        kw = {}
    elif name in _ROSENBROCK_FIXED:
        kw = {"jacobian": kw.get("jacobian")}  # the other keywords are swallowed by kwargs... (src/ODE.jl:366)
        if name == "ode4s":
            kw = {}  # ode4s(F, x0, tspan; jacobian=nothing, kwargs...) passes jacobian=nothing on (src/ODE.jl:437-438)
    elif name in _MULTISTEP:
        kw = {}  # ode_ms(F, x0, tspan, order; kwargs...) ignores keywords (src/ODE.jl:175,212-213)
    o = make_opts(name, F, ts, **kw)
    if y.size != o.dof:
        raise ValueError("%r has %d components, y0 has %d" % (F, o.dof, y.size))
    p = F.param_array()
    L = lib()
    h = C.c_void_p()
    try:
        check(L.ode_b200_solve_single(C.byref(o), y.ctypes.data_as(_lib.dp),
                                      None if p is None else p.ctypes.data_as(_lib.dp),
                                      ts.ctypes.data_as(_lib.dp), ts.size, C.byref(h)))
    except _lib.OdeB200Error as e:
        _raise_like_reference(e)
    try:
        n = L.ode_b200_result_len(h)
        tout = np.empty(n)
        yout = np.empty((n, o.dof))
        check(L.ode_b200_result_copy(h, tout.ctypes.data_as(_lib.dp), yout.ctypes.data_as(_lib.dp)))
        status = L.ode_b200_result_status(h)
        st = dict(n_accept=L.ode_b200_result_counts(h, 0), n_reject=L.ode_b200_result_counts(h, 1),
                  n_rhs=L.ode_b200_result_counts(h, 2), status=status)
    finally:
        L.ode_b200_result_free(h)
    if status == 1 and name != "ode23s":
        print("Warning: dt < minstep.  Stopping.")  # src/runge_kutta.jl:359
    if scalar:
        yout = yout[:, 0]  # vcat_nosplat, src/runge_kutta.jl:477
    return (tout, yout, st) if _stats else (tout, yout)


ode1 = Solver("ode1")
ode2_midpoint = Solver("ode2_midpoint")
ode2_heun = Solver("ode2_heun")
ode4 = Solver("ode4")
ode21 = Solver("ode21")
ode23 = Solver("ode23")
ode45_fe = Solver("ode45_fe")
ode45_dp = Solver("ode45_dp")
ode45 = Solver("ode45")  # Dormand-Prince by default, src/runge_kutta.jl:223
ode78 = Solver("ode78")
# EXTENSION: Cash-Karp 4(5).  README.md:57 of the reference says the coefficients "are also available"; its source has
# no such tableau, so this one carries the published coefficients (Cash & Karp 1990) in bt_dopri5's layout, order (5,4).
ode45_ck = Solver("ode45_ck")
ode23s = Solver("ode23s")
ode4s_s = Solver("ode4s_s")    # Shampine coefficients, src/ODE.jl:433
ode4s_kr = Solver("ode4s_kr")  # Kaps-Rentrop coefficients, src/ODE.jl:419
ode4s = Solver("ode4s")        # = ode4s_s with jacobian=nothing, src/ODE.jl:437
ode4ms = Solver("ode4ms")      # Adams-Bashforth order 4, src/ODE.jl:212
ode5ms = Solver("ode5ms")      # order 5, src/ODE.jl:213 (not exported by the reference either: ODE.ode5ms)


ode_ms1, ode_ms2, ode_ms3 = Solver("ode_ms1"), Solver("ode_ms2"), Solver("ode_ms3")  # ODE.ode_ms(F, x0, tspan, 1..3)


def ode_ms(F, x0, tspan, order, **kw):
    """ode_ms(F, x0, tspan, order; kwargs...) (src/ODE.jl:175): fixed-step Adams-Bashforth of order 1..5 on the tspan
    grid, the first steps at reduced order."""
    if int(order) != order or not 1 <= int(order) <= 5:
        raise ValueError("ode_ms is built for order 1..5 (the reference computes higher orders with Polynomials)")
    return _solve_single("ode_ms%d" % int(order), F, x0, tspan, **kw)



def pinned_empty(shape, dtype=np.float64):
    """An uninitialised numpy array in page-locked, device-mapped host memory (ode_b200_host_alloc).
    Results the library delivers into such arrays are written by the kernel as trajectories finish, so
    there is no copy-back phase after the integration; use them for buffers that are reused across calls
    (`solve_ensemble(..., out=...)`)."""
    import weakref
    dt = np.dtype(dtype)
    n = int(np.prod(shape)) * dt.itemsize
    L = lib()
    p = L.ode_b200_host_alloc(n)
    if not p:
        raise MemoryError(L.ode_b200_last_error().decode())
    buf = (C.c_char * max(n, 1)).from_address(p)
    a = np.frombuffer(buf, dtype=dt, count=int(np.prod(shape))).reshape(shape)
    weakref.finalize(buf, L.ode_b200_host_free, p)  # the array keeps `buf` alive through .base
    return a


def _alg_name(alg):
    return alg.name if isinstance(alg, Solver) else str(alg)


def solve_ensemble(alg, F, y0, tspan, *, params=None, points="final", pts_cap=None, trace_traj=None,
                   trace_cap=65536, out=None, **kw):
    """n independent odeXX calls in one launch.

    y0: [n_traj, dof].  params: None (use F's bound parameters for every
    trajectory) or [n_traj, n_params].  Returns a dict with t_final, y_final,
    n_accept, n_reject, n_rhs, status (+ pts_t, pts_y, pts_n for points != "final",
    + trace rows (t, dt, err, accepted) for trace_traj).  `out`: a dict returned by an earlier call of
    the same shape (or arrays from `pinned_empty`) whose t_final ... status arrays are filled in place;
    with `direct_output=True` the kernel writes into the pinned ones itself as trajectories finish (no copy-back
    phase; pays with one GPU per host, measured no gain with 8 GPUs sharing one host).
    """
    name = _alg_name(alg)
    y0 = np.ascontiguousarray(np.asarray(y0, dtype=np.float64))
    if y0.ndim != 2:
        raise ValueError("y0 must be [n_traj, dof]")
    n, dof = y0.shape
    ts = np.ascontiguousarray(np.asarray(tspan, dtype=np.float64))
    if params is None:
        p = F.param_array()
        stride = 0
    else:
        p = np.ascontiguousarray(np.asarray(params, dtype=np.float64))
        if p.shape != (n, F.n_params):
            raise ValueError("params must be [n_traj, %d]" % F.n_params)
        stride = F.n_params
    o = make_opts(name, F, ts, points=points, params_stride=stride, **kw)
    if dof != o.dof:
        raise ValueError("%r has %d components, y0 has %d" % (F, o.dof, dof))
    want = dict(t_final=((n,), np.float64), y_final=((n, dof), np.float64), n_accept=((n,), np.int32),
                n_reject=((n,), np.int32), n_rhs=((n,), np.int32), status=((n,), np.int32))
    given = out if out is not None else {}
    out = {}
    for k, (shape, dt) in want.items():
        a = given.get(k)
        if a is None:
            a = np.empty(shape, dt)
        elif a.shape != shape or a.dtype != dt or not a.flags.c_contiguous:
            raise ValueError("out[%r] must be a C-contiguous %s array of shape %r" % (k, np.dtype(dt).name, shape))
        out[k] = a
    io = EnsembleIO()
    io.n_traj = n
    io.y0 = _vp(y0)
    io.params = _vp(p)
    io.tspan = _vp(ts)
    io.n_tspan = ts.size
    for k in ("t_final", "y_final", "n_accept", "n_reject", "n_rhs", "status"):
        setattr(io, k, _vp(out[k]))
    io.trace_traj = -1
    if o.points != POINTS["final"]:
        cap = int(pts_cap if pts_cap is not None else ts.size)
        out["pts_t"] = np.zeros((n, cap))
        out["pts_y"] = np.zeros((n, cap, dof))
        out["pts_n"] = np.zeros(n, np.int32)
        io.pts_t, io.pts_y, io.pts_n, io.pts_cap = _vp(out["pts_t"]), _vp(out["pts_y"]), _vp(out["pts_n"]), cap
    tn = np.zeros(1, np.int64)
    if trace_traj is not None:
        out["trace"] = np.zeros((trace_cap, 4))
        io.trace_traj, io.trace, io.trace_cap, io.trace_n = int(trace_traj), _vp(out["trace"]), trace_cap, _vp(tn)
    try:
        check(lib().ode_b200_solve_ensemble(C.byref(o), C.byref(io)))
    except _lib.OdeB200Error as e:
        _raise_like_reference(e)
    if trace_traj is not None:
        out["trace"] = out["trace"][:min(int(tn[0]), trace_cap)]
        out["trace_n"] = int(tn[0])
    out["kernel_ms"] = lib().ode_b200_last_kernel_ms()
    return out
