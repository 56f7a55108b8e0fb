"""ctypes front for oracle/libode_oracle.so (the C++ restatement of ODE.jl).

TEST INFRASTRUCTURE ONLY -- see the header of ode_oracle.cpp.  Importable from
tests/, __graft_entry__.smoke() and bench.py's CPU-baseline legs; never from
the product package.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "libode_oracle.so")

# ids, as include/ode_b200.h
METHODS = dict(ode1=0, feuler=0, ode2_midpoint=1, midpoint=1, ode2_heun=2, heun=2, ode4=3, rk4=3,
               ode21=4, rk21=4, ode23=5, rk23=5, ode45_fe=6, rk45=6, ode45=7, ode45_dp=7, dopri5=7,
               ode78=8, feh78=8, ode23s=9, ode4s=10, ode4s_s=10, ode4s_kr=11, ode4ms=12, ode5ms=13, ode_ms1=14,
               ode_ms2=15, ode_ms3=16, ode_ms4=12, ode_ms5=13, ode45_ck=17, ck45=17)
RHS = dict(linear=0, const=1, timelin=2, rotation=3, lorenz=4, vdp=5, robertson=6, kepler=7,
           reaction_diffusion=8)
POINTS = dict(all=0, specified=1, final=2)


class Opts(C.Structure):
    _fields_ = [("method", C.c_int32), ("rhs", C.c_int32), ("dof", C.c_int32), ("n_params", C.c_int32),
                ("params_stride", C.c_int32), ("points", C.c_int32), ("jacobian", C.c_int32),
                ("device", C.c_int32), ("max_steps", C.c_int64),
                ("reltol", C.c_double), ("abstol", C.c_double), ("minstep", C.c_double),
                ("maxstep", C.c_double), ("initstep", C.c_double),
                ("mathmode", C.c_int32), ("host_output", C.c_int32), ("n_gpus", C.c_int32), ("fp_mode", C.c_int32),
                ("reserved", C.c_int32 * 4)]


def build(force=False):
    """Compile the oracle with the recipe in oracle/Makefile."""
    if force or not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(
            os.path.join(_HERE, "ode_oracle.cpp")):
        subprocess.check_call(["make", "-C", _HERE, "-s"])
    return _SO


_lib = None


def lib():
    global _lib
    if _lib is None:
        build()
        L = C.CDLL(_SO)
        dp = C.POINTER(C.c_double)
        ip = C.POINTER(C.c_int32)
        L.oracle_solve.argtypes = [C.POINTER(Opts), dp, dp, dp, C.c_int, C.c_int, C.POINTER(C.c_void_p)]
        L.oracle_solve.restype = C.c_int
        L.oracle_result_len.argtypes = [C.c_void_p]
        L.oracle_result_len.restype = C.c_int64
        L.oracle_result_trace_len.argtypes = [C.c_void_p]
        L.oracle_result_trace_len.restype = C.c_int64
        L.oracle_result_status.argtypes = [C.c_void_p]
        L.oracle_result_count.argtypes = [C.c_void_p, C.c_int]
        L.oracle_result_count.restype = C.c_int64
        L.oracle_result_copy.argtypes = [C.c_void_p, dp, dp, dp]
        L.oracle_result_copy.restype = None
        L.oracle_result_free.argtypes = [C.c_void_p]
        L.oracle_result_free.restype = None
        L.oracle_solve_ensemble.argtypes = [C.POINTER(Opts), C.c_int64, dp, dp, dp, C.c_int, dp, dp, ip, ip, ip, ip,
                                            C.c_int]
        L.oracle_solve_ensemble.restype = C.c_int
        L.oracle_tableau.argtypes = [C.c_int] * 4
        L.oracle_tableau.restype = C.c_double
        L.oracle_tableau_info.argtypes = [C.c_int, C.c_int]
        L.oracle_pow.argtypes = [C.c_double, C.c_double, C.c_int]
        L.oracle_pow.restype = C.c_double
        L.oracle_log10.argtypes = [C.c_double, C.c_int]
        L.oracle_log10.restype = C.c_double
        L.oracle_pow_v.argtypes = [dp, dp, dp, C.c_int64, C.c_int]
        L.oracle_pow_v.restype = None
        L.oracle_log10_v.argtypes = [dp, dp, C.c_int64, C.c_int]
        L.oracle_log10_v.restype = None
        L.oracle_rhs.argtypes = [C.c_int, C.c_int64, dp, C.c_double, dp, dp]
        L.oracle_rhs.restype = None
        L.oracle_norm2.argtypes = [dp, C.c_int64]
        L.oracle_norm2.restype = C.c_double
        L.oracle_max_threads.restype = C.c_int
        L.oracle_ms_coefficient.argtypes = [C.c_int, C.c_int]
        L.oracle_ms_coefficient.restype = C.c_double
        L.oracle_set_user_rhs.argtypes = [C.c_void_p, C.c_void_p]
        L.oracle_set_user_rhs.restype = None
        L.oracle_set_user_stencil.argtypes = [C.c_void_p, C.c_int]
        L.oracle_set_user_stencil.restype = None
        L.oracle_set_threads.argtypes = [C.c_int]
        L.oracle_set_threads.restype = None
        _lib = L
    return _lib


def _dp(a):
    return a.ctypes.data_as(C.POINTER(C.c_double)) if a is not None else None


def _ip(a):
    return a.ctypes.data_as(C.POINTER(C.c_int32)) if a is not None else None


def make_opts(method, rhs, dof, params, tspan, *, reltol=1e-5, abstol=1e-8, minstep=None, maxstep=None, initstep=0.0,
              points="all", mathmode=0, max_steps=0, params_stride=0, jacobian=0):
    """Options with the reference's defaults (src/runge_kutta.jl:233-239)."""
    span = abs(float(tspan[-1]) - float(tspan[0]))
    o = Opts()
    o.method = METHODS[method] if isinstance(method, str) else int(method)
    o.rhs = RHS[rhs] if isinstance(rhs, str) else int(rhs)
    o.dof = int(dof)
    o.n_params = 0 if params is None else int(np.asarray(params).shape[-1])
    o.params_stride = int(params_stride)
    o.points = POINTS[points] if isinstance(points, str) else int(points)
    o.jacobian = int(jacobian)
    o.device = 0
    o.max_steps = int(max_steps)
    o.reltol = float(reltol)
    o.abstol = float(abstol)
    o.minstep = span / 1e18 if minstep is None else float(minstep)
    o.maxstep = span / 2.5 if maxstep is None else float(maxstep)
    o.initstep = float(initstep)
    o.mathmode = int(mathmode)
    return o


class Solution:
    def __init__(self, t, y, trace, nacc, nrej, nrhs, status):
        self.t, self.y, self.trace = t, y, trace
        self.n_accept, self.n_reject, self.n_rhs, self.status = nacc, nrej, nrhs, status


def solve(method, rhs, y0, tspan, params=None, *, trace=False, **kw):
    """One trajectory: returns Solution(tout, yout[len, dof], trace[n, 4], ...)."""
    L = lib()
    y0 = np.ascontiguousarray(np.atleast_1d(np.asarray(y0, dtype=np.float64)))
    tspan = np.ascontiguousarray(np.asarray(tspan, dtype=np.float64))
    p = np.ascontiguousarray(np.asarray(params if params is not None else [0.0], dtype=np.float64))
    o = make_opts(method, rhs, y0.size, p if params is not None else None, tspan, **kw)
    h = C.c_void_p()
    rc = L.oracle_solve(C.byref(o), _dp(y0), _dp(p), _dp(tspan), tspan.size, int(trace), C.byref(h))
    try:
        if rc != 0:
            raise RuntimeError("oracle error %d" % rc)
        n = L.oracle_result_len(h)
        nt = L.oracle_result_trace_len(h)
        t = np.empty(n)
        y = np.empty((n, y0.size))
        tr = np.empty((nt, 4))
        L.oracle_result_copy(h, _dp(t), _dp(y), _dp(tr) if nt else None)
        return Solution(t, y, tr, L.oracle_result_count(h, 0), L.oracle_result_count(h, 1),
                        L.oracle_result_count(h, 2), L.oracle_result_status(h))
    finally:
        L.oracle_result_free(h)


def solve_ensemble(method, rhs, y0, tspan, params=None, *, threads=0, **kw):
    """Final state + counters of every trajectory (OpenMP).  y0: [n, dof]."""
    L = lib()
    y0 = np.ascontiguousarray(np.asarray(y0, dtype=np.float64))
    n, dof = y0.shape
    tspan = np.ascontiguousarray(np.asarray(tspan, dtype=np.float64))
    if params is None:
        p = np.zeros(1)
        stride = 0
        po = None
    else:
        p = np.ascontiguousarray(np.asarray(params, dtype=np.float64))
        stride = p.shape[-1] if p.ndim == 2 else 0
        po = p
    kw.pop("points", None)
    o = make_opts(method, rhs, dof, po, tspan, points="final", params_stride=stride, **kw)
    tf = np.empty(n)
    yf = np.empty((n, dof))
    na = np.empty(n, np.int32)
    nr = np.empty(n, np.int32)
    nf = np.empty(n, np.int32)
    st = np.empty(n, np.int32)
    if threads <= 0:
        threads = L.oracle_max_threads()
    rc = L.oracle_solve_ensemble(C.byref(o), n, _dp(y0), _dp(p), _dp(tspan), tspan.size, _dp(tf), _dp(yf), _ip(na),
                                 _ip(nr), _ip(nf), _ip(st), threads)
    if rc != 0:
        raise RuntimeError("oracle error %d" % rc)
    return dict(t_final=tf, y_final=yf, n_accept=na, n_reject=nr, n_rhs=nf, status=st, threads=threads)


def tableau(method):
    L = lib()
    m = METHODS[method] if isinstance(method, str) else method
    S = L.oracle_tableau_info(m, 0)
    adaptive = bool(L.oracle_tableau_info(m, 1))
    nb = 2 if adaptive else 1
    a = np.array([[L.oracle_tableau(m, 0, i, j) for j in range(S)] for i in range(S)])
    b = np.array([[L.oracle_tableau(m, 1, i, j) for j in range(S)] for i in range(nb)])
    c = np.array([L.oracle_tableau(m, 2, i, 0) for i in range(S)])
    return dict(S=S, adaptive=adaptive, fsal=bool(L.oracle_tableau_info(m, 2)), order=L.oracle_tableau_info(m, 3),
                a=a, b=b, c=c)


def rhs_eval(rhs, t, y, params):
    L = lib()
    y = np.ascontiguousarray(np.asarray(y, dtype=np.float64))
    p = np.ascontiguousarray(np.asarray(params if params is not None else [0.0], dtype=np.float64))
    f = np.empty_like(y)
    L.oracle_rhs(RHS[rhs] if isinstance(rhs, str) else rhs, y.size, _dp(p), float(t), _dp(y), _dp(f))
    return f


def host_cores():
    """Cores this process may run on (torchrun exports OMP_NUM_THREADS=1, which would hide them)."""
    try:
        return len(os.sched_getaffinity(0))
    except Exception:
        return os.cpu_count() or 1


def use_all_cores():
    n = host_cores()
    lib().oracle_set_threads(n)
    return n


USER_RHS_BASE = 1000
_RHSFN = C.CFUNCTYPE(None, C.c_double, C.POINTER(C.c_double), C.POINTER(C.c_double), C.POINTER(C.c_double), C.c_int)


def solve_user(method, f, y0, tspan, params=None, jac=None, **kw):
    """One trajectory with a python right-hand side f(t, y, p) -> list (evaluated with python floats, i.e.
    binary64 without FMA), run through the C++ restatement.  Used to check run-time registered CUDA functors."""
    L = lib()
    y0a = np.ascontiguousarray(np.atleast_1d(np.asarray(y0, dtype=np.float64)))
    dof = y0a.size
    npar = 0 if params is None else len(params)

    def cb(t, yp, fp, pp, n):
        out = f(t, [yp[i] for i in range(n)], [pp[i] for i in range(npar)])
        for i in range(n):
            fp[i] = out[i]

    def cbj(t, yp, Jp, pp, n):
        out = jac(t, [yp[i] for i in range(n)], [pp[i] for i in range(npar)])
        for i in range(n):
            for j in range(n):
                Jp[i * n + j] = out[i][j]

    c1 = _RHSFN(cb)
    c2 = _RHSFN(cbj) if jac is not None else None
    L.oracle_set_user_rhs(C.cast(c1, C.c_void_p), C.cast(c2, C.c_void_p) if c2 else None)
    try:
        kw2 = dict(kw)
        if jac is not None:
            kw2["jacobian"] = 1
        return solve(method, USER_RHS_BASE, y0a, tspan, params, **kw2)
    finally:
        L.oracle_set_user_rhs(None, None)


STENCIL_BASE = 2000
_POINTFN = C.CFUNCTYPE(C.c_double, C.POINTER(C.c_double), C.POINTER(C.c_double), C.c_double, C.c_int64, C.c_int64)


def solve_user_stencil(method, point, y0, tspan, params, radius=1, **kw):
    """Large system with a python stencil point(w, p, t, i, n) -> float on a periodic grid: w[k] = u[i-radius+k],
    p the parameter list, t the stage time, i the grid index, n the ring length."""
    L = lib()
    npar = len(params)
    width = 2 * radius + 1

    def trampoline(w, p, t, i, n):
        return float(point([w[k] for k in range(width)], [p[q] for q in range(npar)], t, i, n))

    cb = _POINTFN(trampoline)
    L.oracle_set_user_stencil(C.cast(cb, C.c_void_p), int(radius))
    try:
        return solve(method, STENCIL_BASE, y0, tspan, params, **kw)
    finally:
        L.oracle_set_user_stencil(None, 1)
