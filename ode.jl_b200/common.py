"""The common-interface entry `solve(prob, alg; ...)`.

Host mirror of /root/reference src/common.jl:1-123 and src/algorithm_types.jl:
keyword mapping (reltol, abstol, saveat, save_everystep, save_start, dtmin,
dtmax, dt), construction of `Ts`, the `points` choice, and the call
`AlgType(f, u0, Ts; norm, abstol, reltol, maxstep=dtmax, minstep=dtmin,
initstep=dt, points)` (:93-100).  Quirks are kept: `dtmin` defaults to
span/1e-9 (:13) and the return code is spelt `:Succss` (:122).
"""
import warnings

import numpy as np

from .api import Solver, _alg_name
from .rhs import DeviceRHS

# src/ODE.jl:14-17
WARNKEYWORDS = ("save_idxs", "d_discontinuities", "isoutofdomain", "unstable_check", "calck", "progress",
                "internalnorm", "gamma", "beta1", "beta2", "qmax", "qmin", "qoldinit")
ALGORITHMS = ("ode23", "ode45", "ode23s", "ode78", "ode4", "ode4ms", "ode4s")  # src/algorithm_types.jl


class ODEProblem:
    """ODEProblem(f, u0, tspan, p=None): f is a DeviceRHS; p (if given) replaces its parameters."""

    def __init__(self, f, u0, tspan, p=None, mass_matrix=None, callback=None):
        self.f, self.u0, self.tspan, self.p = f, u0, (float(tspan[0]), float(tspan[1])), p
        self.mass_matrix, self.callback = mass_matrix, callback


class ODESolution:
    """What DiffEqBase.build_solution(prob, alg, ts, timeseries; retcode=:Succss) carries."""

    def __init__(self, prob, alg, t, u, retcode="Succss"):
        self.prob, self.alg, self.t, self.u, self.retcode = prob, alg, t, u, retcode

    def __getitem__(self, i):
        return self.u[i]

    def __len__(self):
        return len(self.u)


def _frange(a, step, b):
    """collect(a:step:b) for floats: a + k*step, k = 0..n, with Julia's habit of landing exactly on b when the
    range is meant to hit it (Julia computes float ranges in twice precision; this is the plain-float analogue,
    so an interior element may differ from Julia's by an ulp)."""
    if step == 0 or (b - a) / step < -1e-12:
        return []
    n = int(np.floor((b - a) / step * (1 + 1e-14) + 1e-12))
    out = [a + k * step for k in range(n + 1)]
    if out and abs(out[-1] - b) <= 4e-16 * max(abs(a), abs(b), abs(step)) * max(1, n):
        out[-1] = b
    return out


def build_ts(tspan, saveat):
    """`Ts` of src/common.jl:50-70: the time grid handed to the solver for a given `saveat`.

    saveat a number: tspan[1]+saveat : saveat : tspan[end] (or up to tspan[end]-saveat when the range does not
    land on tspan[end]); a collection: as given.  A trailing tspan[2] is dropped and re-appended, a leading
    tspan[1] is kept once; duplicates are removed in order (`unique`)."""
    t0, t1 = float(tspan[0]), float(tspan[1])
    if np.ndim(saveat) == 0:  # :52-57
        step = float(saveat)
        full = _frange(t0, step, t1)
        if full and full[-1] == t1:  # (tspan[1]:saveat:tspan[end])[end] == tspan[end]
            saveat_vec = _frange(t0 + step, step, t1)
        else:
            saveat_vec = _frange(t0 + step, step, t1 - step)
    else:
        saveat_vec = [float(x) for x in saveat]  # :59
    if saveat_vec and saveat_vec[-1] == t1:  # :62-64
        saveat_vec.pop()
    if saveat_vec and saveat_vec[0] == t0:   # :66-70
        Ts = saveat_vec + [t1]
    else:
        Ts = [t0] + saveat_vec + [t1]
    return list(dict.fromkeys(Ts))  # unique, order kept


def solve(prob, alg, *, verbose=True, save_timeseries=None, saveat=(), reltol=1e-5, abstol=1e-8,
          save_everystep=None, dense=None, save_start=None, callback=None, dtmin=None, dtmax=None,
          timeseries_errors=True, dense_errors=False, dt=0.0, norm=None, **kwargs):
    name = _alg_name(alg)
    if name not in ALGORITHMS:
        raise TypeError("%r is not an ODEjlAlgorithm" % (alg,))
    if not isinstance(prob.f, DeviceRHS):
        raise TypeError("prob.f must be a DeviceRHS (no CPU fallback)")
    tspan = prob.tspan
    span = abs(tspan[1] - tspan[0])
    saveat_is_number = np.ndim(saveat) == 0
    saveat_empty = (not saveat_is_number) and len(saveat) == 0
    if save_everystep is None:
        save_everystep = saveat_empty  # :8
    if save_start is None:             # :10-11
        save_start = True if (save_everystep or saveat_empty or saveat_is_number) else (tspan[0] in list(saveat))
    if dtmin is None:
        dtmin = span / 1e-9  # :13 (sic)
    if dtmax is None:
        dtmax = span / 2.5   # :14
    if verbose and kwargs:   # :21-35
        bad = [k for k in kwargs if k in WARNKEYWORDS]
        if bad:
            warnings.warn("keyword(s) %s are not supported by ODE.jl solvers and are ignored" % ", ".join(bad))
    if save_timeseries is not None:  # :37-40
        if verbose:
            warnings.warn("save_timeseries is deprecated. Use save_everystep instead")
        save_everystep = save_timeseries
    if prob.mass_matrix is not None:  # :42-44
        raise ValueError("This solver is not able to use mass matrices.")
    if callback is not None or prob.callback is not None:  # :46-48
        raise ValueError("ODE is not compatible with callbacks.")

    Ts = build_ts(tspan, saveat)
    points = "all" if save_everystep else "specified"  # :72-76

    u0 = np.asarray(prob.u0, dtype=np.float64)
    is_array = u0.ndim > 0
    sizeu = u0.shape
    f = prob.f.with_params(prob.p)
    solver = alg if isinstance(alg, Solver) else Solver(name)
    kw = dict(norm=norm, abstol=abstol, reltol=reltol, maxstep=dtmax, minstep=dtmin, initstep=dt, points=points)
    ts, ys = solver(f, u0.ravel() if is_array else float(u0), Ts, **kw)  # :93-100
    start = 0
    if not save_start:  # :102-107
        start = 1
        ts = ts[1:]
    if is_array:        # :110-117
        timeseries = [np.reshape(ys[i], sizeu) for i in range(start, len(ys))]
    else:
        timeseries = ys  # (not sliced in the reference either, :116)
    return ODESolution(prob, alg, ts, timeseries, retcode="Succss")
