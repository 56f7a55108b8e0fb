"""Registered right-hand-side tokens.

A Julia closure (or a Python callable) cannot run on the device, so the `F`
argument of the odeXX front-ends is a DeviceRHS: a callable object that
(i) evaluates the registered expression on the host -- handy for reference runs
and for checking the device functors -- and (ii) names the compiled device
functor (`rhs id`, parameter vector) that crosses the C ABI.
"""
import math

import numpy as np

from ._lib import RHS

_DOF = dict(linear=1, const=1, timelin=1, rotation=2, lorenz=3, vdp=2, robertson=3, kepler=4, reaction_diffusion=0,
            reaction_diffusion_field=0)
_NP = dict(linear=1, const=1, timelin=1, rotation=0, lorenz=3, vdp=1, robertson=3, kepler=0, reaction_diffusion=2,
           reaction_diffusion_field=2)


def register_rhs(name, dof, n_params, eval_body, jac_body=None, host=None, params=()):
    """Register a right-hand side from CUDA C++ source at run time and return its DeviceRHS token.

    eval_body : body of  void eval(double t, const double (&y)[dof], double (&f)[dof], const double (&p)[max(n_params,1)])
    jac_body  : optional body of  void jac(double t, const double (&y)[dof], double (&J)[dof][dof], const double (&p)[..])
    host      : optional python callable host(t, y, p) evaluating the same expression (for DeviceRHS.__call__)
    The library compiles its own kernel templates for this functor with NVRTC on first use (no FMA contraction)."""
    from ._lib import lib
    rid = lib().ode_b200_register_rhs(name.encode(), int(dof), int(n_params), eval_body.encode(),
                                      jac_body.encode() if jac_body else None)
    if rid < 0:
        from ._lib import check
        check(rid)
    tok = DeviceRHS.__new__(DeviceRHS)
    tok.name, tok.id, tok.dof, tok.n_params, tok.params = name, rid, int(dof), int(n_params), tuple(params)
    tok._host = host
    return tok


def register_stencil(name, point_body, host=None, params=(0.0, 0.0), radius=1):
    """Register a periodic 1-D stencil of radius 1..4 for the large-system regime from CUDA C++ source.

    point_body : body of  double point(const double (&w)[2*radius+1], const double (&p)[8], double t,
                 long long i, long long n)  -- w[k] = u[i-radius+k], t the stage time, i the global grid index,
                 n the ring length; must `return`.  A 3-point stencil may also use um, u, up and p0, p1.
    host       : optional python callable host(w, p, t, i, n) for DeviceRHS.__call__ on whole arrays
    params     : default parameter values (at most 8); their count is the stencil's n_params"""
    from ._lib import lib, check
    rid = lib().ode_b200_register_stencil(name.encode(), int(radius), len(params), point_body.encode())
    if rid < 0:
        check(rid)
    tok = DeviceRHS.__new__(DeviceRHS)
    tok.name, tok.id, tok.dof, tok.n_params, tok.params = name, rid, 0, len(params), tuple(params)
    tok._host = host
    tok.is_stencil = True
    tok.radius = int(radius)
    return tok


def register_field(name, rhs_body, host=None, params=()):
    """Register a GENERAL right-hand side for one large system (dof = N) from CUDA C++ source: F_i(t, y) may read
    any component of y (non-local couplings, networks, N-body sums, flattened 2-D grids).

    rhs_body : body of  double rhs(long long i, long long n, const double* y, const double (&p)[8], double t)
               returning dy_i/dt
    host     : optional python callable host(t, y, p) -> array, for DeviceRHS.__call__
    params   : default parameter values (at most 8)
    Runs one kernel per Runge-Kutta stage on one GPU (`solve_large`, or the odeXX front-ends on a dof = N state)."""
    from ._lib import lib, check
    rid = lib().ode_b200_register_field(name.encode(), len(params), rhs_body.encode())
    if rid < 0:
        check(rid)
    tok = DeviceRHS.__new__(DeviceRHS)
    tok.name, tok.id, tok.dof, tok.n_params, tok.params = name, rid, 0, len(params), tuple(params)
    tok._host = host
    tok.is_stencil = True  # routed to the large-system entry points
    tok.is_field = True
    return tok


class DeviceRHS:
    """`DeviceRHS("lorenz", 10.0, 28.0, 8/3)`; calling it evaluates F(t, y) on the host."""

    def __init__(self, name, *params):
        if name not in RHS:
            raise KeyError("unknown right-hand side %r; registered: %s" % (name, ", ".join(sorted(RHS))))
        if len(params) == 1 and np.ndim(params[0]) >= 1:
            params = tuple(np.asarray(params[0], dtype=np.float64).ravel().tolist()) if np.ndim(params[0]) == 1 \
                else (np.asarray(params[0], dtype=np.float64),)
        self.name = name
        self.id = RHS[name]
        self.dof = _DOF[name]
        self.n_params = _NP[name]
        self.params = params
        self.is_stencil = name in ("reaction_diffusion", "reaction_diffusion_field")  # a dof = N large-system RHS

    def with_params(self, p):
        if p is None:
            return self
        vals = np.atleast_1d(np.asarray(p, dtype=np.float64)).tolist()
        if self.id >= 1000:  # run-time registered: same functor, new parameter values
            tok = DeviceRHS.__new__(DeviceRHS)
            tok.__dict__.update(self.__dict__)
            tok.params = tuple(vals)
            return tok
        return DeviceRHS(self.name, *vals)

    def param_array(self):
        if self.n_params == 0:
            return None
        if len(self.params) == 1 and isinstance(self.params[0], np.ndarray):
            a = np.ascontiguousarray(self.params[0], dtype=np.float64)
        else:
            a = np.asarray(self.params, dtype=np.float64)
        if a.shape[-1] != self.n_params:
            raise ValueError("%s takes %d parameters, got %r" % (self.name, self.n_params, a.shape))
        return a

    def __call__(self, t, y):
        if self.id >= 3000:
            if getattr(self, "_host", None) is None:
                raise TypeError("run-time field %r was registered without a host evaluator" % self.name)
            return np.asarray(self._host(t, np.asarray(y, dtype=np.float64), list(self.params)), dtype=np.float64)
        if self.id >= 2000:
            if getattr(self, "_host", None) is None:
                raise TypeError("run-time stencil %r was registered without a host evaluator" % self.name)
            u = np.asarray(y, dtype=np.float64)
            n, R = u.size, self.radius
            return np.array([self._host([u[(i + k - R) % n] for k in range(2 * R + 1)], list(self.params), t, i, n)
                             for i in range(n)])
        if self.id >= 1000:
            if getattr(self, "_host", None) is None:
                raise TypeError("run-time RHS %r was registered without a host evaluator" % self.name)
            return self._host(t, y, self.params)
        p = self.params
        n = self.name
        scalar = np.ndim(y) == 0
        v = [float(y)] if scalar else [float(x) for x in y]
        if n == "linear":
            f = [p[0] * v[0]]
        elif n == "const":
            f = [p[0]]
        elif n == "timelin":
            f = [p[0] * t]
        elif n == "rotation":
            f = [-v[1], v[0]]
        elif n == "lorenz":
            f = [p[0] * (v[1] - v[0]), v[0] * (p[1] - v[2]) - v[1], v[0] * v[1] - p[2] * v[2]]
        elif n == "vdp":
            f = [v[1], p[0] * ((1.0 - v[0] * v[0]) * v[1]) - v[0]]
        elif n == "robertson":
            d1 = -p[0] * v[0] + p[2] * v[1] * v[2]
            d3 = p[1] * v[1] * v[1]
            f = [d1, -d1 - d3, d3]
        elif n == "kepler":
            r2 = v[0] * v[0] + v[1] * v[1]
            r3 = r2 * math.sqrt(r2)
            f = [v[2], v[3], -v[0] / r3, -v[1] / r3]
        else:
            u = np.asarray(v)
            f = p[0] * ((np.roll(u, 1) - 2.0 * u) + np.roll(u, -1)) + (p[1] * u) * (1.0 - u)
            return f
        return f[0] if scalar else np.array(f)

    def __repr__(self):
        return "DeviceRHS(%r%s)" % (self.name, "".join(", %r" % (x,) for x in self.params))
