#!/usr/bin/env python3
"""Pin the CPU oracle's bits: runs a fixed set of problems through oracle/ (C++ restatement, deterministic pow/log10
mode) and writes step counts and final states, as exact hex floats, to tests/golden/oracle_runs.json.

The fixture guards the checker itself: tests/test_oracle.py::test_oracle_against_golden_runs fails when a change to
oracle/ or to include/ode_b200_detmath.h moves any bit of these runs, so such a change has to be deliberate
(re-run this script and commit the new file together with the reason).
Run:  python tools/gen_golden_oracle.py
"""
import json
import math
import os
import sys

import numpy as np

ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
sys.path.insert(0, ROOT)
from oracle import oracle as O  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden", "oracle_runs.json")
LORENZ = [10.0, 28.0, 8.0 / 3.0]
ADAPTIVE = ["ode21", "ode23", "ode45_fe", "ode45", "ode78"]
FIXED = ["ode1", "ode2_midpoint", "ode2_heun", "ode4", "ode4s_s", "ode4s_kr", "ode4ms"]


def cases():
    yield dict(name="readme_linear", method="ode45", rhs="linear", y0=[0.5], tspan=[0.0, 1.0], params=[1.01],
               kw=dict(reltol=1e-8, abstol=1e-8, maxstep=0.4, minstep=1e9))  # README.md:19-23 through solve()
    for m in ADAPTIVE + ["ode23s"]:
        yield dict(name="lorenz_" + m, method=m, rhs="lorenz", y0=[1.0, 1.0, 1.0], tspan=[0.0, 2.0], params=LORENZ, kw={})
        yield dict(name="linear_backward_" + m, method=m, rhs="linear", y0=[1.0], tspan=[1.0, 0.0], params=[1.0], kw={})
    for m in FIXED:
        yield dict(name="lorenz_" + m, method=m, rhs="lorenz", y0=[1.0, 1.0, 1.0],
                   tspan=[float(x) for x in np.linspace(0.0, 0.5, 26)], params=LORENZ, kw={})
    yield dict(name="robertson_ode23s", method="ode23s", rhs="robertson", y0=[1.0, 0.0, 0.0], tspan=[0.0, 40.0],
               params=[0.04, 3e7, 1e4], kw=dict(reltol=1e-8, abstol=1e-8))
    yield dict(name="robertson_ode23s_analytic_jacobian", method="ode23s", rhs="robertson", y0=[1.0, 0.0, 0.0],
               tspan=[0.0, 40.0], params=[0.04, 3e7, 1e4], kw=dict(jacobian=1))
    e = 0.5
    yield dict(name="kepler_ode78", method="ode78", rhs="kepler", y0=[1 - e, 0.0, 0.0, math.sqrt((1 + e) / (1 - e))],
               tspan=[0.0, 6.0], params=None, kw=dict(reltol=1e-12, abstol=1e-12))
    yield dict(name="vdp_ode45_fe", method="ode45_fe", rhs="vdp", y0=[2.0, 0.0], tspan=[0.0, 3.0], params=[5.0], kw={})
    n = 64
    u0 = [float(x) for x in 0.5 + 0.4 * np.sin(2 * np.pi * 2 * np.arange(n) / n)]
    for m in ("ode45", "ode23", "ode4"):
        ts = [0.0, 2.0] if m != "ode4" else [float(x) for x in np.linspace(0.0, 0.5, 11)]
        yield dict(name="reaction_diffusion_64_" + m, method=m, rhs="reaction_diffusion", y0=u0, tspan=ts,
                   params=[1.0, 1.0], kw={})


def run(c):
    s = O.solve(c["method"], c["rhs"], c["y0"], c["tspan"], c["params"], points="all", mathmode=0, **c["kw"])
    return dict(n_accept=int(s.n_accept), n_reject=int(s.n_reject), n_rhs=int(s.n_rhs), status=int(s.status),
                n_out=int(len(s.t)), t_final=float(s.t[-1]).hex(), y_final=[float(v).hex() for v in np.atleast_1d(s.y[-1])])


if __name__ == "__main__":
    out = []
    for c in cases():
        out.append(dict(case=c, result=run(c)))
    with open(OUT, "w") as f:
        json.dump(out, f, indent=1)
    print("wrote %d runs to %s" % (len(out), OUT))
