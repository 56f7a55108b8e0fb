"""Randomised parity sweep: many (method, RHS, tolerances, span, direction, step limits, output mode)
combinations, each a small ensemble or a single system, CUDA result vs CPU oracle bit for bit.
Seeded, so a failure reproduces; the case description is part of the assertion message."""
import math

import numpy as np
import pytest

import ode_jl_b200 as m
from oracle import oracle as O

pytestmark = pytest.mark.gpu

RHS = {
    "lorenz": (3, lambda r: [10.0 * r.uniform(0.5, 1.5), 28.0 * r.uniform(0.5, 1.2), 8 / 3 * r.uniform(0.5, 1.5)],
               lambda r, n: np.stack([r.uniform(-10, 10, n), r.uniform(-10, 10, n), r.uniform(5, 35, n)], 1)),
    "vdp": (2, lambda r: [r.uniform(0.1, 8.0)], lambda r, n: r.uniform(-2.5, 2.5, (n, 2))),
    "robertson": (3, lambda r: list(np.array([0.04, 3e7, 1e4]) * 10 ** r.uniform(-0.5, 0.5, 3)),
                  lambda r, n: np.tile([1.0, 0.0, 0.0], (n, 1)) + np.abs(r.normal(0, 1e-3, (n, 3)))),
    "kepler": (4, lambda r: [], lambda r, n: np.stack([1 - (e := r.uniform(0, 0.8, n)), np.zeros(n), np.zeros(n),
                                                        np.sqrt((1 + e) / (1 - e))], 1)),
    "rotation": (2, lambda r: [], lambda r, n: r.uniform(-3, 3, (n, 2))),
    "linear": (1, lambda r: [r.uniform(-3, 1.5)], lambda r, n: r.uniform(0.1, 3, (n, 1))),
    "timelin": (1, lambda r: [r.uniform(-3, 3)], lambda r, n: r.uniform(-1, 1, (n, 1))),
}
ADAPTIVE = ["ode21", "ode23", "ode45_fe", "ode45", "ode78", "ode23s"]
FIXED = ["ode1", "ode2_midpoint", "ode2_heun", "ode4", "ode4ms", "ode4s_s", "ode4s_kr"]


def random_case(r):
    name = r.choice(list(RHS))
    dof, mkp, mky = RHS[name]
    p = mkp(r)
    method = r.choice(ADAPTIVE + FIXED, p=[0.12] * 6 + [0.04] * 7)
    span = float(10 ** r.uniform(-1.5, 0.8))
    if name == "robertson":
        span = float(10 ** r.uniform(-1, 2.5))
    backward = bool(r.random() < 0.15) and name in ("linear", "rotation", "timelin") and method != "ode23s"
    t0 = float(r.uniform(-2, 2))
    t1 = t0 - span if backward else t0 + span
    kw = {}
    if method in ADAPTIVE:
        kw["reltol"] = float(10 ** r.uniform(-9, -3))
        kw["abstol"] = float(10 ** r.uniform(-11, -5))
        if r.random() < 0.3:
            kw["maxstep"] = span / r.uniform(3, 40)
        if r.random() < 0.2 and not backward:
            kw["initstep"] = span * 10 ** r.uniform(-5, -1)
        if r.random() < 0.1:
            kw["minstep"] = span * 10 ** r.uniform(-6, -3)
        if method == "ode21":
            kw["reltol"] = max(kw["reltol"], 1e-5)
    if method in ("ode23s", "ode4s_s", "ode4s_kr") and r.random() < 0.4 and name != "timelin":
        kw["jacobian"] = True
    ninner = int(r.integers(0, 4)) if method in ADAPTIVE else int(r.integers(5, 60))
    inner = np.sort(r.uniform(0, 1, ninner))
    ts = np.concatenate([[t0], t0 + (t1 - t0) * inner, [t1]])
    ts = np.array(sorted(set(ts.tolist()), reverse=backward))
    return name, dof, p, mky, method, ts, kw


@pytest.mark.parametrize("seed", range(40))
def test_random_configurations_match_oracle(seed):
    r = np.random.default_rng(1000 + seed)
    for it in range(14):
        name, dof, p, mky, method, ts, kw = random_case(r)
        F = m.DeviceRHS(name, *p)
        okw = dict(kw)
        if "jacobian" in okw:
            okw["jacobian"] = 1
        desc = "seed %d case %d: %s %s ts=%s kw=%s p=%s" % (seed, it, method, name, ts.tolist(), kw, p)
        n = int(r.integers(1, 40))
        y0 = mky(r, n)
        got = m.solve_ensemble(method, F, y0, ts, max_steps=20000, **kw)
        ref = O.solve_ensemble(method, name, y0, ts, p if p else None, max_steps=20000, **okw)
        for k in ("n_accept", "n_reject", "n_rhs", "status", "t_final"):
            assert np.array_equal(got[k], ref[k], equal_nan=True), desc + " :: " + k
        assert np.array_equal(got["y_final"], ref["y_final"], equal_nan=True), desc + " :: y_final"
        # one trajectory through the (tout, yout) front-end in a random points mode
        i = int(r.integers(0, n))
        skw = dict(kw)
        fixed_vec = method in ("ode1", "ode2_midpoint", "ode2_heun", "ode4") and dof > 1
        if method in ADAPTIVE:
            skw["points"] = str(r.choice(["all", "specified"]))
        yi = y0[i] if (dof > 1 or r.random() < 0.5) else float(y0[i, 0])
        if method in ("ode4s_s", "ode4s_kr", "ode4ms") or fixed_vec or method in ("ode1", "ode2_midpoint", "ode2_heun", "ode4"):
            skw = {k: v for k, v in skw.items() if k == "jacobian"}
            if method in ("ode1", "ode2_midpoint", "ode2_heun", "ode4") and np.ndim(yi) > 0:
                skw = {}
        t, y = getattr(m, method)(F, yi, ts, **skw)
        okw2 = {k: (1 if k == "jacobian" else v) for k, v in skw.items()}
        o = O.solve(method, name, np.atleast_1d(y0[i]), ts, p if p else None, max_steps=0, **okw2)
        yy = y if np.ndim(y) == 2 else y[:, None]
        assert np.array_equal(t, o.t) and np.array_equal(yy, o.y, equal_nan=True), desc + " :: front-end " + str(skw)
