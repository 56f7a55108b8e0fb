#!/usr/bin/env python3
"""End-to-end time of ode_b200_solve_ensemble (host buffers) vs ODE_B200_CHUNKS for a few configs."""
import math, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import ode_jl_b200 as m
rng = np.random.default_rng(20261019)
n = 1 << 18
e = rng.uniform(0, 0.9, n)
cases = {
    "kepler78": ("ode78", m.DeviceRHS("kepler"), np.stack([1 - e, np.zeros(n), np.zeros(n), np.sqrt((1 + e) / (1 - e))], 1), [0.0, 20 * math.pi], dict(reltol=1e-12, abstol=1e-12), None),
    "rober23s": ("ode23s", m.DeviceRHS("robertson", 0.04, 3e7, 1e4), np.tile([1.0, 0.0, 0.0], (n, 1)), [0.0, 1e5], dict(reltol=1e-8, abstol=1e-8), np.array([0.04, 3e7, 1e4]) * 10 ** rng.uniform(-0.5, 0.5, (n, 3))),
    "lorenz45": ("ode45", m.DeviceRHS("lorenz", 10.0, 28.0, 8 / 3), np.stack([rng.uniform(-10, 10, n), rng.uniform(-10, 10, n), rng.uniform(5, 35, n)], 1), [0.0, 10.0], {}, None),
}
for name, (alg, F, y0, ts, kw, par) in cases.items():
    for c in (1, 2, 4, 8, 16):
        os.environ["ODE_B200_CHUNKS"] = str(c)
        best = 1e9
        for _ in range(3):
            t0 = time.perf_counter()
            m.solve_ensemble(alg, F, y0, ts, params=par, **kw)
            best = min(best, time.perf_counter() - t0)
        print("%-9s chunks %2d  %.2f ms" % (name, c, best * 1e3))
