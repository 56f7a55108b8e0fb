#!/usr/bin/env python3
"""One launch of a BASELINE ensemble config for ncu: python tools/config_probe.py c3|c4 [n]"""
import math, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import ode_jl_b200 as m
which = sys.argv[1]
n = int(sys.argv[2]) if len(sys.argv) > 2 else 1 << 18
rng = np.random.default_rng(20261019)
for _ in range(3):
    if which == "c3":
        e = rng.uniform(0, 0.9, n)
        y0 = np.stack([1 - e, np.zeros(n), np.zeros(n), np.sqrt((1 + e) / (1 - e))], 1)
        r = m.solve_ensemble("ode78", m.DeviceRHS("kepler"), y0, [0.0, 20 * math.pi], reltol=1e-12, abstol=1e-12)
    else:
        k = np.array([0.04, 3e7, 1e4]) * 10 ** rng.uniform(-0.5, 0.5, (n, 3))
        r = m.solve_ensemble("ode23s", m.DeviceRHS("robertson", 0.04, 3e7, 1e4), np.tile([1.0, 0.0, 0.0], (n, 1)),
                             [0.0, 1e5], params=k, reltol=1e-8, abstol=1e-8)
    print(which, r["kernel_ms"], int(r["n_accept"].sum() + r["n_reject"].sum()))
