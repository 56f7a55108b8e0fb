#!/usr/bin/env python3
"""One ensemble configuration of BASELINE.json, `reps` launches (for ncu: keep the command short).
usage: one_config.py c2|c3|c4|large [reps] [n]"""
import math
import os
import sys

os.environ.setdefault("ODE_B200_CHUNKS", "1")
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import ode_jl_b200 as m

which = sys.argv[1]
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 1
rng = np.random.default_rng(20261019)
if which == "c2":
    n = int(sys.argv[3]) if len(sys.argv) > 3 else 1 << 20
    y0 = np.stack([rng.uniform(-10, 10, n), rng.uniform(-10, 10, n), rng.uniform(5, 35, n)], 1)
    run = lambda: m.solve_ensemble("ode45", m.DeviceRHS("lorenz", 10.0, 28.0, 8 / 3), y0, [0.0, 10.0])
elif which == "c3":
    n = int(sys.argv[3]) if len(sys.argv) > 3 else 1 << 18
    e = rng.uniform(0, 0.9, n)
    y0 = np.stack([1 - e, np.zeros(n), np.zeros(n), np.sqrt((1 + e) / (1 - e))], 1)
    run = lambda: m.solve_ensemble("ode78", m.DeviceRHS("kepler"), y0, [0.0, 20 * math.pi], reltol=1e-12, abstol=1e-12)
elif which == "c4":
    n = int(sys.argv[3]) if len(sys.argv) > 3 else 1 << 18
    rng.uniform(0, 0.9, n)  # (same stream position as tools/bench_configs.py)
    k = np.array([0.04, 3e7, 1e4]) * 10 ** rng.uniform(-0.5, 0.5, (n, 3))
    y0 = np.tile([1.0, 0.0, 0.0], (n, 1))
    run = lambda: m.solve_ensemble("ode23s", m.DeviceRHS("robertson", 0.04, 3e7, 1e4), y0, [0.0, 1e5], params=k,
                                   reltol=1e-8, abstol=1e-8)
elif which == "large":
    n = int(sys.argv[3]) if len(sys.argv) > 3 else 1 << 26
    u0 = 0.5 + 0.4 * np.sin(2 * np.pi * 8 * np.arange(n) / n)
    run = lambda: m.solve_large("ode45", m.DeviceRHS("reaction_diffusion", 1.0, 1.0), u0, [0.0, 20.0], max_steps=8)
else:
    raise SystemExit(__doc__)
for _ in range(reps):
    r = run()
att = (int(r["n_accept"].sum() + r["n_reject"].sum()) if which != "large" else r["n_accept"] + r["n_reject"])
print("%s n=%d attempts=%d kernel_ms=%.3f" % (which, n, att, r["kernel_ms"]))
