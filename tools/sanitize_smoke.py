#!/usr/bin/env python3
"""Small run of every kernel family for compute-sanitizer (memcheck / racecheck / synccheck):
    compute-sanitizer --tool memcheck python tools/sanitize_smoke.py
Sizes are tiny so the run stays short under the sanitizer."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import ode_jl_b200 as m

rng = np.random.default_rng(0)
L = m.DeviceRHS("lorenz", 10.0, 28.0, 8 / 3)
y0 = np.stack([rng.uniform(-10, 10, 300), rng.uniform(-10, 10, 300), rng.uniform(5, 35, 300)], 1)
for alg in ("ode21", "ode23", "ode45_fe", "ode45", "ode78", "ode23s"):
    r = m.solve_ensemble(alg, L, y0, [0.0, 0.3])
    assert (r["status"] == 0).all(), alg
    r = m.solve_ensemble(alg, L, y0[:40], [0.0, 0.1, 0.2, 0.3], points="all", pts_cap=64, trace_traj=3)
    assert np.isin(r["status"], (0, 5)).all(), alg  # 5 = points buffer overflow (ode21 needs more than 64 points)
    r = m.solve_ensemble(alg, L, y0[:40], [0.0, 0.1, 0.2, 0.3], points="specified")
ts = np.linspace(0, 0.2, 21)
for alg in ("ode1", "ode2_midpoint", "ode2_heun", "ode4", "ode4ms", "ode4s_s", "ode4s_kr"):
    m.solve_ensemble(alg, L, y0[:100], ts)
    getattr(m, alg)(L, [1.0, 1.0, 1.0], ts)
m.solve_ensemble("ode23s", L, y0[:64], [0.0, 0.3], jacobian=True)
m.solve_ensemble("ode78", m.DeviceRHS("kepler"), np.tile([0.6, 0.0, 0.0, 2.0], (64, 1)), [0.0, 1.0], reltol=1e-10, abstol=1e-10)
t, y = m.ode45(m.DeviceRHS("linear", 1.01), 0.5, [0.0, 1.0], reltol=1e-8, abstol=1e-8)
assert len(t) == 12
F = m.DeviceRHS("reaction_diffusion", 1.0, 1.0)
for n in (700, 4096, 5003):
    i = np.arange(n)
    u0 = 0.5 + 0.4 * np.sin(2 * np.pi * 8 * i / n)
    for alg in ("ode45", "ode23", "ode78"):
        r = m.solve_large(alg, F, u0, [0.0, 1.0], trace=True)
        assert r["status"] == 0
    r = m.solve_large("ode45", F, u0, [0.0, 0.3, 1.0], points="specified")
print("sanitize smoke ok")
