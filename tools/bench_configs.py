#!/usr/bin/env python3
"""Throughput of the other ensemble configurations of BASELINE.json (configs[2], configs[3]) and a
few extra method/RHS pairs.  Device time of the kernel (ode_b200_last_kernel_ms), host buffers."""
import math
import os
import sys

os.environ.setdefault("ODE_B200_CHUNKS", "1")  # one launch per call, so kernel_ms is the kernel alone

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import ode_jl_b200 as m

rng = np.random.default_rng(20261019)
n = int(sys.argv[1]) if len(sys.argv) > 1 else 1 << 18


RESULTS = []
# algorithmic FP64 operations per attempt, where they were counted
FLOPS = {"C3 kepler ode78 tol1e-12": 753, "C4 robertson ode23s tol1e-8": 250, "lorenz ode45": 297, "lorenz ode45_fe": 250,
         "lorenz ode23": 152}


def run(name, alg, F, y0, tspan, **kw):
    best = None
    for _ in range(3):
        r = m.solve_ensemble(alg, F, y0, tspan, **kw)
        best = r["kernel_ms"] if best is None else min(best, r["kernel_ms"])
    att = int(r["n_accept"].sum() + r["n_reject"].sum())
    print("%-34s n=%d attempts=%d (%.1f/traj, rej %.2f%%) kernel %.3f ms  %.3e attempts/s  status!=0: %d"
          % (name, len(y0), att, att / len(y0), 100.0 * r["n_reject"].sum() / att, best, att / best * 1e3,
             int((r["status"] != 0).sum())))
    RESULTS.append({"config": name, "method": alg, "n_traj": int(len(y0)), "attempts": att, "kernel_ms": best,
                    "attempts_per_s": att / best * 1e3,
                    "algorithmic_tflops": (FLOPS[name] * att / best * 1e3 / 1e12) if name in FLOPS else None})
    return r


e = rng.uniform(0, 0.9, n)
y0 = np.stack([1 - e, np.zeros(n), np.zeros(n), np.sqrt((1 + e) / (1 - e))], 1)
r = run("C3 kepler ode78 tol1e-12", "ode78", m.DeviceRHS("kepler"), y0, [0.0, 20 * math.pi], reltol=1e-12, abstol=1e-12)
en = lambda y: 0.5 * (y[:, 2] ** 2 + y[:, 3] ** 2) - 1 / np.hypot(y[:, 0], y[:, 1])
print("   energy drift max %.3e" % np.max(np.abs(en(r["y_final"]) - en(y0))))
k = np.array([0.04, 3e7, 1e4]) * 10 ** rng.uniform(-0.5, 0.5, (n, 3))
y0 = np.tile([1.0, 0.0, 0.0], (n, 1))
run("C4 robertson ode23s tol1e-8", "ode23s", m.DeviceRHS("robertson", 0.04, 3e7, 1e4), y0, [0.0, 1e5], params=k,
    reltol=1e-8, abstol=1e-8)
y0 = np.stack([rng.uniform(-10, 10, n), rng.uniform(-10, 10, n), rng.uniform(5, 35, n)], 1)
L = m.DeviceRHS("lorenz", 10.0, 28.0, 8 / 3)
for alg in ("ode23", "ode45_fe", "ode45", "ode78"):
    run("lorenz " + alg, alg, L, y0, [0.0, 10.0])
y0 = np.stack([rng.uniform(-2, 2, n), rng.uniform(-2, 2, n)], 1)
run("vdp mu=5 ode45", "ode45", m.DeviceRHS("vdp", 5.0), y0, [0.0, 20.0])
run("vdp mu=50 ode23s", "ode23s", m.DeviceRHS("vdp", 50.0), y0, [0.0, 20.0])

if os.environ.get("ODE_B200_BENCH_CONFIGS_JSON"):
    import json
    with open(os.environ["ODE_B200_BENCH_CONFIGS_JSON"], "w") as f:
        json.dump(RESULTS, f, indent=1)
