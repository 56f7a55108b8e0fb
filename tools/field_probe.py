#!/usr/bin/env python3
"""Stage-per-kernel (field.cuh) against the fused stencil kernel on the same reaction-diffusion system.
usage: python tools/field_probe.py [log2 N]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import ode_jl_b200 as m

lg = int(sys.argv[1]) if len(sys.argv) > 1 else 24
n = 1 << lg
u0 = 0.5 + 0.4 * np.sin(2 * np.pi * 8 * np.arange(n) / n)
out = np.empty(n)
res = {}
for name in ("reaction_diffusion", "reaction_diffusion_field"):
    F = m.DeviceRHS(name, 1.0, 1.0)
    for _ in range(2):
        r = m.solve_large("ode45", F, u0, [0.0, 20.0], out=out)
    att = r["n_accept"] + r["n_reject"]
    res[name] = r["y"].copy()
    print("%-26s N=2^%d: %d attempts, device %.2f ms, %.4f ms per attempt, %.0f GB/s at 384 B/point, %.0f GB/s at 32 B/point"
          % (name, lg, att, r["kernel_ms"], r["kernel_ms"] / att, 384.0 * n / (r["kernel_ms"] / att * 1e-3) / 1e9,
             32.0 * n / (r["kernel_ms"] / att * 1e-3) / 1e9))
print("same bits:", np.array_equal(res["reaction_diffusion"], res["reaction_diffusion_field"]))
