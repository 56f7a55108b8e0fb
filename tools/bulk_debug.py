#!/usr/bin/env python3
"""Debug aid for kernel variants of the fused large-system kernel: one accepted attempt at several sizes, state dumped
for a bitwise diff between library builds.  usage: ODE_B200_LIB=... python tools/bulk_debug.py tag"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import ode_jl_b200 as m
tag = sys.argv[1]
F = m.DeviceRHS("reaction_diffusion", 1.0, 1.0)
for n in (360000, 380000, 1 << 20):
    u0 = 0.5 + 0.4 * np.sin(2 * np.pi * 8 * np.arange(n) / n)
    r = m.solve_large("ode45", F, u0, [0.0, 20.0], trace=True, max_steps=2, initstep=1e-3)
    print(tag, n, "acc/rej", r["n_accept"], r["n_reject"], "trace", r["trace"][:2].tolist())
    np.save("/tmp/dbg_%s_%d.npy" % (tag, n), r["y"])
