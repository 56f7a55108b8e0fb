"""Randomised parity sweep: many (method, RHS, tolerances, span, direction, step limits, output mode)
combinations, each a small ensemble or a single system, CUDA result vs CPU oracle bit for bit.
Seeded, so a failure reproduces; the case description is part of the assertion message."""
import math

import os

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


@pytest.mark.parametrize("seed", range(int(os.environ.get("ODE_B200_FUZZ_SEEDS", "40"))))
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


@pytest.mark.parametrize("seed", range(int(os.environ.get("ODE_B200_FUZZ_LARGE_SEEDS", "8"))))
def test_random_large_system_configurations_match_oracle(seed):
    """Random N, tableau, stencil parameters, tolerances, spans, output mode; and for a third of the cases a random
    slab decomposition (ranks emulated as sequential sessions on one GPU) that must give the single-slab bits."""
    import torch
    from test_gpu_large import lockstep
    r = np.random.default_rng(5000 + seed)
    for it in range(5):
        n = int(r.integers(200, 20000))
        method = str(r.choice(["ode23", "ode45_fe", "ode45", "ode78"]))
        kappa, rr = float(r.uniform(0.05, 1.5)), float(r.uniform(0.0, 2.0))
        i = np.arange(n)
        y0 = 0.5 + 0.4 * np.sin(2 * np.pi * int(r.integers(1, 9)) * i / n) + 0.05 * r.uniform(-1, 1, n)
        span = float(10 ** r.uniform(-1, 0.7))
        kw = dict(reltol=float(10 ** r.uniform(-8, -3)), abstol=float(10 ** r.uniform(-10, -6)))
        mode = str(r.choice(["final", "specified", "all"]))
        ts = [0.0, span] if mode == "final" else [0.0] + sorted(r.uniform(0, span, int(r.integers(1, 4))).tolist()) + [span]
        F = m.DeviceRHS("reaction_diffusion", kappa, rr)
        desc = "seed %d case %d: %s n=%d kappa=%g r=%g ts=%s kw=%s mode=%s" % (seed, it, method, n, kappa, rr, ts, kw, mode)
        got = m.solve_large(method, F, y0, ts, trace=True, points=mode, **kw)
        ref = O.solve(method, "reaction_diffusion", y0, ts, [kappa, rr], trace=True, points=mode, **kw)
        assert np.array_equal(got["trace"], ref.trace), desc
        if mode == "final":
            assert np.array_equal(got["y"], ref.y[-1]), desc
        else:
            assert np.array_equal(got["t"], ref.t) and np.array_equal(got["y"], ref.y), desc
        if it % 3 == 0:
            world = int(r.integers(2, 6))
            dev = torch.device("cuda", 0)
            st = torch.cuda.current_stream(dev).cuda_stream
            engines, sends, recvs, parts = [], [], [], []
            for k in range(world):
                lo, hi = m.shard_bounds(n, world, k)
                e = m.CudaSlabEngine(method, F, hi - lo, n, k, world, [0.0, span], st, **kw)
                yl = torch.tensor(y0[lo:hi], dtype=torch.float64, device=dev)
                e.set_state(yl.data_ptr())
                engines.append(e)
                parts.append(yl)
                sends.append(torch.zeros(m._lib.XCHG_DOUBLES, dtype=torch.float64, device=dev))
                recvs.append(torch.zeros(world, m._lib.XCHG_DOUBLES, dtype=torch.float64, device=dev))
            ctls = lockstep(engines, sends, recvs)
            outs = []
            for e, yl in zip(engines, parts):
                o = torch.empty_like(yl)
                e.get_state(o.data_ptr())
                outs.append(o)
            torch.cuda.synchronize()
            fin = O.solve(method, "reaction_diffusion", y0, [0.0, span], [kappa, rr], points="final", **kw)
            assert np.array_equal(torch.cat(outs).cpu().numpy(), fin.y[-1]), desc + " world=%d" % world
            assert ctls[0]["n_accept"] == fin.n_accept and ctls[0]["n_reject"] == fin.n_reject
            for e in engines:
                e.close()
