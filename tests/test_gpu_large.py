"""GPU parity for the large-system regime (fused attempt kernel + device control)
against the CPU oracle: same attempt trace (t, dt, err, accept) and same final
state, bit for bit, for the single slab and for emulated multi-rank slabs."""
import math

import numpy as np
import pytest

import ode_jl_b200 as m
from oracle import oracle as O

pytestmark = pytest.mark.gpu
RD = [1.0, 1.0]


def u0(n):
    i = np.arange(n)
    return 0.5 + 0.4 * np.sin(2 * np.pi * 8 * i / n)


def oracle_rd(method, y0, tspan, **kw):
    return O.solve(method, "reaction_diffusion", y0, tspan, RD, trace=True, points="final", **kw)


@pytest.mark.parametrize("n", [4096, 5000, 1 << 16])
def test_reaction_diffusion_dopri5_bit_exact(n):
    """Config 5 at oracle-sized N: identical accept/reject sequence, dt, err and final state."""
    y0 = u0(n)
    got = m.solve_large("ode45", m.DeviceRHS("reaction_diffusion", *RD), y0, [0.0, 20.0], trace=True)
    ref = oracle_rd("ode45", y0, [0.0, 20.0])
    assert got["status"] == 0 and got["t_final"] == ref.t[-1]
    assert (got["n_accept"], got["n_reject"], got["n_rhs"]) == (ref.n_accept, ref.n_reject, ref.n_rhs)
    assert np.array_equal(got["trace"], ref.trace)
    assert np.array_equal(got["y"], ref.y[-1])


@pytest.mark.parametrize("method", ["ode21", "ode23", "ode45_fe", "ode45", "ode78", "ode45_ck"])
def test_all_tableaus_large_bit_exact(method):
    n = 3000
    y0 = u0(n) + 0.05 * np.cos(2 * np.pi * 3 * np.arange(n) / n)
    kw = dict(reltol=1e-4) if method == "ode21" else {}
    tspan = [0.0, 0.5] if method == "ode21" else [0.0, 3.0]
    got = m.solve_large(method, m.DeviceRHS("reaction_diffusion", 0.7, 1.3), y0, tspan, trace=True, **kw)
    ref = O.solve(method, "reaction_diffusion", y0, tspan, [0.7, 1.3], trace=True, points="final", **kw)
    assert np.array_equal(got["trace"], ref.trace)
    assert np.array_equal(got["y"], ref.y[-1])
    assert (got["n_accept"], got["n_reject"], got["n_rhs"]) == (ref.n_accept, ref.n_reject, ref.n_rhs)


def lockstep(engines, sends, recvs, poll_every=4):
    """Run R slab sessions of one process in lockstep; the 'all-gather' is a device copy."""
    import torch

    def exchange():
        allv = torch.cat([s.view(1, -1) for s in sends], 0)
        for r in recvs:
            r.copy_(allv)

    for ph in range(4):
        for e, s, r in zip(engines, sends, recvs):
            e.init_phase(ph, s.data_ptr(), r.data_ptr())
        if ph < 3:
            exchange()
    while True:
        for _ in range(poll_every):
            for e, s in zip(engines, sends):
                e.attempt(s.data_ptr())
            exchange()
            for e, r in zip(engines, recvs):
                e.control(r.data_ptr())
        ctls = [e.poll() for e in engines]
        if all(c["done"] for c in ctls):
            return ctls


@pytest.mark.parametrize("world", [2, 3, 8])
def test_slab_decomposition_equals_single_slab(world):
    """Domain decomposition with ghost refresh + gathered error partials gives the same bits as
    one slab (ranks emulated as sequential sessions on one GPU; no kernel waits on another)."""
    import torch
    n = 20000
    y0 = u0(n)
    F = m.DeviceRHS("reaction_diffusion", *RD)
    single = m.solve_large("ode45", F, y0, [0.0, 20.0], trace=True)
    dev = torch.device("cuda", 0)
    st = torch.cuda.current_stream(dev).cuda_stream
    engines, sends, recvs, parts = [], [], [], []
    for r in range(world):
        lo, hi = m.shard_bounds(n, world, r)
        e = m.CudaSlabEngine("ode45", F, hi - lo, n, r, world, [0.0, 20.0], st)
        yl = torch.tensor(y0[lo:hi], dtype=torch.float64, device=dev)
        e.set_state(yl.data_ptr())
        engines.append(e)
        parts.append(yl)
        sends.append(torch.zeros(m._lib.XCHG_DOUBLES, dtype=torch.float64, device=dev))
        recvs.append(torch.zeros(world, m._lib.XCHG_DOUBLES, dtype=torch.float64, device=dev))
    trs = torch.zeros(4096, 4, dtype=torch.float64, device=dev)
    engines[0].set_trace(trs.data_ptr(), 4096)
    ctls = lockstep(engines, sends, recvs)
    out = []
    for e, yl in zip(engines, parts):
        o = torch.empty_like(yl)
        e.get_state(o.data_ptr())
        out.append(o)
    torch.cuda.synchronize()
    y = torch.cat(out).cpu().numpy()
    assert all(c["status"] == 0 for c in ctls)
    assert len({(c["n_accept"], c["n_reject"], c["t_final"]) for c in ctls}) == 1
    assert (ctls[0]["n_accept"], ctls[0]["n_reject"]) == (single["n_accept"], single["n_reject"])
    assert np.array_equal(trs[:ctls[0]["trace_n"]].cpu().numpy(), single["trace"])
    assert np.array_equal(y, single["y"])
    for e in engines:
        e.close()


WIDE_BODY = """
const double x = (double)i / (double)n;
const double lap = (((-w[0] + 16.0 * w[1]) - 30.0 * w[2]) + 16.0 * w[3]) - w[4];
return (p[0] * (1.0 + 0.5 * x)) * (lap / 12.0) + p[1] * ((x - p[2] * t) * w[2]);
"""


@pytest.mark.parametrize("world", [2, 3])
def test_slab_decomposition_radius2_index_dependent_stencil(world):
    """A registered 5-point stencil whose coefficients depend on the global grid index and on time: the slabs
    must agree on the index base (uneven split: n % world != 0) and exchange 2-wide halos per RHS sweep."""
    import torch
    n = 20011
    y0 = u0(n)
    F = m.register_stencil("wide_slab_%d" % world, WIDE_BODY, params=(0.6, 0.3, 0.1), radius=2)
    single = m.solve_large("ode45", F, y0, [0.0, 2.0], trace=True)
    dev = torch.device("cuda", 0)
    st = torch.cuda.current_stream(dev).cuda_stream
    engines, sends, recvs, parts = [], [], [], []
    for r in range(world):
        lo, hi = m.shard_bounds(n, world, r)
        e = m.CudaSlabEngine("ode45", F, hi - lo, n, r, world, [0.0, 2.0], st)
        yl = torch.tensor(y0[lo:hi], dtype=torch.float64, device=dev)
        e.set_state(yl.data_ptr())
        engines.append(e)
        parts.append(yl)
        sends.append(torch.zeros(m._lib.XCHG_DOUBLES, dtype=torch.float64, device=dev))
        recvs.append(torch.zeros(world, m._lib.XCHG_DOUBLES, dtype=torch.float64, device=dev))
    ctls = lockstep(engines, sends, recvs)
    out = []
    for e, yl in zip(engines, parts):
        o = torch.empty_like(yl)
        e.get_state(o.data_ptr())
        out.append(o)
    torch.cuda.synchronize()
    y = torch.cat(out).cpu().numpy()
    assert all(c["status"] == 0 for c in ctls)
    assert (ctls[0]["n_accept"], ctls[0]["n_reject"]) == (single["n_accept"], single["n_reject"])
    assert np.array_equal(y, single["y"])
    for e in engines:
        e.close()


def test_solve_large_distributed_world1():
    import torch
    n = 10000
    y0 = u0(n)
    F = m.DeviceRHS("reaction_diffusion", *RD)
    single = m.solve_large("ode45", F, y0, [0.0, 5.0])
    y, ctl = m.solve_large_distributed("ode45", F, torch.tensor(y0, device="cuda"), n, [0.0, 5.0])
    assert np.array_equal(y.cpu().numpy(), single["y"]) and ctl["n_accept"] == single["n_accept"]


def test_large_properties_at_full_size_prefix():
    """N = 2^22 (beyond what the oracle covers in the test budget): size-independent checks --
    the solution stays in [0, 1.0000001], relaxes towards u = 1, is 8-fold periodic like u0, and the
    step count grows with N as the 2-norm (not RMS) controller implies."""
    n = 1 << 22
    y0 = u0(n)
    got = m.solve_large("ode45", m.DeviceRHS("reaction_diffusion", *RD), y0, [0.0, 20.0], trace=True)
    y = got["y"]
    assert got["status"] == 0 and got["t_final"] == 20.0
    assert y.min() > 0.9 and y.max() <= 1.0 + 1e-7
    assert np.max(np.abs(y - np.roll(y, n // 8))) < 1e-6
    small = m.solve_large("ode45", m.DeviceRHS("reaction_diffusion", *RD), u0(1 << 14), [0.0, 20.0])
    assert got["n_accept"] >= small["n_accept"]
    tr = got["trace"]
    assert np.all(np.diff(tr[tr[:, 3] == 1.0, 0]) > 0)


def test_large_error_paths():
    F = m.DeviceRHS("reaction_diffusion", *RD)
    with pytest.raises(ValueError, match="Zero time span"):
        m.solve_large("ode45", F, u0(4096), [1.0, 1.0])
    import torch
    with pytest.raises(ValueError, match="adaptive RK Butcher table"):  # the adaptive driver with a fixed-step tableau
        m.CudaSlabEngine("ode4", F, 4096, 4096, 0, 1, [0.0, 1.0], torch.cuda.current_stream().cuda_stream)
    with pytest.raises(m.OdeB200Error):  # the Rosenbrock solvers need a dense N x N Jacobian: ensemble path only
        m.solve_large("ode23s", F, u0(4096), [0.0, 1.0])
    with pytest.raises(m.OdeB200Error):
        m.solve_large("ode45", F, u0(8), [0.0, 1.0])
    got = m.solve_large("ode45", F, u0(4096), [0.0, 20.0], max_steps=7)
    assert got["status"] == 2 and got["n_accept"] + got["n_reject"] == 7


@pytest.mark.parametrize("method", ["ode45", "ode23"])
def test_large_points_specified_bit_exact(method):
    """points=:specified on the large system: Hermite rows at interior tspan points (src/runge_kutta.jl:321-326,479-492)."""
    n = 6000
    y0 = u0(n)
    ts = [0.0, 0.013, 0.5, 0.51, 3.0, 7.5, 10.0]
    got = m.solve_large(method, m.DeviceRHS("reaction_diffusion", *RD), y0, ts, points="specified", trace=True)
    ref = O.solve(method, "reaction_diffusion", y0, ts, RD, trace=True, points="specified")
    assert got["y"].shape == (len(ts), n)
    assert np.array_equal(got["trace"], ref.trace)
    assert np.array_equal(got["y"], ref.y)
    assert np.array_equal(got["t"], ref.t)


def test_large_nan_state_and_backward_direction():
    """NaN / overflow in the state (src/runge_kutta.jl:432,358) and backward integration over a short span on a
    smooth state (test/runtests.jl:58 analogue)."""
    n = 4096
    F = m.DeviceRHS("reaction_diffusion", *RD)
    # (1) NaN in u0: hinit already yields dt = NaN; one rejected attempt, then the non-finite-dt guard
    y0 = u0(n)
    y0[1234] = np.nan
    got = m.solve_large("ode45", F, y0, [0.0, 1.0], trace=True, max_steps=400)
    ref = O.solve("ode45", "reaction_diffusion", y0, [0.0, 1.0], RD, trace=True, points="final", max_steps=400)
    assert got["status"] == ref.status == 3 and (got["n_accept"], got["n_reject"]) == (ref.n_accept, ref.n_reject)
    assert np.array_equal(got["trace"], ref.trace, equal_nan=True)
    # (2) an overflowing point with a forced finite first step: every attempt trips the NaN guard, dt shrinks by 5
    #     per attempt (err = 10, timeout = 5) until |dt| < minstep stops the integration with a truncated result
    y0 = u0(n)
    y0[77] = 1e160
    kw = dict(initstep=0.1, minstep=1e-9)
    got = m.solve_large("ode45", F, y0, [0.0, 1.0], trace=True, **kw)
    ref = O.solve("ode45", "reaction_diffusion", y0, [0.0, 1.0], RD, trace=True, points="final", **kw)
    assert got["status"] == ref.status == 1 and got["n_accept"] == ref.n_accept == 0 and got["n_reject"] == ref.n_reject > 5
    assert np.array_equal(got["trace"], ref.trace, equal_nan=True) and np.all(got["trace"][:, 2] == 10.0)
    assert got["t_final"] == 0.0 and np.array_equal(got["y"], y0)
    # (3) backward
    y0 = 0.5 + 0.001 * np.sin(2 * np.pi * np.arange(n) / n)
    got = m.solve_large("ode45", m.DeviceRHS("reaction_diffusion", 1e-3, 1.0), y0, [0.5, 0.0], trace=True)
    ref = O.solve("ode45", "reaction_diffusion", y0, [0.5, 0.0], [1e-3, 1.0], trace=True, points="final")
    assert np.array_equal(got["trace"], ref.trace) and np.array_equal(got["y"], ref.y[-1]) and got["t_final"] == 0.0


@pytest.mark.parametrize("method", ["ode45", "ode23", "ode45_fe", "ode78"])
def test_large_overflowing_first_step_recovers_like_the_reference(method):
    """A first step far too long overflows the stages to Inf.  The reference multiplies zero tableau coefficients,
    so Inf*0 = NaN reaches the trial state, the NaN guard (:432) cuts dt by 5 and the integration recovers.  The
    fused kernel drops zero coefficients, notices the non-finite result and repeats the attempt with its literal
    instantiation: same trace, same final state, bit for bit -- single slab and two slabs."""
    import torch
    n = 6000
    y0 = u0(n)
    F = m.DeviceRHS("reaction_diffusion", *RD)
    kw = dict(initstep=1e7)
    got = m.solve_large(method, F, y0, [0.0, 1.0], trace=True, **kw)
    ref = O.solve(method, "reaction_diffusion", y0, [0.0, 1.0], RD, trace=True, points="final", **kw)
    if method in ("ode45", "ode78"):
        assert ref.status == 0 and np.any(ref.trace[:, 2] == 10.0)  # the NaN guard fired and the run recovered
    # (ode45_fe meets err = Inf -> dt = NaN, where the reference spins forever and both sides stop with status 3)
    assert got["status"] == ref.status
    assert (got["n_accept"], got["n_reject"], got["n_rhs"]) == (ref.n_accept, ref.n_reject, ref.n_rhs)
    assert np.array_equal(got["trace"], ref.trace, equal_nan=True)
    assert np.array_equal(got["y"], ref.y[-1])
    if method != "ode45":
        return
    dev = torch.device("cuda", 0)
    st = torch.cuda.current_stream(dev).cuda_stream
    world = 2
    engines, sends, recvs, parts = [], [], [], []
    for r in range(world):
        lo, hi = m.shard_bounds(n, world, r)
        e = m.CudaSlabEngine(method, F, hi - lo, n, r, world, [0.0, 1.0], st, **kw)
        yl = torch.tensor(y0[lo:hi], dtype=torch.float64, device=dev)
        e.set_state(yl.data_ptr())
        engines.append(e)
        parts.append(yl)
        sends.append(torch.zeros(m._lib.XCHG_DOUBLES, dtype=torch.float64, device=dev))
        recvs.append(torch.zeros(world, m._lib.XCHG_DOUBLES, dtype=torch.float64, device=dev))
    ctls = lockstep(engines, sends, recvs)
    out = []
    for e, yl in zip(engines, parts):
        o = torch.empty_like(yl)
        e.get_state(o.data_ptr())
        out.append(o)
    torch.cuda.synchronize()
    assert (ctls[0]["n_accept"], ctls[0]["n_reject"]) == (ref.n_accept, ref.n_reject)
    assert np.array_equal(torch.cat(out).cpu().numpy(), ref.y[-1])
    for e in engines:
        e.close()


@pytest.mark.parametrize("method", ["ode45", "ode23"])
def test_large_points_all_through_the_front_end(method):
    """odeXX(F, u0, tspan) with the default points=:all on a dof = N state: every accepted step plus the
    Hermite values at interior tspan points (src/runge_kutta.jl:327-340), bit for bit."""
    n = 3000
    y0 = u0(n)
    ts = [0.0, 0.013, 0.5, 3.0, 4.0]
    F = m.DeviceRHS("reaction_diffusion", *RD)
    t, y, st = getattr(m, method).stats(F, y0, ts)
    ref = O.solve(method, "reaction_diffusion", y0, ts, RD, points="all")
    assert np.array_equal(t, ref.t) and np.array_equal(y, ref.y)
    assert (st["n_accept"], st["n_reject"], st["n_rhs"]) == (ref.n_accept, ref.n_reject, ref.n_rhs)
    # long integration: more rows than the first guess -> deterministic rerun with the reported size
    t2, y2 = getattr(m, method)(F, y0, [0.0, 20.0], reltol=1e-7)
    ref2 = O.solve(method, "reaction_diffusion", y0, [0.0, 20.0], RD, points="all", reltol=1e-7)
    assert len(t2) == len(ref2.t) > 66 and np.array_equal(t2, ref2.t) and np.array_equal(y2, ref2.y)
    t3, y3 = getattr(m, method)(F, y0, ts, points="specified")
    ref3 = O.solve(method, "reaction_diffusion", y0, ts, RD, points="specified")
    assert np.array_equal(t3, ref3.t) and np.array_equal(y3, ref3.y)


# ---- general gather right-hand sides ("fields"): one kernel per stage (csrc/field.cuh) -------------------------
@pytest.mark.parametrize("method", ["ode21", "ode23", "ode45_fe", "ode45", "ode78", "ode45_ck"])
def test_field_path_equals_fused_stencil_path_and_oracle(method):
    """The built-in reaction_diffusion_field is the reaction-diffusion system written as a gather over the global
    state: the stage-per-kernel path must give the fused kernel's bits, and the oracle's."""
    n = 5000
    y0 = u0(n)
    ts = [0.0, 0.7, 3.0]
    a = m.solve_large(method, m.DeviceRHS("reaction_diffusion_field", *RD), y0, [0.0, 3.0], trace=True)
    b = m.solve_large(method, m.DeviceRHS("reaction_diffusion", *RD), y0, [0.0, 3.0], trace=True)
    assert np.array_equal(a["y"], b["y"]) and np.array_equal(a["trace"], b["trace"])
    assert (a["n_accept"], a["n_reject"], a["n_rhs"]) == (b["n_accept"], b["n_reject"], b["n_rhs"])
    for points in ("specified", "all"):
        t, y, st = getattr(m, method).stats(m.DeviceRHS("reaction_diffusion_field", *RD), y0, ts, points=points)
        o = O.solve(method, "reaction_diffusion", y0, ts, RD, points=points)
        assert np.array_equal(t, o.t) and np.array_equal(y, o.y), points
        assert (st["n_accept"], st["n_reject"], st["n_rhs"]) == (o.n_accept, o.n_reject, o.n_rhs)


NONLOCAL_BODY = """
const long long j = (i * 7 + 3) % n;            // a long-range partner
const double mirror = y[n - 1 - i];
return (p[0] * (y[j] - y[i]) + p[1] * (mirror * y[i])) - (p[2] * t) * ((y[i] * y[i]) * y[i]);
"""


def nonlocal_host(t, y, p):
    n = len(y)
    return [(p[0] * (y[(i * 7 + 3) % n] - y[i]) + p[1] * (y[n - 1 - i] * y[i])) - (p[2] * t) * ((y[i] * y[i]) * y[i])
            for i in range(n)]


def test_user_field_nonlocal_coupling_matches_oracle():
    F = m.register_field("nonlocal", NONLOCAL_BODY, host=nonlocal_host, params=(0.8, 0.3, 0.5))
    n = 400
    y0 = 0.5 + 0.4 * np.sin(2 * np.pi * 3 * np.arange(n) / n)
    ts = [0.0, 0.4, 1.5]
    for method, points in (("ode45", "specified"), ("ode23", "all"), ("ode78", "specified"), ("ode45_fe", "all")):
        t, y, st = getattr(m, method).stats(F, y0, ts, points=points)
        o = O.solve_user(method, nonlocal_host, y0, ts, [0.8, 0.3, 0.5], points=points)
        assert np.array_equal(t, o.t) and np.array_equal(y, o.y), (method, points)
        assert (st["n_accept"], st["n_reject"], st["n_rhs"]) == (o.n_accept, o.n_reject, o.n_rhs)
    # below 32 components the ensemble path (register_rhs) is the one that restates the reference's small-vector norm
    with pytest.raises(m.OdeB200Error):
        m.ode45(F, y0[:3], [0.0, 1.0])
    assert np.allclose(F(0.3, y0[:5]), nonlocal_host(0.3, list(y0[:5]), [0.8, 0.3, 0.5]))


def test_field_overflowing_first_step_recovers_like_the_reference():
    n = 3000
    y0 = u0(n)
    kw = dict(initstep=1e7)
    got = m.solve_large("ode45", m.DeviceRHS("reaction_diffusion_field", *RD), y0, [0.0, 1.0], trace=True, **kw)
    ref = O.solve("ode45", "reaction_diffusion", y0, [0.0, 1.0], RD, trace=True, points="final", **kw)
    assert ref.status == 0 and np.any(ref.trace[:, 2] == 10.0)
    assert got["status"] == 0 and (got["n_accept"], got["n_reject"], got["n_rhs"]) == (ref.n_accept, ref.n_reject, ref.n_rhs)
    assert np.array_equal(got["trace"], ref.trace, equal_nan=True) and np.array_equal(got["y"], ref.y[-1])


@pytest.mark.parametrize("method", ["ode1", "ode2_midpoint", "ode2_heun", "ode4", "ode4ms", "ode5ms", "ode_ms1", "ode_ms2",
                                    "ode_ms3"])
def test_fixed_step_rk_on_a_large_state(method):
    """oderk_fixed (src/runge_kutta.jl:176-212) on a dof = N state: one state per tspan entry, bit for bit, for
    the built-in system (stencil and gather ids), a registered radius-2 stencil and a registered field."""
    n = 3000
    y0 = u0(n)
    ts = np.linspace(0.0, 0.5, 11)
    o = O.solve(method, "reaction_diffusion", y0, ts, RD)
    for name in ("reaction_diffusion", "reaction_diffusion_field"):
        t, y = getattr(m, method)(m.DeviceRHS(name, *RD), y0, ts)
        assert np.array_equal(t, o.t) and np.array_equal(y, o.y), name
    r = m.solve_large(method, m.DeviceRHS("reaction_diffusion", *RD), y0, ts)  # points = final
    assert np.array_equal(r["y"], o.y[-1]) and r["n_rhs"] == o.n_rhs
    n = 400
    y0 = 0.5 + 0.4 * np.sin(2 * np.pi * 3 * np.arange(n) / n)
    W = m.register_stencil("wide_fixed_" + method, WIDE_BODY, params=(0.6, 0.3, 0.1), radius=2)

    def wide_host(w, p, t, i, nn):
        x = float(i) / float(nn)
        lap = (((-w[0] + 16.0 * w[1]) - 30.0 * w[2]) + 16.0 * w[3]) - w[4]
        return (p[0] * (1.0 + 0.5 * x)) * (lap / 12.0) + p[1] * ((x - p[2] * t) * w[2])

    t, y = getattr(m, method)(W, y0, ts)
    o = O.solve_user_stencil(method, wide_host, y0, ts, [0.6, 0.3, 0.1], radius=2)
    assert np.array_equal(t, o.t) and np.array_equal(y, o.y)
    F = m.register_field("nonlocal_fixed_" + method, NONLOCAL_BODY, params=(0.8, 0.3, 0.5))
    t, y = getattr(m, method)(F, y0, ts)
    o = O.solve_user(method, nonlocal_host, y0, ts, [0.8, 0.3, 0.5])
    assert np.array_equal(t, o.t) and np.array_equal(y, o.y)


# ---- C5 at its stated sizes ("N = 2^16 ... 2^20 full span and N = 2^26 for the first 5 attempts")
def test_reaction_diffusion_2_20_full_span_bit_exact():
    n = 1 << 20
    y0 = u0(n)
    got = m.solve_large("ode45", m.DeviceRHS("reaction_diffusion", *RD), y0, [0.0, 20.0], trace=True)
    O.use_all_cores()
    ref = oracle_rd("ode45", y0, [0.0, 20.0])
    assert got["status"] == 0 and got["t_final"] == ref.t[-1]
    assert (got["n_accept"], got["n_reject"], got["n_rhs"]) == (ref.n_accept, ref.n_reject, ref.n_rhs)
    assert np.array_equal(got["trace"], ref.trace)
    assert np.array_equal(got["y"], ref.y[-1])


def test_reaction_diffusion_2_26_first_five_attempts_bit_exact():
    """BASELINE.json configs[4] at its full size N = 2^26: the first five step attempts (t, dt, err, accept) and the
    state they lead to, bit for bit against the oracle (which takes ~20 s for them on 8 cores)."""
    n = 1 << 26
    y0 = u0(n)
    got = m.solve_large("ode45", m.DeviceRHS("reaction_diffusion", *RD), y0, [0.0, 20.0], trace=True, max_steps=5)
    O.use_all_cores()
    ref = oracle_rd("ode45", y0, [0.0, 20.0], max_steps=5)
    assert got["status"] == ref.status == 2  # the max_steps guard
    assert (got["n_accept"], got["n_reject"], got["n_rhs"]) == (ref.n_accept, ref.n_reject, ref.n_rhs)
    assert got["n_accept"] + got["n_reject"] == 5
    assert np.array_equal(got["trace"], ref.trace)
    assert got["t_final"] == ref.t[-1]
    assert np.array_equal(got["y"], ref.y[-1])
    m.lib().ode_b200_trim()


# ---- wide ghost zones: ode78 with a radius-3 / radius-4 stencil has H = 40 / 52 > one warp ------------------------
def _wide_stencil(radius):
    """A (2R+1)-point smoothing stencil with index- and time-dependent weights; CUDA body and host twin."""
    terms = " + ".join("(%.17g * w[%d])" % (1.0 / (1 + abs(k - radius)), k) for k in range(2 * radius + 1))
    norm = sum(1.0 / (1 + abs(k - radius)) for k in range(2 * radius + 1))
    body = """
const double x = (double)i / (double)n;
const double avg = (%s) / %.17g;
return (p[0] * (1.0 + 0.25 * x)) * (avg - w[%d]) + (p[1] * t) * (w[%d] * (1.0 - w[%d]));
""" % (terms, norm, radius, radius, radius)

    def host(w, p, t, i, nn):
        x = float(i) / float(nn)
        s = None
        for k in range(2 * radius + 1):
            term = (1.0 / (1 + abs(k - radius))) * w[k]
            s = term if s is None else s + term
        avg = s / norm
        return (p[0] * (1.0 + 0.25 * x)) * (avg - w[radius]) + (p[1] * t) * (w[radius] * (1.0 - w[radius]))

    return body, host


@pytest.mark.parametrize("radius", [3, 4])
def test_ode78_wide_stencil_ghost_zone_beyond_one_warp(radius):
    """H = 40 (radius 3) and H = 52 (radius 4): the hinit edge packing / ghost fill must cover all H cells, on one
    slab (periodic wrap) and across two slabs.  Oracle: the C++ restatement driven by the python twin of the stencil."""
    import torch
    n = 1500
    y0 = u0(n) + 0.03 * np.cos(2 * np.pi * 5 * np.arange(n) / n)
    body, host = _wide_stencil(radius)
    F = m.register_stencil("wide_r%d_ode78" % radius, body, params=(0.9, 0.4), radius=radius)
    ts = [0.0, 3.0]
    got = m.solve_large("ode78", F, y0, ts, trace=True)
    ref = O.solve_user_stencil("ode78", host, y0, ts, [0.9, 0.4], radius=radius, trace=True, points="final")
    assert np.array_equal(got["trace"], ref.trace)
    assert np.array_equal(got["y"], ref.y[-1])
    assert (got["n_accept"], got["n_reject"], got["n_rhs"]) == (ref.n_accept, ref.n_reject, ref.n_rhs)
    dev = torch.device("cuda", 0)
    st = torch.cuda.current_stream(dev).cuda_stream
    world = 2
    engines, sends, recvs, parts = [], [], [], []
    for r in range(world):
        lo, hi = m.shard_bounds(n, world, r)
        e = m.CudaSlabEngine("ode78", F, hi - lo, n, r, world, ts, st)
        yl = torch.tensor(y0[lo:hi], dtype=torch.float64, device=dev)
        e.set_state(yl.data_ptr())
        engines.append(e)
        parts.append(yl)
        sends.append(torch.zeros(m._lib.XCHG_DOUBLES, dtype=torch.float64, device=dev))
        recvs.append(torch.zeros(world, m._lib.XCHG_DOUBLES, dtype=torch.float64, device=dev))
    ctls = lockstep(engines, sends, recvs)
    assert ctls[0]["ghost"] == (40 if radius == 3 else 52)
    out = []
    for e, yl in zip(engines, parts):
        o = torch.empty_like(yl)
        e.get_state(o.data_ptr())
        out.append(o)
    torch.cuda.synchronize()
    assert (ctls[0]["n_accept"], ctls[0]["n_reject"]) == (ref.n_accept, ref.n_reject)
    assert np.array_equal(torch.cat(out).cpu().numpy(), ref.y[-1])
    for e in engines:
        e.close()


# ---- output rows around rejected attempts ---------------------------------------------------------------------------
@pytest.mark.parametrize("points", ["all", "specified"])
def test_large_output_rows_survive_rejected_attempts(points):
    """Dense tspan (almost every accepted step holds an interior output point) on a run with rejected attempts:
    a rejection right after such a step must not touch the rows already written (the control kernel clears its
    row selection on every invocation)."""
    n = 1 << 14
    y0 = u0(n)
    ts = np.linspace(0.0, 20.0, 161)
    F = m.DeviceRHS("reaction_diffusion", *RD)
    r = m.solve_large("ode45", F, y0, ts, points=points, trace=True)
    ref = O.solve("ode45", "reaction_diffusion", y0, ts, RD, points=points, trace=True)
    assert ref.n_reject >= 1
    assert np.array_equal(r["trace"], ref.trace)
    assert np.array_equal(r["t"], ref.t) and np.array_equal(r["y"], ref.y)


def test_slab_points_specified_with_rejections_two_ranks():
    """The same through two slab sessions with the explicit exchange (the non-fused control kernel)."""
    import torch
    n = 1 << 14
    y0 = u0(n)
    ts = np.linspace(0.0, 20.0, 41)
    F = m.DeviceRHS("reaction_diffusion", *RD)
    ref = O.solve("ode45", "reaction_diffusion", y0, ts, RD, points="specified")
    assert ref.n_reject >= 1
    dev = torch.device("cuda", 0)
    st = torch.cuda.current_stream(dev).cuda_stream
    world = 2
    engines, sends, recvs, outs = [], [], [], []
    for r in range(world):
        lo, hi = m.shard_bounds(n, world, r)
        e = m.CudaSlabEngine("ode45", F, hi - lo, n, r, world, [0.0, 20.0], st, points="specified")
        o = torch.zeros(len(ts), hi - lo, dtype=torch.float64, device=dev)
        e.set_output(ts, o.data_ptr())
        yl = torch.tensor(y0[lo:hi], dtype=torch.float64, device=dev)
        e.set_state(yl.data_ptr())
        engines.append(e)
        outs.append(o)
        sends.append(torch.zeros(m._lib.XCHG_DOUBLES, dtype=torch.float64, device=dev))
        recvs.append(torch.zeros(world, m._lib.XCHG_DOUBLES, dtype=torch.float64, device=dev))
    ctls = lockstep(engines, sends, recvs)
    torch.cuda.synchronize()
    assert (ctls[0]["n_accept"], ctls[0]["n_reject"]) == (ref.n_accept, ref.n_reject)
    assert np.array_equal(torch.cat(outs, 1).cpu().numpy(), ref.y)
    for e in engines:
        e.close()


@pytest.mark.gpu
def test_guard_zones_behind_the_slab_stay_untouched(monkeypatch):
    """compute-sanitizer is not available on every pool, so the library carries its own bounds check for the arrays
    the tile / ghost indexing writes: every state and stage array has a zeroed tail behind the last ghost cell, and
    with ODE_B200_GUARD_CHECK=1 a solve fails if any kernel wrote there.  Ragged sizes (partial last tile, sizes
    around tile multiples), every tableau (ghost widths 4 .. 24), wide registered stencils (ghost widths up to 52),
    the stage-per-kernel field path, output modes."""
    monkeypatch.setenv("ODE_B200_GUARD_CHECK", "1")
    F = m.DeviceRHS("reaction_diffusion", 1.0, 1.0)
    for n in (33, 495, 496, 497, 511, 512, 513, 700, 991, 992, 993, 4096, 5003, 65537, 300007):
        r = m.solve_large("ode45", F, u0(n), [0.0, 0.6])
        assert r["status"] == 0, n
    for method in ("ode21", "ode23", "ode45_fe", "ode45", "ode45_ck", "ode78"):
        for n in (481, 700, 5003):
            assert m.solve_large(method, F, u0(n), [0.0, 0.5])["status"] == 0, (method, n)
    for radius in (2, 3, 4):
        body, _ = _wide_stencil(radius) if radius > 2 else ("return p[0] * ((((-w[0] + 16.0*w[1]) - 30.0*w[2]) + 16.0*w[3]) - w[4]) / 12.0;", None)
        S = m.register_stencil("guard_r%d" % radius, body, params=(0.9, 0.4) if radius > 2 else (0.1,), radius=radius)
        for method in ("ode45", "ode78"):
            assert m.solve_large(method, S, u0(1500), [0.0, 0.4])["status"] == 0, (radius, method)
    G = m.DeviceRHS("reaction_diffusion_field", 1.0, 1.0)
    for n in (700, 5003):
        assert m.solve_large("ode45", G, u0(n), [0.0, 0.5])["status"] == 0
    r = m.solve_large("ode45", F, u0(5003), [0.0, 0.2, 0.5], points="specified")
    assert r["status"] == 0
    t, y = m.ode45(F, u0(700), [0.0, 0.3, 0.5])  # points = :all through the front end
    assert t[-1] == 0.5
    # the check itself: with one double of a tail poked on purpose the same solve must fail
    monkeypatch.setenv("ODE_B200_GUARD_SELFTEST", "1")
    with pytest.raises(Exception, match="guard check: 1 doubles"):
        m.solve_large("ode45", F, u0(700), [0.0, 0.5])
    monkeypatch.delenv("ODE_B200_GUARD_SELFTEST")
    assert m.solve_large("ode45", F, u0(700), [0.0, 0.5])["status"] == 0
