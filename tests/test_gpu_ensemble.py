"""GPU parity: the CUDA kernels (through the C ABI) against the CPU oracle, bit for bit.

Every comparison here is exact: same accept/reject sequence, same step and RHS
counts, same final bits.  The oracle runs in mathmode 0 (the shared deterministic
pow/log10); test_oracle.py ties that to libm and to the reference's known answers.
"""
import math

import numpy as np
import pytest

import ode_jl_b200 as m
from oracle import oracle as O

pytestmark = pytest.mark.gpu

LORENZ = [10.0, 28.0, 8.0 / 3.0]
ADAPTIVE = ["ode21", "ode23", "ode45_fe", "ode45", "ode78", "ode45_ck"]  # ode45_ck: the Cash-Karp extension
FIXED = ["ode1", "ode2_midpoint", "ode2_heun", "ode4"]


def lorenz_ics(n, seed=20261019):
    rng = np.random.default_rng(seed)
    return np.stack([rng.uniform(-10, 10, n), rng.uniform(-10, 10, n), rng.uniform(5, 35, n)], 1)


def assert_same(got, ref, keys=("y_final", "t_final", "n_accept", "n_reject", "n_rhs", "status")):
    for k in keys:
        a, b = np.asarray(got[k]), np.asarray(ref[k])
        if not np.array_equal(a, b, equal_nan=a.dtype.kind == "f"):
            bad = np.flatnonzero(np.any(np.atleast_2d((a != b).reshape(len(a), -1)), axis=1))
            raise AssertionError("%s differs on %d of %d trajectories (first %s)" % (k, bad.size, len(a), bad[:5]))


def test_device_detmath_equals_host_detmath():
    import ctypes as C
    rng = np.random.default_rng(1)
    n = 200000
    x = np.exp(rng.uniform(-40, 40, n))
    x[:8] = [0.0, 1.0, np.inf, 5e-324, 1e-310, 10.0, 2.0, 1e308]
    y = np.array([0.5, 1 / 3, 0.2, 0.125, -0.7, 3.3])[rng.integers(0, 6, n)]
    op, ol = np.empty(n), np.empty(n)
    dp = C.POINTER(C.c_double)
    c = lambda a: a.ctypes.data_as(dp)
    m._lib.check(m.lib().ode_b200_selftest_detmath(0, c(x), c(y), c(op), c(ol), n))
    hp, hl = np.empty(n), np.empty(n)
    O.lib().oracle_pow_v(c(x), c(y), c(hp), n, 0)
    O.lib().oracle_log10_v(c(x), c(hl), n, 0)
    assert np.array_equal(op, hp, equal_nan=True)
    assert np.array_equal(ol, hl, equal_nan=True)


@pytest.mark.parametrize("tend", [2.0, 10.0])
def test_lorenz_dopri5_ensemble_bit_exact(tend):
    """Config 2 strict-parity subset: first 4096 trajectories."""
    y0 = lorenz_ics(4096)
    got = m.solve_ensemble("ode45", m.DeviceRHS("lorenz", *LORENZ), y0, [0.0, tend])
    ref = O.solve_ensemble("ode45", "lorenz", y0, [0.0, tend], LORENZ)
    assert_same(got, ref)
    assert np.all(got["status"] == 0) and np.all(got["t_final"] == tend)


@pytest.mark.parametrize("method", ADAPTIVE)
def test_all_adaptive_tableaus_bit_exact(method):
    n = 256
    y0 = lorenz_ics(n, 5)
    got = m.solve_ensemble(method, m.DeviceRHS("lorenz", *LORENZ), y0, [0.0, 1.5], reltol=1e-4)
    ref = O.solve_ensemble(method, "lorenz", y0, [0.0, 1.5], LORENZ, reltol=1e-4)
    assert_same(got, ref)
    rng = np.random.default_rng(9)
    y0 = np.stack([rng.uniform(-2, 2, n), rng.uniform(-2, 2, n)], 1)
    mu = rng.uniform(0.5, 5.0, (n, 1))
    got = m.solve_ensemble(method, m.DeviceRHS("vdp", 1.0), y0, [0.0, 4.0], params=mu)
    ref = O.solve_ensemble(method, "vdp", y0, [0.0, 4.0], mu)
    assert_same(got, ref)
    # backward in time on a smooth problem (test/runtests.jl:58)
    y0 = rng.uniform(0.5, 2.0, (n, 1))
    got = m.solve_ensemble(method, m.DeviceRHS("linear", 1.0), y0, [1.0, 0.0])
    ref = O.solve_ensemble(method, "linear", y0, [1.0, 0.0], [1.0])
    assert_same(got, ref)


def test_kepler_feh78_tight_tolerance_bit_exact_and_energy():
    """Config 3: ode78, tol 1e-12, 10 periods; energy drift check."""
    rng = np.random.default_rng(20261019)
    n = 512
    e = rng.uniform(0, 0.9, n)
    y0 = np.stack([1 - e, np.zeros(n), np.zeros(n), np.sqrt((1 + e) / (1 - e))], 1)
    ts = [0.0, 20 * math.pi]
    got = m.solve_ensemble("ode78", m.DeviceRHS("kepler"), y0, ts, reltol=1e-12, abstol=1e-12)
    ref = O.solve_ensemble("ode78", "kepler", y0, ts, None, reltol=1e-12, abstol=1e-12)
    assert_same(got, ref)
    en = lambda y: 0.5 * (y[:, 2] ** 2 + y[:, 3] ** 2) - 1 / np.hypot(y[:, 0], y[:, 1])
    assert np.max(np.abs(en(got["y_final"]) - en(y0))) < 2e-9


def test_robertson_ode23s_ensemble_bit_exact():
    """Config 4: per-trajectory rate constants, FD Jacobian, in-register 3x3 LU."""
    rng = np.random.default_rng(20261019)
    n = 512
    k = np.array([0.04, 3e7, 1e4]) * 10 ** rng.uniform(-0.5, 0.5, (n, 3))
    y0 = np.tile([1.0, 0.0, 0.0], (n, 1))
    kw = dict(reltol=1e-8, abstol=1e-8)
    got = m.solve_ensemble("ode23s", m.DeviceRHS("robertson", 0.04, 3e7, 1e4), y0, [0.0, 1e5], params=k, **kw)
    ref = O.solve_ensemble("ode23s", "robertson", y0, [0.0, 1e5], k, **kw)
    assert_same(got, ref)
    assert np.all(np.abs(got["y_final"].sum(1) - 1.0) < 1e-6)


def test_rober_kat_on_gpu():
    """test/runtests.jl:78-96 through the ode23s front-end."""
    F = m.DeviceRHS("robertson", 0.04, 3e7, 1e4)
    t, y, st = m.ode23s.stats(F, [1.0, 0.0, 0.0], [0.0, 1e11], abstol=1e-8, reltol=1e-8, maxstep=1e11 / 10,
                              minstep=1e11 / 1e18)
    ref = np.array([0.2083340149701255e-07, 0.8333360770334713e-13, 0.9999999791665050])
    assert np.max(np.abs(y[-1] - ref)) < 2e-10
    o = O.solve("ode23s", "robertson", [1.0, 0.0, 0.0], [0.0, 1e11], [0.04, 3e7, 1e4], abstol=1e-8, reltol=1e-8,
                maxstep=1e10, minstep=1e-7)
    assert np.array_equal(t, o.t) and np.array_equal(y, o.y)
    assert (st["n_accept"], st["n_reject"], st["n_rhs"]) == (o.n_accept, o.n_reject, o.n_rhs) == (905, 12, 6377)


@pytest.mark.parametrize("method", ADAPTIVE + ["ode23s"])
@pytest.mark.parametrize("points", ["all", "specified"])
def test_single_system_tout_yout_bit_exact(method, points):
    """The (tout, yout) contract incl. Hermite / Rosenbrock dense output at interior tspan points."""
    F = m.DeviceRHS("lorenz", *LORENZ)
    ts = [0.0, 0.3, 0.31, 1.0, 1.7, 2.0]
    t, y, st = getattr(m, method).stats(F, [1.0, 1.0, 1.0], ts, points=points)
    o = O.solve(method, "lorenz", [1.0, 1.0, 1.0], ts, LORENZ, points=points)
    assert np.array_equal(t, o.t) and np.array_equal(y, o.y)
    assert (st["n_accept"], st["n_reject"], st["n_rhs"]) == (o.n_accept, o.n_reject, o.n_rhs)
    if points == "specified" and method != "ode23s":
        assert np.array_equal(t, ts)


@pytest.mark.parametrize("method", ADAPTIVE + ["ode23s"])
def test_attempt_trace_bit_exact(method):
    """Accept/reject sequence, dt and err of every attempt of one trajectory."""
    y0 = lorenz_ics(8, 3)
    tr = 5
    got = m.solve_ensemble(method, m.DeviceRHS("lorenz", *LORENZ), y0, [0.0, 3.0], trace_traj=tr)
    o = O.solve(method, "lorenz", y0[tr], [0.0, 3.0], LORENZ, trace=True)
    assert got["trace_n"] == len(o.trace)
    assert np.array_equal(got["trace"], o.trace)
    assert o.n_reject > 0 or method in ("ode21", "ode23s", "ode45_ck")  # (no rejection on this trajectory for those)


@pytest.mark.parametrize("method", FIXED)
def test_fixed_step_bit_exact(method):
    F = m.DeviceRHS("lorenz", *LORENZ)
    ts = np.linspace(0.0, 1.0, 201)
    t, y = getattr(m, method)(F, [1.0, 1.0, 1.0], ts)
    o = O.solve(method, "lorenz", [1.0, 1.0, 1.0], ts, LORENZ)
    assert np.array_equal(t, o.t) and np.array_equal(y, o.y)
    y0 = lorenz_ics(300, 4)
    got = m.solve_ensemble(method, F, y0, ts)
    for i in (0, 7, 299):
        oi = O.solve(method, "lorenz", y0[i], ts, LORENZ)
        assert np.array_equal(got["y_final"][i], oi.y[-1])


MULTISTEP = ["ode4ms", "ode5ms", "ode_ms1", "ode_ms2", "ode_ms3"]  # ode_ms, src/ODE.jl:175-213
ALL = FIXED + MULTISTEP + ADAPTIVE + ["ode4s_s", "ode4s_kr", "ode4s", "ode23s"]


@pytest.mark.parametrize("solver", ALL)
def test_reference_runtests_analytic(solver):
    """test/runtests.jl:40-70 (T = Float64, tol 1e-2) through the odeXX front-ends."""
    tol = 1e-2
    f = getattr(m, solver)
    t, y = f(m.DeviceRHS("const", 6.0), 0.0, np.arange(0, 11) * 0.1)
    assert np.max(np.abs(y - 6 * t)) < tol and t.dtype == np.float64
    ts = np.arange(0, 1001) * 0.001
    t, y = f(m.DeviceRHS("timelin", 2.0), 0.0, ts)
    assert np.max(np.abs(y - t ** 2)) < tol
    t, y = f(m.DeviceRHS("linear", 1.0), 1.0, ts)
    assert np.max(np.abs(y - np.exp(t))) < tol
    if solver != "ode23s":
        t, y = f(m.DeviceRHS("linear", 1.0), 1.0, ts[::-1].copy())
        assert np.max(np.abs(y - np.exp(t - 1))) < tol
    ts = np.arange(0, int(2 * math.pi / 0.001) + 1) * 0.001
    t, y = f(m.DeviceRHS("rotation"), [1.0, 2.0], ts)
    ex = np.stack([np.cos(t) - 2 * np.sin(t), 2 * np.cos(t) + np.sin(t)], 1)
    assert np.max(np.abs(y - ex)) < tol


def test_ode23s_negative_start():
    """test/runtests.jl:75."""
    t, y = m.ode23s(m.DeviceRHS("rotation"), [1.0, 2.0], [-5.0, 0.0])
    assert len(t) > 1


def test_readme_solve_common_interface(capsys):
    """README.md:17-23 / test/common.jl: solve(prob, ode45(), reltol=1e-8, abstol=1e-8)."""
    prob = m.ODEProblem(m.DeviceRHS("linear", 1.01), 0.5, (0.0, 1.0))
    sol = m.solve(prob, m.ode45(), reltol=1e-8, abstol=1e-8)
    assert sol.retcode == "Succss" and len(sol.t) == 12 and sol.t[-1] == 1.0
    assert sol[-1] == 1.3728005108017367
    assert abs(sol[-1] - 0.5 * math.exp(1.01)) < 5e-9
    assert isinstance(float(sol[1]), float)
    # 2-D array problem keeps its shape (test/common.jl:21-23)
    prob2 = m.ODEProblem(m.DeviceRHS("rotation"), np.array([[1.0], [2.0]]), (0.0, 1.0))
    for alg in (m.ode23(), m.ode45(), m.ode78()):
        s2 = m.solve(prob2, alg, dt=1 / 16, dtmin=np.finfo(float).eps, abstol=1e-6, reltol=1e-3)
        assert s2[1].shape == (2, 1)
    # saveat, save_everystep=false -> points=:specified
    s3 = m.solve(prob, m.ode45(), saveat=[0.25, 0.5], reltol=1e-8, abstol=1e-8)
    assert list(s3.t) == [0.25, 0.5, 1.0] or list(s3.t) == [0.0, 0.25, 0.5, 1.0]


def test_reference_error_behaviour(capsys):
    F = m.DeviceRHS("linear", 1.0)
    with pytest.raises(ValueError, match="Zero time span"):
        m.ode45(F, 1.0, [0.0, 0.0])
    with pytest.raises(ValueError, match="initstep has wrong sign"):
        m.ode45(F, 1.0, [0.0, 1.0], initstep=-0.1)
    with pytest.raises(ValueError, match="Unrecognized option points"):
        m.ode45(F, 1.0, [0.0, 1.0], points="some")
    # dt < minstep on the first rejection: warning + truncated solution (src/runge_kutta.jl:358-360)
    t, y = m.ode45(m.DeviceRHS("lorenz", *LORENZ), [1.0, 1.0, 1.0], [0.0, 10.0], minstep=1e10)
    assert "Warning: dt < minstep.  Stopping." in capsys.readouterr().out
    o = O.solve("ode45", "lorenz", [1.0, 1.0, 1.0], [0.0, 10.0], LORENZ, minstep=1e10)
    assert t[-1] < 10.0 and np.array_equal(t, o.t) and np.array_equal(y, o.y)
    # max_steps guard instead of the reference's endless loop
    got = m.solve_ensemble("ode45", m.DeviceRHS("lorenz", *LORENZ), lorenz_ics(4), [0.0, 10.0], max_steps=50)
    assert np.all(got["status"] == 2) and np.all(got["n_accept"] + got["n_reject"] == 50)


def test_empty_and_ragged_ensembles():
    F = m.DeviceRHS("lorenz", *LORENZ)
    got = m.solve_ensemble("ode45", F, np.zeros((0, 3)), [0.0, 1.0])
    assert got["y_final"].shape == (0, 3)
    for n in (1, 31, 33, 129, 1000):
        y0 = lorenz_ics(n, n)
        got = m.solve_ensemble("ode45", F, y0, [0.0, 0.5])
        ref = O.solve_ensemble("ode45", "lorenz", y0, [0.0, 0.5], LORENZ)
        assert_same(got, ref)


def test_chunked_host_path_matches_oracle_at_chunk_boundaries(monkeypatch):
    """ODE_B200_CHUNKS=4 turns on the two-stream chunk pipeline inside ode_b200_solve_ensemble; results must
    not depend on the chunking (checked against the oracle around every chunk boundary)."""
    monkeypatch.setenv("ODE_B200_CHUNKS", "4")
    n = 4 * 16384 + 17
    y0 = lorenz_ics(n, 77)
    got = m.solve_ensemble("ode45", m.DeviceRHS("lorenz", *LORENZ), y0, [0.0, 1.0])
    assert np.all(got["status"] == 0) and np.all(got["t_final"] == 1.0)
    idx = np.unique(np.concatenate([np.arange(0, 64)] + [np.arange(n * c // 4 - 32, n * c // 4 + 32) for c in (1, 2, 3)]
                                   + [np.arange(n - 64, n)]))
    ref = O.solve_ensemble("ode45", "lorenz", y0[idx], [0.0, 1.0], LORENZ)
    for k in ("y_final", "t_final", "n_accept", "n_reject", "n_rhs", "status"):
        assert np.array_equal(got[k][idx], ref[k]), k


@pytest.mark.parametrize("method", ["ode4s_s", "ode4s_kr", "ode4ms"])
@pytest.mark.parametrize("jacobian", [False, True])
def test_rosenbrock_fixed_and_adams_bashforth_bit_exact(method, jacobian):
    """oderosenbrock (src/ODE.jl:366-438) and ode4ms (:175-212) on the tspan grid, all points."""
    if method == "ode4ms" and jacobian:
        pytest.skip("no Jacobian in Adams-Bashforth")
    ts = np.linspace(0.0, 1.0, 101)
    kw = dict(jacobian=True) if jacobian else {}
    for rhs, y0, p in (("lorenz", [1.0, 1.0, 1.0], LORENZ), ("kepler", [0.4, 0.0, 0.0, 2.0], []), ("vdp", [2.0, 0.0], [5.0])):
        t, y, st = getattr(m, method).stats(m.DeviceRHS(rhs, *p), y0, ts, **kw)
        o = O.solve(method, rhs, y0, ts, p if p else None, jacobian=int(jacobian))
        assert np.array_equal(t, o.t) and np.array_equal(y, o.y) and st["n_rhs"] == o.n_rhs
    n = 300
    y0 = lorenz_ics(n, 8)
    got = m.solve_ensemble(method, m.DeviceRHS("lorenz", *LORENZ), y0, ts, **kw)
    ref = O.solve_ensemble(method, "lorenz", y0, ts, LORENZ, jacobian=int(jacobian))
    assert_same(got, ref)


def test_ode23s_analytic_jacobian_bit_exact():
    rng = np.random.default_rng(4)
    n = 256
    k = np.array([0.04, 3e7, 1e4]) * 10 ** rng.uniform(-0.5, 0.5, (n, 3))
    y0 = np.tile([1.0, 0.0, 0.0], (n, 1))
    kw = dict(reltol=1e-8, abstol=1e-8)
    got = m.solve_ensemble("ode23s", m.DeviceRHS("robertson", 0.04, 3e7, 1e4), y0, [0.0, 1e4], params=k, jacobian=True, **kw)
    ref = O.solve_ensemble("ode23s", "robertson", y0, [0.0, 1e4], k, jacobian=1, **kw)
    assert_same(got, ref)
    fd = m.solve_ensemble("ode23s", m.DeviceRHS("robertson", 0.04, 3e7, 1e4), y0, [0.0, 1e4], params=k, **kw)
    assert np.all(got["n_rhs"] < fd["n_rhs"])  # no dof+1 RHS calls per accepted step
    t, y = m.ode23s(m.DeviceRHS("vdp", 10.0), [2.0, 0.0], [0.0, 1.0, 3.0], jacobian=True)
    o = O.solve("ode23s", "vdp", [2.0, 0.0], [0.0, 1.0, 3.0], [10.0], jacobian=1)
    assert np.array_equal(t, o.t) and np.array_equal(y, o.y)


def test_solve_common_interface_ode4s_ode4ms():
    """test/common.jl:8: algs = [ode23(), ode45(), ode78(), ode4(), ode4ms(), ode4s()]."""
    prob = m.ODEProblem(m.DeviceRHS("linear", 1.01), 0.5, (0.0, 1.0))
    for alg in (m.ode4(), m.ode4ms(), m.ode4s()):
        sol = m.solve(prob, alg, dt=1 / 16, abstol=1e-6, reltol=1e-3)
        assert sol.retcode == "Succss" and isinstance(float(sol[1]), float)


def test_nan_inf_and_blowup_paths_match_oracle(capsys):
    """isoutofdomain / NaN guard (src/runge_kutta.jl:432), minstep stop (:358-360) and the non-finite-dt guard:
    trajectories with NaN / Inf / overflowing states inside an otherwise healthy ensemble."""
    y0 = lorenz_ics(64, 21)
    y0[3, 1] = np.nan
    y0[10, 0] = np.inf
    y0[20] = [1e200, -1e200, 1e200]
    F = m.DeviceRHS("lorenz", *LORENZ)
    for method in ("ode45", "ode23", "ode78"):
        got = m.solve_ensemble(method, F, y0, [0.0, 1.0], max_steps=5000)
        ref = O.solve_ensemble(method, "lorenz", y0, [0.0, 1.0], LORENZ, max_steps=5000)
        assert_same(got, ref, keys=("n_accept", "n_reject", "n_rhs", "status", "t_final"))
        ok = got["status"] == 0
        assert np.array_equal(got["y_final"][ok], ref["y_final"][ok]) and ok.sum() >= 61
        assert set(got["status"][[3, 10, 20]]) <= {1, 2, 3}
    # exponential growth to overflow: y' = 1e3 y on [0, 10]
    got = m.solve_ensemble("ode45", m.DeviceRHS("linear", 1e3), np.array([[1.0], [2.0]]), [0.0, 10.0], max_steps=100000)
    ref = O.solve_ensemble("ode45", "linear", np.array([[1.0], [2.0]]), [0.0, 10.0], [1e3], max_steps=100000)
    assert_same(got, ref, keys=("n_accept", "n_reject", "status", "t_final"))
    assert np.array_equal(got["y_final"], ref["y_final"], equal_nan=True)
    # ode23s with a NaN start: the loop guard `minstep < abs(h)` is false for NaN h (src/ODE.jl:296)
    got = m.solve_ensemble("ode23s", F, y0[:8], [0.0, 0.5], max_steps=2000)
    ref = O.solve_ensemble("ode23s", "lorenz", y0[:8], [0.0, 0.5], LORENZ, max_steps=2000)
    assert_same(got, ref, keys=("n_accept", "n_reject", "n_rhs", "status"))


def test_many_output_points_per_trajectory():
    """points=:specified with thousands of save points per trajectory."""
    ts = np.linspace(0.0, 2.0, 2001)
    y0 = lorenz_ics(16, 2)
    got = m.solve_ensemble("ode45", m.DeviceRHS("lorenz", *LORENZ), y0, ts, points="specified")
    assert got["pts_y"].shape == (16, 2001, 3) and np.all(got["pts_n"] == 2001)
    for i in (0, 7, 15):
        o = O.solve("ode45", "lorenz", y0[i], ts, LORENZ, points="specified")
        assert np.array_equal(got["pts_y"][i], o.y) and np.array_equal(got["pts_t"][i], o.t)
    got = m.solve_ensemble("ode45", m.DeviceRHS("lorenz", *LORENZ), y0, ts, points="all", pts_cap=4096)
    for i in (3, 12):
        o = O.solve("ode45", "lorenz", y0[i], ts, LORENZ, points="all")
        k = got["pts_n"][i]
        assert k == len(o.t) and np.array_equal(got["pts_y"][i, :k], o.y) and np.array_equal(got["pts_t"][i, :k], o.t)
    small = m.solve_ensemble("ode45", m.DeviceRHS("lorenz", *LORENZ), y0, ts, points="all", pts_cap=100)
    assert np.all(small["status"] == 5) and np.array_equal(small["pts_n"], got["pts_n"])  # overflow reports the need


def test_full_size_lorenz_ensemble_properties():
    """BASELINE config 2 at full size (2^20 trajectories): size-independent checks.
    (a) a random sample of 2048 trajectories equals the oracle bit for bit;
    (b) permuting the ensemble permutes the results (trajectories are independent of the dynamic scheduling);
    (c) every trajectory ends exactly at tend with status OK."""
    n = 1 << 20
    y0 = lorenz_ics(n)
    F = m.DeviceRHS("lorenz", *LORENZ)
    got = m.solve_ensemble("ode45", F, y0, [0.0, 10.0])
    assert np.all(got["status"] == 0) and np.all(got["t_final"] == 10.0)
    rng = np.random.default_rng(1)
    idx = rng.choice(n, 2048, replace=False)
    ref = O.solve_ensemble("ode45", "lorenz", y0[idx], [0.0, 10.0], LORENZ)
    for k in ("y_final", "n_accept", "n_reject", "n_rhs"):
        assert np.array_equal(got[k][idx], ref[k]), k
    perm = rng.permutation(n)
    got2 = m.solve_ensemble("ode45", F, y0[perm], [0.0, 10.0])
    for k in ("y_final", "n_accept", "n_reject"):
        assert np.array_equal(got2[k], got[k][perm]), k
    att = int(got["n_accept"].sum() + got["n_reject"].sum())
    assert att == 381568754  # the bench's workload: attempts per 2^20-trajectory pass (seed 20261019)


def test_full_size_kepler_and_robertson_invariants():
    """BASELINE configs 3 and 4 at full size (2^18): conserved quantities."""
    rng = np.random.default_rng(20261019)
    n = 1 << 18
    e = rng.uniform(0, 0.9, n)
    y0 = np.stack([1 - e, np.zeros(n), np.zeros(n), np.sqrt((1 + e) / (1 - e))], 1)
    got = m.solve_ensemble("ode78", m.DeviceRHS("kepler"), y0, [0.0, 20 * math.pi], reltol=1e-12, abstol=1e-12)
    assert np.all(got["status"] == 0)
    yf = got["y_final"]
    en = lambda y: 0.5 * (y[:, 2] ** 2 + y[:, 3] ** 2) - 1 / np.hypot(y[:, 0], y[:, 1])
    am = lambda y: y[:, 0] * y[:, 3] - y[:, 1] * y[:, 2]
    assert np.max(np.abs(en(yf) - en(y0))) < 1e-9
    assert np.max(np.abs(am(yf) - am(y0))) < 1e-9
    assert np.max(np.hypot(yf[:, 0] - y0[:, 0], yf[:, 1] - y0[:, 1])) < 2e-7  # 10 closed orbits
    k = np.array([0.04, 3e7, 1e4]) * 10 ** rng.uniform(-0.5, 0.5, (n, 3))
    y0 = np.tile([1.0, 0.0, 0.0], (n, 1))
    got = m.solve_ensemble("ode23s", m.DeviceRHS("robertson", 0.04, 3e7, 1e4), y0, [0.0, 1e5], params=k,
                           reltol=1e-8, abstol=1e-8)
    assert np.all(got["status"] == 0) and np.all(got["t_final"] == 1e5)
    assert np.max(np.abs(got["y_final"].sum(1) - 1.0)) < 1e-7 and got["y_final"].min() > -1e-9


@pytest.mark.parametrize("method", FIXED + MULTISTEP + ["ode4s_s"])
def test_fixed_step_blowup_matches_oracle(method):
    """Fixed-step solvers on a trajectory that overflows: Inf/NaN pattern as in the reference (every tableau
    coefficient is multiplied, so Inf*0 = NaN)."""
    ts = np.linspace(0.0, 2.0, 41)
    y0 = np.array([[1e150, -1e150, 1e150], [1.0, 1.0, 1.0]])
    got = m.solve_ensemble(method, m.DeviceRHS("lorenz", *LORENZ), y0, ts, points="specified")
    for i in range(2):
        o = O.solve(method, "lorenz", y0[i], ts, LORENZ)
        k = len(o.t)  # the Rosenbrock solvers stop at a singular W (Julia raises SingularException there)
        assert got["pts_n"][i] == k and got["status"][i] == o.status, (method, i)
        assert np.array_equal(got["pts_y"][i, :k], o.y, equal_nan=True), (method, i)
    if method != "ode4s_s":  # the L-stable Rosenbrock step survives 1e150; the explicit ones overflow
        assert not np.all(np.isfinite(got["pts_y"][0]))


@pytest.mark.gpu
@pytest.mark.parametrize("method", ["ode21", "ode23", "ode45_fe", "ode45", "ode78", "ode45_ck"])
def test_hinit_prepass_equals_inline_hinit(method):
    """Ensembles of >= 4096 trajectories take the initial step and f0 from ens_adapt_init_kernel; smaller
    ones compute hinit inside the persistent kernel.  Same bits either way (forward and backward spans,
    per-trajectory parameters), and the same as the oracle on a few trajectories."""
    rng = np.random.default_rng(77)
    n = 6000
    y0 = np.stack([rng.uniform(-10, 10, n), rng.uniform(-10, 10, n), rng.uniform(5, 35, n)], 1)
    par = np.stack([rng.uniform(8, 12, n), rng.uniform(20, 30, n), rng.uniform(2, 3, n)], 1)
    F = m.DeviceRHS("lorenz", *LORENZ)
    # backward spans can run away after a rejection (reference quirk): bound the attempts
    for ts in ([0.0, 0.7], [0.3, -0.2]):
        kw = dict(params=par, max_steps=400)
        big = m.solve_ensemble(method, F, y0, ts, **kw)
        halves = [m.solve_ensemble(method, F, y0[s], ts, params=par[s], max_steps=400)
                  for s in (slice(0, 3000), slice(3000, n))]
        for k in ("y_final", "t_final", "n_accept", "n_reject", "n_rhs", "status"):
            assert np.array_equal(big[k], np.concatenate([h[k] for h in halves])), (k, ts)
        for i in (0, 2999, 5999):
            o = O.solve(method, "lorenz", y0[i], ts, list(par[i]), points="final", max_steps=400)
            assert np.array_equal(big["y_final"][i], o.y[-1]) and big["n_rhs"][i] == o.n_rhs and big["status"][i] == o.status


@pytest.mark.gpu
def test_pinned_output_buffers_are_written_in_place():
    """Outputs in page-locked memory (ode_b200_host_alloc) are written by the kernel itself; pageable ones are
    staged and copied.  Same bits, and `out=` really reuses the arrays."""
    rng = np.random.default_rng(5)
    n = 5000
    y0 = np.stack([rng.uniform(-10, 10, n), rng.uniform(-10, 10, n), rng.uniform(5, 35, n)], 1)
    F = m.DeviceRHS("lorenz", *LORENZ)
    ref = m.solve_ensemble("ode45", F, y0, [0.0, 1.0])
    out = dict(y_final=m.pinned_empty((n, 3)), t_final=m.pinned_empty((n,)), n_accept=m.pinned_empty((n,), np.int32),
               n_rhs=m.pinned_empty((n,), np.int32))
    out["y_final"][:] = -1.0
    got = m.solve_ensemble("ode45", F, y0, [0.0, 1.0], out=out, direct_output=True)
    assert got["y_final"] is out["y_final"] and got["n_accept"] is out["n_accept"]
    for k in ("y_final", "t_final", "n_accept", "n_reject", "n_rhs", "status"):
        assert np.array_equal(got[k], ref[k]), k
    # a blow-up goes through the literal second pass, which also writes the mapped buffers
    y0[7] = [1e200, -1e200, 1e200]
    ref = m.solve_ensemble("ode45", F, y0, [0.0, 1.0])
    got = m.solve_ensemble("ode45", F, y0, [0.0, 1.0], out=out, direct_output=True)
    for k in ("y_final", "t_final", "n_accept", "n_reject", "n_rhs", "status"):
        assert np.array_equal(got[k], ref[k], equal_nan=True), k
    got = m.solve_ensemble("ode45", F, y0, [0.0, 1.0], out=out)  # same buffers, staged copy
    for k in ("y_final", "t_final", "n_accept", "n_reject", "n_rhs", "status"):
        assert np.array_equal(got[k], ref[k], equal_nan=True), k
    with pytest.raises(ValueError):
        m.solve_ensemble("ode45", F, y0, [0.0, 1.0], out=dict(y_final=np.empty((n, 2))))


@pytest.mark.gpu
def test_shared_denominator_division_is_ieee_division():
    """recip_prepare/div_by (csrc/common.cuh) against the compiler's a / b on 2^30 operand pairs."""
    assert m.lib().ode_b200_selftest_div(0, 1 << 30, 20261019) == 0


@pytest.mark.gpu
def test_cuda_path_against_committed_golden_runs():
    """The CUDA path against tests/golden/oracle_runs.json directly (no live oracle in the loop): step counts and the
    final state, as exact hex floats, of every committed run."""
    import json
    import os
    gold = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "oracle_runs.json")))
    for entry in gold:
        c, want = entry["case"], entry["result"]
        F = m.DeviceRHS(c["rhs"], *(c["params"] or []))
        kw = dict(c["kw"])
        if "jacobian" in kw:
            kw["jacobian"] = True
        y0 = c["y0"] if len(c["y0"]) > 1 else c["y0"][0]
        fixed = c["method"] in ("ode1", "ode2_midpoint", "ode2_heun", "ode4", "ode4s_s", "ode4s_kr", "ode4ms")
        if fixed and (len(c["y0"]) > 1 or c["rhs"] == "reaction_diffusion"):
            kw = {k: v for k, v in kw.items() if k == "jacobian"}
        t, y, st = getattr(m, c["method"]).stats(F, y0, c["tspan"], **kw)
        got = dict(n_accept=int(st["n_accept"]), n_reject=int(st["n_reject"]), n_rhs=int(st["n_rhs"]), status=int(st["status"]),
                   n_out=len(t), t_final=float(t[-1]).hex(), y_final=[float(v).hex() for v in np.atleast_1d(y[-1])])
        assert got == want, c["name"]


def test_plain_c_program_solves_the_readme_problem(tmp_path):
    """tests/abi_smoke.c, built with gcc against include/ode_b200.h: the README problem (config 1) through
    ode_b200_solve_single from plain C -- 11 accepted / 0 rejected steps, y(1) = 1.3728005108017367."""
    import subprocess
    from test_abi import build_abi_smoke
    r = subprocess.run([build_abi_smoke(tmp_path)], capture_output=True, text=True)
    assert r.returncode == 0, r.stdout + r.stderr
    assert "11 accepted, 0 rejected" in r.stdout


@pytest.mark.parametrize("order", [1, 2, 3, 4, 5])
def test_ode_ms_all_orders_bit_exact(order):
    """ode_ms(F, x0, tspan, order) (src/ODE.jl:175-209; ode4ms :212, ode5ms :213): every tspan point of a single
    system and of an ensemble with per-trajectory parameters equals the oracle's, bit for bit."""
    ts = np.linspace(0.0, 1.0, 201)
    t, y = m.ode_ms(m.DeviceRHS("lorenz", 10.0, 28.0, 8 / 3), [1.0, 1.0, 1.0], ts, order)
    ref = O.solve("ode_ms%d" % order, "lorenz", [1.0, 1.0, 1.0], ts, [10.0, 28.0, 8 / 3])
    assert np.array_equal(t, ref.t) and np.array_equal(y, ref.y)
    if order == 5:
        t5, y5 = m.ode5ms(m.DeviceRHS("lorenz", 10.0, 28.0, 8 / 3), [1.0, 1.0, 1.0], ts)
        assert np.array_equal(y5, ref.y)
    rng = np.random.default_rng(order)
    n = 777
    y0 = rng.uniform(-2, 2, (n, 2))
    mu = rng.uniform(0.5, 3.0, (n, 1))
    got = m.solve_ensemble("ode_ms%d" % order, m.DeviceRHS("vdp", 1.0), y0, ts, params=mu, points="all", pts_cap=len(ts))
    exp = O.solve_ensemble("ode_ms%d" % order, "vdp", y0, ts, mu)
    assert np.array_equal(got["y_final"], exp["y_final"]) and np.all(got["pts_n"] == len(ts))
    k = int(rng.integers(n))
    one = O.solve("ode_ms%d" % order, "vdp", y0[k], ts, mu[k])
    assert np.array_equal(got["pts_y"][k], one.y)
    with pytest.raises(ValueError):
        m.ode_ms(m.DeviceRHS("vdp", 1.0), [1.0, 0.0], ts, 6)


def test_relaxed_fp_mode_is_close_to_but_is_not_the_parity_path():
    """opts.fp_mode = ODE_B200_FP_FMA (strict_fp=False): the Dormand-Prince kernels compiled with FMA contraction.
    Opt-in, never the parity path: on non-chaotic problems (and Lorenz over a short horizon) the final states agree
    with the strict / oracle result to 1e-10 relative (the north-star tolerance) and the step counts are the same
    except where a decision sat within rounding of the threshold; the bits are NOT the oracle's."""
    rng = np.random.default_rng(11)
    n = 8192
    cases = [("vdp", m.DeviceRHS("vdp", 1.0), rng.uniform(-2, 2, (n, 2)), [0.0, 5.0]),
             ("lorenz", m.DeviceRHS("lorenz", 10.0, 28.0, 8 / 3),
              np.stack([rng.uniform(-10, 10, n), rng.uniform(-10, 10, n), rng.uniform(5, 35, n)], 1), [0.0, 2.0])]
    for name, F, y0, ts in cases:
        strict = m.solve_ensemble("ode45", F, y0, ts)
        fma = m.solve_ensemble("ode45", F, y0, ts, strict_fp=False)
        assert np.all(fma["status"] == 0)
        scale = np.maximum(np.abs(strict["y_final"]).max(axis=1), 1e-3)
        rel = np.abs(fma["y_final"] - strict["y_final"]).max(axis=1) / scale
        same_steps = (fma["n_accept"] == strict["n_accept"]) & (fma["n_reject"] == strict["n_reject"])
        print("%s: max rel diff %.3e over %d trajectories, %d with different step counts, %d bit-identical"
              % (name, rel[same_steps].max(), n, int((~same_steps).sum()), int((rel == 0).sum())))
        assert rel[same_steps].max() < 1e-10
        assert (~same_steps).sum() <= n // 1000  # a flipped accept/reject needs err within ~1e-15 of 1
        assert (rel > 0).any()  # it really is different arithmetic
    with pytest.raises(m.OdeB200Error):  # no FMA build of the other methods: refused, not silently strict
        m.solve_ensemble("ode23", cases[0][1], cases[0][2], cases[0][3], strict_fp=False)
    with pytest.raises(m.OdeB200Error):
        m.solve_large("ode45", m.DeviceRHS("reaction_diffusion", 1.0, 1.0), np.linspace(0.1, 0.9, 4096), [0.0, 1.0], strict_fp=False)


@pytest.mark.gpu
@pytest.mark.parametrize("method,rhs", [("ode45", "lorenz"), ("ode45", "lorenz_pp"), ("ode78", "kepler"), ("ode23", "vdp")])
def test_tail_handoff_second_launch_bit_exact(method, rhs, monkeypatch):
    """Ensembles that keep every lane busy hand the tail over to a second, re-packed launch of the same kernel
    (ensemble_adapt.cuh, ODE_B200_DRAIN): a trajectory's state crosses the launches as the same doubles, so the results
    must not move by a bit -- against the run without the hand-off, and against the oracle."""
    rng = np.random.default_rng(5)
    n = 160000  # > the ~95k resident lanes of the persistent grid
    par = None
    if rhs.startswith("lorenz"):
        y0 = np.stack([rng.uniform(-10, 10, n), rng.uniform(-10, 10, n), rng.uniform(5, 35, n)], 1)
        F, ts, kw, oname, op = m.DeviceRHS("lorenz", *LORENZ), [0.0, 0.6], {}, "lorenz", LORENZ
        if rhs == "lorenz_pp":  # per-trajectory parameters: the resumed lane reloads them
            par = np.stack([rng.uniform(8, 12, n), rng.uniform(20, 30, n), rng.uniform(2, 3, n)], 1)
            kw = dict(params=par)
    elif rhs == "kepler":
        e = rng.uniform(0, 0.9, n)
        y0 = np.stack([1 - e, np.zeros(n), np.zeros(n), np.sqrt((1 + e) / (1 - e))], 1)
        F, ts, kw, oname, op = m.DeviceRHS("kepler"), [0.0, 2.0], dict(reltol=1e-10, abstol=1e-10), "kepler", None
    else:
        y0 = np.stack([rng.uniform(-2, 2, n), rng.uniform(-2, 2, n)], 1)
        F, ts, kw, oname, op = m.DeviceRHS("vdp", 5.0), [0.0, 1.0], {}, "vdp", [5.0]
    L = m.lib()
    monkeypatch.setenv("ODE_B200_DRAIN", "0")
    L.ode_b200_launch_count(1)
    plain = m.solve_ensemble(method, F, y0, ts, **kw)
    l0 = L.ode_b200_launch_count(1)
    monkeypatch.setenv("ODE_B200_DRAIN", "1")
    got = m.solve_ensemble(method, F, y0, ts, **kw)
    l1 = L.ode_b200_launch_count(1)
    if l1 == l0:
        pytest.skip("library built without ODEB200_DRAIN (the default: the hand-off measured slower, common.cuh)")
    assert l1 == l0 + 5, (l0, l1)  # 4 sort kernels + the second launch
    assert_same(got, plain)
    assert int(got["status"].max()) == 0
    O.use_all_cores()
    sub = slice(0, 20000)
    ref = O.solve_ensemble(method, oname, y0[sub], ts, op if par is None else par[sub],
                           **{k: v for k, v in kw.items() if k != "params"})
    for k in ("y_final", "t_final", "n_accept", "n_reject", "n_rhs", "status"):
        assert np.array_equal(got[k][sub], ref[k]), k


@pytest.mark.gpu
@pytest.mark.parametrize("method,rhs,n", [("ode45", "lorenz", 200000), ("ode45", "lorenz", 6000), ("ode78", "lorenz", 70001),
                                          ("ode23", "lorenz", 9999), ("ode78", "kepler", 30011), ("ode45", "vdp", 20000),
                                          ("ode45_ck", "linear", 5001)])
def test_streaming_prepass_from_pinned_y0_bit_exact(method, rhs, n, monkeypatch):
    """y0 in page-locked host memory + one parameter vector: the pre-pass reads y0 over PCIe chunk by chunk while the
    persistent kernel is already integrating (ens_adapt_stream_init_kernel) instead of an upload in front of
    everything.  Same bits as the staged path and as the oracle; ragged sizes exercise the last partial chunk, the
    functors state sizes 1..4 (a chunk of y0 is fetched in 16-byte pieces)."""
    rng = np.random.default_rng(11)
    kw = {}
    if rhs == "lorenz":
        src = np.stack([rng.uniform(-10, 10, n), rng.uniform(-10, 10, n), rng.uniform(5, 35, n)], 1)
        F, par, ts = m.DeviceRHS("lorenz", *LORENZ), LORENZ, [0.0, 0.5]
    elif rhs == "kepler":
        e = rng.uniform(0, 0.9, n)
        src = np.stack([1 - e, np.zeros(n), np.zeros(n), np.sqrt((1 + e) / (1 - e))], 1)
        F, par, ts, kw = m.DeviceRHS("kepler"), None, [0.0, 1.5], dict(reltol=1e-10, abstol=1e-10)
    elif rhs == "vdp":
        src = np.stack([rng.uniform(-2, 2, n), rng.uniform(-2, 2, n)], 1)
        F, par, ts = m.DeviceRHS("vdp", 5.0), [5.0], [0.0, 1.0]
    else:
        src = rng.uniform(0.5, 2.0, (n, 1))
        F, par, ts = m.DeviceRHS("linear", 1.01), [1.01], [0.0, 1.0]
    y0 = m.pinned_empty(src.shape)
    y0[:] = src
    monkeypatch.setenv("ODE_B200_STREAM_INIT", "0")
    plain = m.solve_ensemble(method, F, y0, ts, **kw)
    monkeypatch.setenv("ODE_B200_STREAM_INIT", "1")
    for direct in (False, True):
        got = m.solve_ensemble(method, F, y0, ts, direct_output=direct, **kw)
        assert_same(got, plain)
    pageable = m.solve_ensemble(method, F, np.array(y0), ts, **kw)  # a pageable copy takes the staged path
    assert_same(pageable, plain)
    sub = slice(0, min(n, 20000))
    O.use_all_cores()
    ref = O.solve_ensemble(method, rhs, np.array(y0[sub]), ts, par, **kw)
    for k in ("y_final", "t_final", "n_accept", "n_reject", "n_rhs", "status"):
        assert np.array_equal(got[k][sub], ref[k]), k


@pytest.mark.gpu
@pytest.mark.parametrize("pinned_out", [False, True])
def test_progressive_copy_back_bit_exact(pinned_out, monkeypatch):
    """Staged outputs of a large ensemble go back to the host chunk by chunk while the kernel still runs (the lane that
    finishes a chunk's last trajectory raises a host flag, the calling thread issues the copies).  Same bits as the
    copy phase after the kernel, also for chunks completed by the literal pass (a blown-up trajectory) and a ragged
    last chunk."""
    rng = np.random.default_rng(21)
    n = 150001
    y0 = np.stack([rng.uniform(-10, 10, n), rng.uniform(-10, 10, n), rng.uniform(5, 35, n)], 1)
    y0[40000] = [1e200, -1e200, 1e200]  # overflows in the first attempt: re-integrated by the literal pass
    F = m.DeviceRHS("lorenz", *LORENZ)
    ts = [0.0, 0.5]
    keys = ("t_final", "y_final", "n_accept", "n_reject", "status")

    def outputs():
        if not pinned_out:
            return None
        return dict(t_final=m.pinned_empty(n), y_final=m.pinned_empty((n, 3)), n_accept=m.pinned_empty(n, np.int32),
                    n_reject=m.pinned_empty(n, np.int32), status=m.pinned_empty(n, np.int32))

    monkeypatch.setenv("ODE_B200_PROGRESSIVE", "0")
    plain = m.solve_ensemble("ode45", F, y0, ts, max_steps=500, out=outputs())
    monkeypatch.setenv("ODE_B200_PROGRESSIVE", "1")
    got = m.solve_ensemble("ode45", F, y0, ts, max_steps=500, out=outputs())
    for k in keys:
        assert np.array_equal(got[k], plain[k], equal_nan=True), k
    assert got["status"][40000] != 0 and int((got["status"] != 0).sum()) == 1
    O.use_all_cores()
    sub = slice(30000, 50000)
    ref = O.solve_ensemble("ode45", "lorenz", y0[sub], ts, LORENZ, max_steps=500)
    for k in keys:
        assert np.array_equal(got[k][sub], ref[k], equal_nan=True), k


@pytest.mark.gpu
@pytest.mark.parametrize("method,rhs,n", [("ode45", "lorenz", 5001), ("ode45", "lorenz", 100003), ("ode78", "kepler", 4099),
                                          ("ode23s", "robertson", 4097), ("ode23", "vdp", 333)])
def test_device_entry_point_writes_nothing_outside_its_output_arrays(method, rhs, n):
    """Own bounds check of the ensemble kernels (compute-sanitizer is not available on every pool): every output array
    of ode_b200_solve_ensemble_dev sits between two canary margins in one device allocation; after the solve (pre-pass,
    persistent kernel, literal pass) the margins must be intact and the payload complete."""
    import ctypes as C
    import torch
    from ode_jl_b200 import _lib
    dev = torch.device("cuda", 0)
    rng = np.random.default_rng(9)
    if rhs == "lorenz":
        y0 = np.stack([rng.uniform(-10, 10, n), rng.uniform(-10, 10, n), rng.uniform(5, 35, n)], 1)
        F, ts, kw = m.DeviceRHS("lorenz", *LORENZ), [0.0, 0.5], {}
    elif rhs == "kepler":
        e = rng.uniform(0, 0.9, n)
        y0 = np.stack([1 - e, np.zeros(n), np.zeros(n), np.sqrt((1 + e) / (1 - e))], 1)
        F, ts, kw = m.DeviceRHS("kepler"), [0.0, 1.0], dict(reltol=1e-10, abstol=1e-10)
    elif rhs == "robertson":
        y0 = np.tile([1.0, 0.0, 0.0], (n, 1))
        F, ts, kw = m.DeviceRHS("robertson", 0.04, 3e7, 1e4), [0.0, 10.0], dict(reltol=1e-6, abstol=1e-8)
    else:
        y0 = np.stack([rng.uniform(-2, 2, n), rng.uniform(-2, 2, n)], 1)
        F, ts, kw = m.DeviceRHS("vdp", 5.0), [0.0, 1.0], {}
    dof = y0.shape[1]
    M = 1024  # canary elements on either side
    CAN64, CAN32 = float.fromhex("0x1.dead1p+700"), 0x5EADBEEF

    def guarded(count, dtype):
        t = torch.full((count + 2 * M,), CAN64 if dtype == torch.float64 else CAN32, dtype=dtype, device=dev)
        return t, t[M:M + count]

    bufs = {k: guarded(c, dt) for k, c, dt in (("t_final", n, torch.float64), ("y_final", n * dof, torch.float64),
                                               ("n_accept", n, torch.int32), ("n_reject", n, torch.int32),
                                               ("n_rhs", n, torch.int32), ("status", n, torch.int32))}
    d_y0 = torch.from_numpy(np.ascontiguousarray(y0)).to(dev)
    par = F.param_array()
    d_par = torch.from_numpy(np.ascontiguousarray(par if par is not None else np.zeros(1))).to(dev)
    d_ts = torch.tensor(ts, dtype=torch.float64, device=dev)
    opts = m.make_opts(method, F, np.array(ts), points="final", **kw)
    io = _lib.EnsembleIO()
    io.n_traj, io.n_tspan, io.trace_traj = n, 2, -1
    io.y0, io.params, io.tspan = d_y0.data_ptr(), d_par.data_ptr(), d_ts.data_ptr()
    for k in bufs:
        setattr(io, k, bufs[k][1].data_ptr())
    _lib.check(m.lib().ode_b200_solve_ensemble_dev(C.byref(opts), C.byref(io), C.c_void_p(torch.cuda.current_stream().cuda_stream)))
    torch.cuda.synchronize()
    for k, (whole, view) in bufs.items():
        can = CAN64 if whole.dtype == torch.float64 else CAN32
        assert bool((whole[:M] == can).all()) and bool((whole[M + view.numel():] == can).all()), "canary of %s overwritten" % k
        assert not bool((view == can).any()), "%s not completely written" % k
    ref = m.solve_ensemble(method, F, y0, ts, **kw)
    assert np.array_equal(bufs["y_final"][1].cpu().numpy().reshape(n, dof), ref["y_final"])
    assert np.array_equal(bufs["status"][1].cpu().numpy(), ref["status"])
