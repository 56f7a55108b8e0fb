"""Host-side mirror logic that needs no GPU: the `solve` keyword mapping of src/common.jl:50-76, option defaults of
src/runge_kutta.jl:233-239, the registered-RHS tokens and argument errors raised before anything reaches the device."""
import numpy as np
import pytest

import ode_jl_b200 as m
from ode_jl_b200.common import build_ts


def test_ts_from_saveat_like_common_jl():
    # no saveat: Ts = [tspan[1], tspan[2]]
    assert build_ts((0.0, 1.0), ()) == [0.0, 1.0]
    # vector saveat; a trailing tspan[2] is popped and re-appended (:62-64), a leading tspan[1] kept once (:66-67)
    assert build_ts((0.0, 1.0), [0.25, 0.5]) == [0.0, 0.25, 0.5, 1.0]
    assert build_ts((0.0, 1.0), [0.25, 0.5, 1.0]) == [0.0, 0.25, 0.5, 1.0]
    assert build_ts((0.0, 1.0), [0.0, 0.5]) == [0.0, 0.5, 1.0]
    assert build_ts((0.0, 1.0), [0.5, 0.5]) == [0.0, 0.5, 1.0]  # unique
    # number saveat that lands on tspan[2] (:53-54) and one that does not (:55-56)
    assert build_ts((0.0, 1.0), 0.25) == [0.0, 0.25, 0.5, 0.75, 1.0]
    ts = build_ts((0.0, 1.0), 0.1)
    assert len(ts) == 11 and ts[0] == 0.0 and ts[-1] == 1.0 and np.allclose(ts, np.arange(11) * 0.1, atol=1e-15)
    assert build_ts((0.0, 1.0), 0.3) == [0.0, 0.3, 0.6, 1.0]
    assert build_ts((1.0, 0.0), -0.5) == [1.0, 0.5, 0.0]


def test_option_defaults_follow_the_reference():
    F = m.DeviceRHS("lorenz", 10.0, 28.0, 8 / 3)
    o = m.make_opts("ode45", F, [2.0, 12.0])
    assert (o.reltol, o.abstol, o.initstep) == (1e-5, 1e-8, 0.0)
    assert o.minstep == 10.0 / 1e18 and o.maxstep == 10.0 / 2.5  # src/runge_kutta.jl:235-236 (code, not README)
    assert o.points == 0 and o.method == 7 and o.rhs == 4 and o.dof == 3 and o.n_params == 3
    assert m.make_opts("ode23s", F, [0.0, 1.0], jacobian=True).jacobian == 1
    with pytest.raises(TypeError):
        m.make_opts("ode45", F, [0.0, 1.0], jacobian=True)
    with pytest.raises(ValueError, match="Unrecognized option points"):
        m.make_opts("ode45", F, [0.0, 1.0], points="dense")
    with pytest.raises(NotImplementedError):
        m.make_opts("ode45", F, [0.0, 1.0], norm=lambda x: 0.0)
    with pytest.raises(TypeError, match="DeviceRHS"):
        m.make_opts("ode45", lambda t, y: y, [0.0, 1.0])


def test_device_rhs_tokens():
    with pytest.raises(KeyError):
        m.DeviceRHS("pendulum")
    F = m.DeviceRHS("vdp", 5.0)
    assert (F.id, F.dof, F.n_params) == (5, 2, 1) and F.param_array().tolist() == [5.0]
    assert m.DeviceRHS("kepler").param_array() is None
    with pytest.raises(ValueError):
        m.DeviceRHS("lorenz", 1.0).param_array()
    assert F.with_params([7.0]).params == (7.0,) and F.with_params(None) is F
    assert repr(F) == "DeviceRHS('vdp', 5.0)"
    # markers: ode45() is the algorithm object, ode45(F, y0, tspan) the solver call (src/algorithm_types.jl)
    assert m.ode45() is m.ode45 and repr(m.ode45()) == "ode45()"


def test_solve_argument_errors_before_the_device():
    prob = m.ODEProblem(m.DeviceRHS("linear", 1.01), 0.5, (0.0, 1.0))
    with pytest.raises(ValueError, match="mass matrices"):
        m.solve(m.ODEProblem(prob.f, 0.5, (0.0, 1.0), mass_matrix=np.eye(1)), m.ode45())
    with pytest.raises(ValueError, match="callbacks"):
        m.solve(prob, m.ode45(), callback=lambda: None)
    with pytest.raises(TypeError):
        m.solve(prob, "euler")
    with pytest.raises(TypeError):
        m.solve(m.ODEProblem(lambda u, p, t: u, 0.5, (0.0, 1.0)), m.ode45())
    with pytest.raises(TypeError, match="no keyword arguments"):
        m.ode4(m.DeviceRHS("rotation"), [1.0, 2.0], [0.0, 1.0], reltol=1e-3)  # oderk_fixed on vectors
    assert m.shard_bounds(10, 4, 3) == (8, 10)


def test_registered_tokens_evaluate_on_the_host_without_a_gpu():
    """register_rhs / register_stencil / register_field only record source text (compilation happens at first use on
    the device); the tokens carry ids, shapes and an optional host evaluator, like DeviceRHS."""
    import ode_jl_b200 as m
    S = m.register_stencil("cpu_lap4", "return p[0] * (w[0] + w[4]);", host=lambda w, p, t, i, n: p[0] * (w[0] + w[4]),
                           params=(2.0,), radius=2)
    assert S.id >= 2000 and S.is_stencil and S.radius == 2 and S.n_params == 1
    u = np.arange(6.0)
    assert np.array_equal(S(0.0, u), [2.0 * (u[(i - 2) % 6] + u[(i + 2) % 6]) for i in range(6)])
    F = m.register_field("cpu_mirror", "return p[0] * (y[n - 1 - i] - y[i]);", params=(0.5,),
                         host=lambda t, y, p: p[0] * (y[::-1] - y))
    assert F.id >= 3000 and F.is_stencil and F.is_field and F.n_params == 1
    assert np.array_equal(F(0.0, u), 0.5 * (u[::-1] - u))
    R = m.register_rhs("cpu_decay", 1, 1, "f[0] = -p[0] * y[0];", params=(3.0,), host=lambda t, y, p: [-p[0] * y[0]])
    assert 1000 <= R.id < 2000 and R.dof == 1
    assert R(0.0, [2.0]) == [-6.0]
    with pytest.raises(m.OdeB200Error):
        m.register_stencil("cpu_too_wide", "return w[0];", params=(), radius=9)
    with pytest.raises(m.OdeB200Error):
        m.register_field("cpu_too_many", "return 0.0;", params=tuple(range(9)))
