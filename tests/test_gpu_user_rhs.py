"""Right-hand sides registered at run time (NVRTC): the library compiles its own kernel templates for a
functor given as CUDA C++ source.  Checked (a) against the built-in functor of the same expression, bit for
bit, and (b) for a functor that is NOT built in, against the C++ oracle driven by a python callback that
evaluates the same expression in binary64 without FMA."""
import math

import numpy as np
import pytest

import ode_jl_b200 as m
from oracle import oracle as O

pytestmark = pytest.mark.gpu

LORENZ_SRC = """
    f[0] = p[0] * (y[1] - y[0]);
    f[1] = y[0] * (p[1] - y[2]) - y[1];
    f[2] = y[0] * y[1] - p[2] * y[2];
"""
LORENZ_JAC = """
    J[0][0] = -p[0];       J[0][1] = p[0];  J[0][2] = 0.0;
    J[1][0] = p[1] - y[2]; J[1][1] = -1.0;  J[1][2] = -y[0];
    J[2][0] = y[1];        J[2][1] = y[0];  J[2][2] = -p[2];
"""
# Brusselator: not a built-in
BRUSS_SRC = """
    const double x = y[0], z = y[1];
    f[0] = p[0] + (x * x) * z - (p[1] + 1.0) * x;
    f[1] = p[1] * x - (x * x) * z;
"""
BRUSS_JAC = """
    const double x = y[0], z = y[1];
    J[0][0] = (2.0 * x) * z - (p[1] + 1.0);  J[0][1] = x * x;
    J[1][0] = p[1] - (2.0 * x) * z;          J[1][1] = -(x * x);
"""


def bruss(t, y, p):
    x, z = y
    return [p[0] + (x * x) * z - (p[1] + 1.0) * x, p[1] * x - (x * x) * z]


def bruss_jac(t, y, p):
    x, z = y
    return [[(2.0 * x) * z - (p[1] + 1.0), x * x], [p[1] - (2.0 * x) * z, -(x * x)]]


@pytest.fixture(scope="module")
def user_lorenz():
    return m.register_rhs("user_lorenz", 3, 3, LORENZ_SRC, LORENZ_JAC, params=(10.0, 28.0, 8 / 3))


@pytest.fixture(scope="module")
def user_bruss():
    return m.register_rhs("brusselator", 2, 2, BRUSS_SRC, BRUSS_JAC, host=bruss, params=(1.0, 3.0))


@pytest.mark.parametrize("method", ["ode21", "ode23", "ode45_fe", "ode45", "ode78", "ode45_ck", "ode23s", "ode4", "ode1",
                                    "ode4s_s", "ode4s_kr", "ode4ms"])
def test_user_rhs_equals_builtin_functor(method, user_lorenz):
    rng = np.random.default_rng(3)
    y0 = np.stack([rng.uniform(-10, 10, 200), rng.uniform(-10, 10, 200), rng.uniform(5, 35, 200)], 1)
    ts = np.linspace(0.0, 0.5, 26) if m.lib().ode_b200_method_is_adaptive(m._lib.METHODS[method]) == 0 else [0.0, 1.0]
    a = m.solve_ensemble(method, user_lorenz, y0, ts)
    b = m.solve_ensemble(method, m.DeviceRHS("lorenz", 10.0, 28.0, 8 / 3), y0, ts)
    for k in ("y_final", "t_final", "n_accept", "n_reject", "n_rhs", "status"):
        assert np.array_equal(a[k], b[k]), k
    if method in ("ode23s", "ode4s_s"):
        a = m.solve_ensemble(method, user_lorenz, y0, ts, jacobian=True)
        b = m.solve_ensemble(method, m.DeviceRHS("lorenz", 10.0, 28.0, 8 / 3), y0, ts, jacobian=True)
        assert np.array_equal(a["y_final"], b["y_final"]) and np.array_equal(a["n_rhs"], b["n_rhs"])


@pytest.mark.parametrize("method", ["ode23", "ode45", "ode78", "ode23s"])
def test_user_rhs_not_builtin_matches_oracle_callback(method, user_bruss):
    ts = [0.0, 2.5, 10.0]
    for y0 in ([1.5, 3.0], [0.2, 0.1]):
        t, y, st = getattr(m, method).stats(user_bruss, y0, ts)
        o = O.solve_user(method, bruss, y0, ts, [1.0, 3.0])
        assert np.array_equal(t, o.t) and np.array_equal(y, o.y)
        assert (st["n_accept"], st["n_reject"], st["n_rhs"]) == (o.n_accept, o.n_reject, o.n_rhs)
    if method == "ode23s":
        t, y = m.ode23s(user_bruss, [1.5, 3.0], ts, jacobian=True)
        o = O.solve_user("ode23s", bruss, [1.5, 3.0], ts, [1.0, 3.0], jac=bruss_jac)
        assert np.array_equal(t, o.t) and np.array_equal(y, o.y)
    # per-trajectory parameters through the ensemble entry point, final states vs the oracle
    rng = np.random.default_rng(8)
    n = 64
    par = np.stack([rng.uniform(0.5, 1.5, n), rng.uniform(1.0, 3.5, n)], 1)
    y0 = rng.uniform(0.5, 3.0, (n, 2))
    got = m.solve_ensemble(method, user_bruss, y0, [0.0, 5.0], params=par)
    for i in (0, 17, 63):
        o = O.solve_user(method, bruss, y0[i], [0.0, 5.0], list(par[i]), points="final")
        assert np.array_equal(got["y_final"][i], o.y[-1]) and got["n_accept"][i] == o.n_accept
    assert np.allclose(user_bruss(0.0, [1.5, 3.0]), bruss(0.0, [1.5, 3.0], [1.0, 3.0]))  # host evaluator
    if method != "ode23s":  # >= 4096 trajectories: hinit comes from the NVRTC-compiled pre-pass kernel
        n = 5000
        par = np.stack([rng.uniform(0.5, 1.5, n), rng.uniform(1.0, 3.5, n)], 1)
        y0 = rng.uniform(0.5, 3.0, (n, 2))
        big = m.solve_ensemble(method, user_bruss, y0, [0.0, 2.0], params=par)
        small = m.solve_ensemble(method, user_bruss, y0[:100], [0.0, 2.0], params=par[:100])
        for k in ("y_final", "n_accept", "n_reject", "n_rhs"):
            assert np.array_equal(big[k][:100], small[k]), k


def test_user_rhs_errors():
    bad = m.register_rhs("broken", 2, 0, "f[0] = y[1]; f[1] = undefined_symbol;")
    with pytest.raises(m.OdeB200Error, match="NVRTC compilation"):
        m.ode45(bad, [1.0, 0.0], [0.0, 1.0])
    with pytest.raises(m.OdeB200Error):
        m.register_rhs("toolarge", 64, 0, "f[0] = 0.0;")
    nojac = m.register_rhs("nojac", 1, 1, "f[0] = p[0] * y[0];", params=(1.0,))
    with pytest.raises(m.OdeB200Error, match="without a Jacobian"):
        m.ode23s(nojac, [1.0], [0.0, 1.0], jacobian=True)
    t, y = m.ode45(nojac, 1.0, [0.0, 1.0])
    assert abs(y[-1] - np.e) < 1e-4


RD_BODY = "return p0 * ((um - 2.0 * u) + up) + (p1 * u) * (1.0 - u);"
NAGUMO_BODY = "return p0 * ((um - 2.0 * u) + up) + ((p1 * u) * (1.0 - u)) * (u - 0.25);"  # not built in


def nagumo(w, p, t, i, n):
    um, u, up = w
    return p[0] * ((um - 2.0 * u) + up) + ((p[1] * u) * (1.0 - u)) * (u - 0.25)


def test_user_stencil_large_system():
    """Run-time registered 3-point stencils on the fused large-system kernels."""
    n = 3000
    i = np.arange(n)
    u0 = 0.5 + 0.4 * np.sin(2 * np.pi * 3 * i / n)
    rd = m.register_stencil("user_rd", RD_BODY, params=(1.0, 1.0))
    for method in ("ode45", "ode23", "ode78"):
        a = m.solve_large(method, rd, u0, [0.0, 3.0], trace=True)
        b = m.solve_large(method, m.DeviceRHS("reaction_diffusion", 1.0, 1.0), u0, [0.0, 3.0], trace=True)
        assert np.array_equal(a["y"], b["y"]) and np.array_equal(a["trace"], b["trace"])
    ng = m.register_stencil("nagumo", NAGUMO_BODY, host=nagumo, params=(0.8, 2.0))
    n = 600
    u0 = 0.5 + 0.45 * np.sin(2 * np.pi * 2 * np.arange(n) / n)
    ts = [0.0, 0.4, 2.0]
    for method, points in (("ode45", "specified"), ("ode23", "all")):
        t, y, st = getattr(m, method).stats(ng, u0, ts, points=points)
        o = O.solve_user_stencil(method, nagumo, u0, ts, [0.8, 2.0], points=points)
        assert np.array_equal(t, o.t) and np.array_equal(y, o.y)
        assert (st["n_accept"], st["n_reject"], st["n_rhs"]) == (o.n_accept, o.n_reject, o.n_rhs)
    w = u0[:5]
    assert np.allclose(ng(0.0, w), [nagumo([w[k - 1], w[k], w[(k + 1) % 5]], [0.8, 2.0], 0.0, k, 5) for k in range(5)])
    with pytest.raises(m.OdeB200Error, match="NVRTC compilation"):
        m.solve_large("ode45", m.register_stencil("bad", "return undefined_thing;"), u0, [0.0, 1.0])


# 4th-order 5-point Laplacian with a space-dependent diffusivity and a travelling time-dependent source:
# uses the whole stencil interface (radius 2, three parameters, stage time, grid index, ring length)
WIDE_BODY = """
const double x = (double)i / (double)n;
const double kap = p[0] * (1.0 + 0.5 * x);
const double lap = (((-w[0] + 16.0 * w[1]) - 30.0 * w[2]) + 16.0 * w[3]) - w[4];
const double s = x - p[2] * t;
return kap * (lap / 12.0) + p[1] * (s * (1.0 - s));
"""


def wide(w, p, t, i, n):
    x = float(i) / float(n)
    kap = p[0] * (1.0 + 0.5 * x)
    lap = (((-w[0] + 16.0 * w[1]) - 30.0 * w[2]) + 16.0 * w[3]) - w[4]
    s = x - p[2] * t
    return kap * (lap / 12.0) + p[1] * (s * (1.0 - s))


@pytest.mark.gpu
def test_user_stencil_radius2_time_and_index_dependent():
    F = m.register_stencil("wide5", WIDE_BODY, host=wide, params=(0.6, 0.3, 0.1), radius=2)
    n = 700
    u0 = 0.2 + 0.1 * np.cos(2 * np.pi * 3 * np.arange(n) / n)
    ts = [0.0, 0.3, 1.5]
    for method, points in (("ode45", "specified"), ("ode23", "all"), ("ode78", "specified"), ("ode45", "all")):
        t, y, st = getattr(m, method).stats(F, u0, ts, points=points)
        o = O.solve_user_stencil(method, wide, u0, ts, [0.6, 0.3, 0.1], radius=2, points=points)
        assert np.array_equal(t, o.t) and np.array_equal(y, o.y), (method, points)
        assert (st["n_accept"], st["n_reject"], st["n_rhs"]) == (o.n_accept, o.n_reject, o.n_rhs)
    # backward in time, and a radius the tile cannot hold for a 13-stage tableau
    t, y = m.ode45(F, u0, [0.5, 0.0], points="specified")
    o = O.solve_user_stencil("ode45", wide, u0, [0.5, 0.0], [0.6, 0.3, 0.1], radius=2, points="specified")
    assert np.array_equal(y, o.y)
    with pytest.raises(m.OdeB200Error):
        m.register_stencil("toowide", "return w[0];", params=(), radius=5)


@pytest.mark.gpu
@pytest.mark.parametrize("dof", [8, 16])
def test_user_rhs_wider_state_matches_oracle(dof):
    """A ring of `dof` cubic oscillators: the per-thread register design, the in-register LU of the Rosenbrock
    solvers and the fixed-step kernels for a state wider than any built-in functor (register_rhs allows
    dof <= 16).  Bit-exact against the oracle driven by the same expression in Python."""
    D = dof
    src = "\n".join("f[%d] = p[0]*((y[%d] - 2.0*y[%d]) + y[%d]) - p[1]*(y[%d]*y[%d])*y[%d];"
                    % (i, (i - 1) % D, i, (i + 1) % D, i, i, i) for i in range(D))

    def host(t, y, p):
        return [p[0] * ((y[(i - 1) % D] - 2.0 * y[i]) + y[(i + 1) % D]) - p[1] * (y[i] * y[i]) * y[i] for i in range(D)]

    F = m.register_rhs("ring%d" % D, D, 2, src, None, host=host, params=(1.5, 0.3))
    y0 = [math.sin(1.0 + i) for i in range(D)]
    methods = ("ode45", "ode78", "ode23s", "ode4", "ode4s_s", "ode4ms") if D == 8 else ("ode45", "ode23s")
    for meth in methods:
        ts = [0.0, 1.0, 3.0] if meth in ("ode45", "ode78", "ode23s") else list(np.linspace(0, 1, 21))
        t, y, st = getattr(m, meth).stats(F, y0, ts)
        o = O.solve_user(meth, host, y0, ts, [1.5, 0.3])
        assert np.array_equal(t, o.t) and np.array_equal(y, o.y), meth
        assert (st["n_accept"], st["n_rhs"]) == (o.n_accept, o.n_rhs), meth


# Two interleaved species (X at even, Y at odd grid indices) with fixed (Dirichlet) ends, written as ONE radius-2
# stencil that branches on the grid index: the registered-stencil interface covers multi-component and
# non-periodic method-of-lines systems without a special code path (the periodic wrap only feeds points whose
# value the functor ignores).
BRUSS_MOL_BODY = """
const long long cell = i >> 1, ncell = n >> 1;
if (cell == 0 || cell == ncell - 1) return 0.0;           // boundary cells keep their value
const double lap = (w[0] - 2.0 * w[2]) + w[4];            // same species two indices away
if ((i & 1) == 0) {                                       // X: A + X^2 Y - (B+1) X + D lap
  const double x = w[2], y = w[3];
  return ((p[0] + (x * x) * y) - (p[1] + 1.0) * x) + p[2] * lap;
}
const double x = w[1], y = w[2];                          // Y: B X - X^2 Y + D lap
return (p[1] * x - (x * x) * y) + p[2] * lap;
"""


def bruss_mol(w, p, t, i, n):
    cell, ncell = i >> 1, n >> 1
    if cell == 0 or cell == ncell - 1:
        return 0.0
    lap = (w[0] - 2.0 * w[2]) + w[4]
    if (i & 1) == 0:
        x, y = w[2], w[3]
        return ((p[0] + (x * x) * y) - (p[1] + 1.0) * x) + p[2] * lap
    x, y = w[1], w[2]
    return (p[1] * x - (x * x) * y) + p[2] * lap


@pytest.mark.gpu
def test_user_stencil_two_species_dirichlet_ends():
    F = m.register_stencil("bruss_mol", BRUSS_MOL_BODY, host=bruss_mol, params=(1.0, 3.0, 0.02), radius=2)
    ncell = 300
    xs = np.arange(ncell) / (ncell - 1)
    u0 = np.empty(2 * ncell)
    u0[0::2] = 1.0 + np.sin(2 * np.pi * xs)
    u0[1::2] = 3.0
    ts = [0.0, 0.5, 2.0]
    for method, points in (("ode45", "specified"), ("ode23", "all")):
        t, y, st = getattr(m, method).stats(F, u0, ts, points=points)
        o = O.solve_user_stencil(method, bruss_mol, u0, ts, [1.0, 3.0, 0.02], radius=2, points=points)
        assert np.array_equal(t, o.t) and np.array_equal(y, o.y), method
        assert (st["n_accept"], st["n_reject"], st["n_rhs"]) == (o.n_accept, o.n_reject, o.n_rhs)
        assert np.array_equal(y[-1][:2], u0[:2]) and np.array_equal(y[-1][-2:], u0[-2:])  # the ends did not move


@pytest.mark.gpu
def test_user_field_and_stencil_error_paths():
    u0 = 0.5 + 0.1 * np.sin(2 * np.pi * np.arange(256) / 256)
    with pytest.raises(m.OdeB200Error, match="NVRTC compilation"):
        m.solve_large("ode45", m.register_field("bad_field", "return undefined_thing;"), u0, [0.0, 1.0])
    with pytest.raises(m.OdeB200Error):
        m.register_field("too_many_params", "return 0.0;", params=tuple(range(9)))
    with pytest.raises(ValueError, match="takes 3 parameters"):  # fewer parameters than the functor was registered with
        F = m.register_stencil("three_params", "return p[2] * w[1];", params=(1.0, 2.0, 3.0))
        m.solve_large("ode45", F.with_params([1.0]), u0, [0.0, 1.0])
    with pytest.raises(m.OdeB200Error):  # fields gather from the whole state: single slab only
        import torch
        G = m.register_field("mirror_err", "return p[0] * (y[n - 1 - i] - y[i]);", params=(0.5,))
        m.CudaSlabEngine("ode45", G, 128, 256, 0, 2, [0.0, 1.0], torch.cuda.current_stream().cuda_stream)
    # a stencil and a field without parameters
    t, y = m.ode45(m.register_stencil("no_params", "return (um - 2.0 * u) + up;", params=()), u0, [0.0, 0.1], points="specified")
    t2, y2 = m.ode45(m.register_field("no_params_f", "return (y[i == 0 ? n - 1 : i - 1] - 2.0 * y[i]) + y[i == n - 1 ? 0 : i + 1];"),
                     u0, [0.0, 0.1], points="specified")
    assert np.array_equal(y, y2)
