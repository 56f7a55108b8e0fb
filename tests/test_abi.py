"""CPU-side checks of the C ABI: the library loads, exports every symbol that
include/ode_b200.h declares, its tableaus equal the reference's, and a solve
without a GPU fails loudly instead of falling back to anything."""
import ctypes as C
import json
import os
import re
from fractions import Fraction

import numpy as np
import pytest

import ode_jl_b200 as m

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "ode_b200.h")


def declared_symbols():
    src = open(HEADER).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(ode_b200_[a-z0-9_]+)\s*\(", src)))


def test_every_declared_symbol_is_exported():
    L = m.lib()
    names = declared_symbols()
    assert len(names) >= 25
    for n in names:
        assert hasattr(L, n), "libode_b200.so does not export %s" % n


def test_ids_and_info():
    L = m.lib()
    assert L.ode_b200_version() == 1
    assert L.ode_b200_method_id(b"ode45") == 7 and L.ode_b200_method_id(b"ode45_dp") == 7
    assert L.ode_b200_method_id(b"ode23s") == 9 and L.ode_b200_method_id(b"nope") < 0
    assert L.ode_b200_rhs_id(b"lorenz") == 4 and L.ode_b200_rhs_dof(4) == 3 and L.ode_b200_rhs_nparams(4) == 3
    assert [L.ode_b200_method_stages(i) for i in range(10)] == [1, 2, 2, 4, 2, 4, 6, 7, 13, 3]
    # isFSAL (src/runge_kutta.jl:62) holds for dopri5 only among the RK tableaus
    assert [L.ode_b200_method_is_fsal(i) for i in range(9)] == [0, 0, 0, 0, 0, 0, 0, 1, 0]
    assert [L.ode_b200_method_is_adaptive(i) for i in range(10)] == [0, 0, 0, 0, 1, 1, 1, 1, 1, 1]


GOLD = dict(bt_feuler=0, bt_midpoint=1, bt_heun=2, bt_rk4=3, bt_rk21=4, bt_rk23=5, bt_rk45=6, bt_dopri5=7, bt_feh78=8)


@pytest.mark.parametrize("name", sorted(GOLD))
def test_library_tableaus_equal_reference_source(name):
    g = json.load(open(os.path.join(ROOT, "tests", "golden", "tableaus.json")))[name]
    L = m.lib()
    mid = GOLD[name]
    f = lambda s: float(Fraction(s).numerator) / float(Fraction(s).denominator)
    S = len(g["c"])
    for i in range(S):
        assert L.ode_b200_tableau(mid, 2, i, 0) == f(g["c"][i])
        for j in range(S):
            assert L.ode_b200_tableau(mid, 0, i, j) == f(g["a"][i][j])
    for r in range(len(g["b"])):
        for j in range(S):
            assert L.ode_b200_tableau(mid, 1, r, j) == f(g["b"][r][j])


@pytest.mark.skipif(m.lib().ode_b200_device_count() > 0, reason="a GPU is present")
def test_no_cpu_fallback():
    with pytest.raises(m.OdeB200Error) as ei:
        m.ode45(m.DeviceRHS("linear", 1.01), 0.5, [0.0, 1.0])
    assert ei.value.code == -6 and "no CPU fallback" in str(ei.value)
    with pytest.raises(m.OdeB200Error):
        m.solve_ensemble("ode45", m.DeviceRHS("lorenz", 10.0, 28.0, 8 / 3), np.ones((4, 3)), [0.0, 1.0])
    with pytest.raises(TypeError):
        m.ode45(lambda t, y: y, 0.5, [0.0, 1.0])  # host closures are refused, not run on the CPU


def test_host_rhs_tokens_match_oracle_rhs():
    """DeviceRHS.__call__ (host evaluation for reference runs) states the same expressions."""
    from oracle import oracle as O
    rng = np.random.default_rng(3)
    cases = [("lorenz", [10.0, 28.0, 8 / 3], 3), ("vdp", [5.0], 2), ("robertson", [0.04, 3e7, 1e4], 3),
             ("kepler", [], 4), ("rotation", [], 2), ("linear", [1.01], 1), ("timelin", [2.0], 1), ("const", [6.0], 1)]
    for name, p, d in cases:
        y = rng.uniform(0.1, 2.0, d)
        a = np.atleast_1d(m.DeviceRHS(name, *p)(0.3, y))
        b = O.rhs_eval(name, 0.3, y, p if p else None)
        assert np.array_equal(a, b), name
