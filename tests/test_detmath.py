"""ode_b200_detmath.h (host build, via the oracle library) against mpmath.

The deterministic pow/log10 are the only transcendental functions on the hot
path (src/runge_kutta.jl:436, src/ODE.jl:105-108,357).  They must be exactly
reproducible CPU<->GPU (checked on the GPU in test_gpu_*.py) and as close to
correctly rounded as any libm the reference could be running on.
"""
import numpy as np
import mpmath as mp
import pytest

from oracle import oracle as O
import ctypes as C


def _powv(x, y, mode=0):
    o = np.empty_like(x)
    dp = C.POINTER(C.c_double)
    O.lib().oracle_pow_v(x.ctypes.data_as(dp), y.ctypes.data_as(dp), o.ctypes.data_as(dp), x.size, mode)
    return o


def test_pow_correctly_rounded_on_controller_domain():
    mp.mp.prec = 200
    rng = np.random.default_rng(11)
    n = 20000
    x = np.exp(rng.uniform(-35, 35, n))
    y = np.array([0.5, 1.0 / 3, 0.2, 0.125, 0.25, 0.3333333333333333])[rng.integers(0, 6, n)]
    got = _powv(x, y)
    bad = sum(1 for i in range(n) if float(mp.power(mp.mpf(float(x[i])), mp.mpf(float(y[i])))) != got[i])
    assert bad == 0


@pytest.mark.parametrize("y", [1.0 / 3, 0.2, 0.125])
def test_controller_roots_correctly_rounded(y):
    """The K-th root path of the controllers' exponents (dm_root3/5/8): correctly rounded on every sample over the whole
    positive range (subnormal to 1.8e308), at exact powers, next to 1, and equal to the general path except where that
    one misrounds (it never does on these samples either)."""
    mp.mp.prec = 300
    rng = np.random.default_rng(int(1 / y))
    n = 20000
    x = np.concatenate([np.exp(rng.uniform(-700, 700, n // 2)), np.exp(rng.uniform(-3, 3, n // 2)),
                        [5e-324, 1e-310, 2.2250738585072014e-308, 2.0 ** -1022, 1.0, 2.0, 8.0, 32.0, 243.0, 256.0,
                         1.0000000000000002, 0.9999999999999999, 1.7976931348623157e308]])
    got = _powv(x, np.full(x.size, y))
    bad = sum(1 for i in range(x.size) if float(mp.power(mp.mpf(float(x[i])), mp.mpf(y))) != got[i])
    assert bad == 0


def test_pow10_and_log10_correctly_rounded():
    mp.mp.prec = 200
    rng = np.random.default_rng(12)
    n = 10000
    e = rng.uniform(-12, 3, n)
    got = _powv(np.full(n, 10.0), e)
    # |y*log(x)| reaches 28 here, so a near-tie can round the other way: bound the error instead
    worst = 0.0
    for i in range(n):
        ex = mp.power(10, mp.mpf(float(e[i])))
        worst = max(worst, abs(float((mp.mpf(float(got[i])) - ex) / mp.mpf(float(np.spacing(got[i]))))))
    assert worst < 0.50001
    x = np.exp(rng.uniform(-40, 40, n))
    o = np.empty(n)
    dp = C.POINTER(C.c_double)
    O.lib().oracle_log10_v(x.ctypes.data_as(dp), o.ctypes.data_as(dp), n, 0)
    assert sum(1 for i in range(n) if float(mp.log10(mp.mpf(float(x[i])))) != o[i]) == 0


def test_pow_special_cases():
    L = O.lib()
    inf = float("inf")
    assert L.oracle_pow(0.0, 0.2, 0) == 0.0
    assert L.oracle_pow(inf, 0.2, 0) == inf          # err == 0 -> 1/err = Inf -> newdt = maxstep
    assert L.oracle_pow(inf, -0.2, 0) == 0.0
    assert np.isnan(L.oracle_pow(float("nan"), 0.2, 0))
    assert L.oracle_pow(1.0, 0.2, 0) == 1.0
    assert L.oracle_pow(4.0, 0.5, 0) == 2.0
    assert L.oracle_pow(5e-324, 0.5, 0) == float(mp.sqrt(mp.mpf(5e-324)))
    assert L.oracle_pow(1e308, 2.0, 0) == inf
    assert L.oracle_log10(1000.0, 0) == 3.0 and L.oracle_log10(1.0, 0) == 0.0
    assert L.oracle_log10(0.0, 0) == -inf


def test_detmath_close_to_libm():
    rng = np.random.default_rng(13)
    x = np.exp(rng.uniform(-30, 30, 100000))
    y = np.full_like(x, 0.2)
    a, b = _powv(x, y, 0), _powv(x, y, 1)
    # both are within an ulp of each other, and they differ on well under 1% of inputs
    assert np.max(np.abs(a - b) / np.spacing(np.abs(a))) <= 1.0
    assert np.mean(a != b) < 0.01


def test_tables_header_is_what_the_generator_writes():
    """include/ode_b200_detmath_tables.h is generated (tools/gen_detmath_tables.py, mpmath): the committed file must be
    byte for byte the generator's output, so a table can never drift from its documented construction."""
    import io
    import os
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    sys.path.insert(0, os.path.join(root, "tools"))
    import gen_detmath_tables as G
    buf = io.StringIO()
    G.main(buf)
    mp.mp.prec = 200  # (the generator raises the working precision; restore the tests' own)
    assert buf.getvalue() == open(os.path.join(root, "include", "ode_b200_detmath_tables.h")).read()
