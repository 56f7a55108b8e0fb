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


# ---- the header from plain C, and the struct declarations of the other bindings against it --------------------------
def build_abi_smoke(tmp):
    import subprocess
    exe = os.path.join(str(tmp), "abi_smoke")
    libdir = os.path.join(ROOT, "ode.jl_b200")
    cc = "/usr/bin/gcc" if os.path.exists("/usr/bin/gcc") else "gcc"
    subprocess.check_call([cc, "-std=c11", "-Wall", "-Wextra", "-Werror", "-I" + os.path.join(ROOT, "include"),
                           os.path.join(ROOT, "tests", "abi_smoke.c"), "-o", exe, "-L" + libdir, "-lode_b200", "-lm",
                           "-Wl,-rpath," + libdir])
    return exe


def test_header_compiles_as_c11_and_layouts_hold(tmp_path):
    """tests/abi_smoke.c: static_asserts on sizeof/offsetof of the ABI structs; the program itself solves the README
    problem on a GPU box and must get ODE_B200_E_NO_DEVICE here."""
    import subprocess
    exe = build_abi_smoke(tmp_path)
    r = subprocess.run([exe], capture_output=True, text=True)
    assert r.returncode == 0, r.stdout + r.stderr
    if m.lib().ode_b200_device_count() < 1:
        assert "ODE_B200_E_NO_DEVICE" in r.stdout


def c_struct_fields(name):
    """[(ctype, field, array_len)] of `typedef struct { ... } name;` in the header."""
    src = re.sub(r"/\*.*?\*/", "", open(HEADER).read(), flags=re.S)
    body = re.search(r"typedef struct\s*\{([^}]*)\}\s*%s\s*;" % name, src).group(1)
    out = []
    for decl in body.split(";"):
        decl = " ".join(decl.split())
        if not decl:
            continue
        mm = re.match(r"(const )?([a-z0-9_]+)\s*(\*?)\s*(.*)$", decl)
        ctype, ptr, names = mm.group(2), mm.group(3), mm.group(4)
        for nm in names.split(","):
            nm = nm.strip()
            p = ptr or ("*" if nm.startswith("*") else "")
            nm = nm.lstrip("* ")
            arr = re.match(r"([a-z0-9_]+)\[(\d+)\]", nm)
            out.append((ctype + p, arr.group(1) if arr else nm, int(arr.group(2)) if arr else 0))
    return out


def test_python_and_julia_struct_declarations_match_the_header():
    from ode_jl_b200 import _lib
    ctype_of = {"int32_t": C.c_int32, "int64_t": C.c_int64, "double": C.c_double}
    for cname, py in (("ode_b200_opts", _lib.Opts), ("ode_b200_ensemble_io", _lib.EnsembleIO),
                      ("ode_b200_large_stats", _lib.LargeStats)):
        want = c_struct_fields(cname)
        assert [f for _, f, _ in want] == [f[0] for f in py._fields_], cname
        for (ct, fname, arr), (pname, ptype) in zip(want, py._fields_):
            if ct.endswith("*"):
                assert ptype is C.c_void_p, (cname, fname)
            elif arr:
                assert ptype._length_ == arr and ptype._type_ is ctype_of[ct], (cname, fname)
            else:
                assert ptype is ctype_of[ct], (cname, fname)
    from oracle import oracle as O
    assert [f[0] for f in O.Opts._fields_] == [f for _, f, _ in c_struct_fields("ode_b200_opts")]
    assert C.sizeof(_lib.Opts) == 112 and C.sizeof(O.Opts) == 112
    # julia/ODEB200.jl: same fields, same order, same widths
    jl = open(os.path.join(ROOT, "julia", "ODEB200.jl")).read()
    jtype = {"int32_t": "Int32", "int64_t": "Int64", "double": "Float64"}
    for cname, jname in (("ode_b200_opts", "Opts"), ("ode_b200_ensemble_io", "EnsembleIO"),
                         ("ode_b200_large_stats", "LargeStats")):
        body = re.search(r"^struct %s\n(.*?)^end" % jname, jl, flags=re.S | re.M).group(1)
        body = re.sub(r"#.*", "", body)
        fields = re.findall(r"([a-z0-9_]+)::([A-Za-z0-9{},]+)", body)
        want = c_struct_fields(cname)
        assert [f for f, _ in fields] == [f for _, f, _ in want], jname
        for (ct, fname, arr), (_, jt) in zip(want, fields):
            if ct.endswith("*"):
                assert jt.startswith("Ptr{"), (jname, fname)
            elif arr:
                assert jt == "NTuple{%d,%s}" % (arr, jtype[ct]), (jname, fname)
            else:
                assert jt == jtype[ct], (jname, fname)
    # every ccall in the shim names a symbol the header declares
    for sym in set(re.findall(r"\(:(ode_b200_[a-z0-9_]+), LIB\[\]\)", jl)):
        assert sym in declared_symbols(), sym
