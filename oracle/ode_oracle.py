"""Independent pure-Python restatement of ODE.jl's hot path (python floats).

TEST INFRASTRUCTURE ONLY.  Written separately from oracle/ode_oracle.cpp so the
two can be cross-checked bit for bit (python floats are IEEE binary64 and never
fuse multiply-add; `**` and math.log10 are the C library's, i.e. this file
corresponds to the C++ oracle's mathmode=1).  Small cases only.

Follows /root/reference: src/runge_kutta.jl:66-157 (tableaus), :176-212
(oderk_fixed), :232-369 (oderk_adapt), :371-399 (rk_embedded_step!), :401-442
(stepsize_hw92!), :444-458 (calc_next_k!), :479-492 (hermite_interp!);
src/ODE.jl:85-112 (hinit), :232-243 (fdjacobian), :252-361 (ode23s).
"""
import math
from fractions import Fraction as Fr

NAN = float("nan")
INF = float("inf")


# -- Julia Base semantics ------------------------------------------------------
def jmin(a, b):
    if a != a or b != b:
        return NAN
    if a < b:
        return a
    if b < a:
        return b
    return a if math.copysign(1.0, a) < 0 else b


def jmax(a, b):
    if a != a or b != b:
        return NAN
    if a > b:
        return a
    if b > a:
        return b
    return b if math.copysign(1.0, a) < 0 else a


def jsign(x):
    return 1.0 if x > 0 else (-1.0 if x < 0 else x)


def jeps(x):
    if math.isinf(x) or x != x:
        return NAN
    return math.ulp(abs(x))


def norminf(v):
    m = abs(v[0])
    for x in v[1:]:
        a = abs(x)
        m = m if (m != m or m > a) else a
    return m


def norm2(v):
    """generic_norm2 (n < 32)."""
    mx = norminf(v)
    if mx == 0.0 or math.isinf(mx):
        return mx
    n = len(v)
    assert n < 32, "python restatement covers the small-vector norm only"
    if math.isfinite(n * mx * mx) and mx * mx != 0.0:
        s = v[0] * v[0]
        for x in v[1:]:
            s += x * x
        return math.sqrt(s)
    s = (abs(v[0]) / mx) ** 2
    for x in v[1:]:
        s += (abs(x) / mx) * (abs(x) / mx)
    return mx * math.sqrt(s)


def jpow(x, y):
    """Base ^ for non-negative bases via the C library."""
    if x == 0.0:
        return 0.0 if y > 0 else INF
    if math.isinf(x):
        return INF if y > 0 else 0.0
    if x != x:
        return NAN
    try:
        return math.pow(x, y)
    except OverflowError:
        return INF


# -- tableaus (rationals exactly as in the source) -----------------------------
def _q(rows):
    return [[Fr(x) for x in r] for r in rows]


def _parse(txt):
    return [[Fr(tok) for tok in line.split()] for line in txt.strip().splitlines()]


TABLEAUS = {
    "ode1": dict(order=(1,), a=_q([[0]]), b=_q([[1]]), c=[Fr(0)]),
    "ode2_midpoint": dict(order=(2,), a=_parse("0 0\n1/2 0"), b=_parse("0 1"), c=[Fr(0), Fr(1, 2)]),
    "ode2_heun": dict(order=(2,), a=_parse("0 0\n1 0"), b=_parse("1/2 1/2"), c=[Fr(0), Fr(1)]),
    "ode4": dict(order=(4,), a=_parse("0 0 0 0\n1/2 0 0 0\n0 1/2 0 0\n0 0 1 0"), b=_parse("1/6 1/3 1/3 1/6"),
                 c=[Fr(0), Fr(1, 2), Fr(1, 2), Fr(1)]),
    "ode21": dict(order=(2, 1), a=_parse("0 0\n1 0"), b=_parse("1/2 1/2\n1 0"), c=[Fr(0), Fr(1)]),
    "ode23": dict(order=(2, 3), a=_parse("0 0 0 0\n1/2 0 0 0\n0 3/4 0 0\n2/9 1/3 4/9 0"),
                  b=_parse("7/24 1/4 1/3 1/8\n2/9 1/3 4/9 0"), c=[Fr(0), Fr(1, 2), Fr(3, 4), Fr(1)]),
    "ode45_fe": dict(order=(4, 5), a=_parse("""
        0 0 0 0 0 0
        1/4 0 0 0 0 0
        3/32 9/32 0 0 0 0
        1932/2197 -7200/2197 7296/2197 0 0 0
        439/216 -8 3680/513 -845/4104 0 0
        -8/27 2 -3544/2565 1859/4104 -11/40 0"""),
                     b=_parse("25/216 0 1408/2565 2197/4104 -1/5 0\n16/135 0 6656/12825 28561/56430 -9/50 2/55"),
                     c=[Fr(0), Fr(1, 4), Fr(3, 8), Fr(12, 13), Fr(1), Fr(1, 2)]),
    "ode45": dict(order=(5, 4), a=_parse("""
        0 0 0 0 0 0 0
        1/5 0 0 0 0 0 0
        3/40 9/40 0 0 0 0 0
        44/45 -56/15 32/9 0 0 0 0
        19372/6561 -25360/2187 64448/6561 -212/729 0 0 0
        9017/3168 -355/33 46732/5247 49/176 -5103/18656 0 0
        35/384 0 500/1113 125/192 -2187/6784 11/84 0"""),
                  b=_parse("35/384 0 500/1113 125/192 -2187/6784 11/84 0\n"
                           "5179/57600 0 7571/16695 393/640 -92097/339200 187/2100 1/40"),
                  c=[Fr(0), Fr(1, 5), Fr(3, 10), Fr(4, 5), Fr(8, 9), Fr(1), Fr(1)]),
    "ode78": dict(order=(7, 8), a=_parse("""
        0 0 0 0 0 0 0 0 0 0 0 0 0
        2/27 0 0 0 0 0 0 0 0 0 0 0 0
        1/36 1/12 0 0 0 0 0 0 0 0 0 0 0
        1/24 0 1/8 0 0 0 0 0 0 0 0 0 0
        5/12 0 -25/16 25/16 0 0 0 0 0 0 0 0 0
        1/20 0 0 1/4 1/5 0 0 0 0 0 0 0 0
        -25/108 0 0 125/108 -65/27 125/54 0 0 0 0 0 0 0
        31/300 0 0 0 61/225 -2/9 13/900 0 0 0 0 0 0
        2 0 0 -53/6 704/45 -107/9 67/90 3 0 0 0 0 0
        -91/108 0 0 23/108 -976/135 311/54 -19/60 17/6 -1/12 0 0 0 0
        2383/4100 0 0 -341/164 4496/1025 -301/82 2133/4100 45/82 45/164 18/41 0 0 0
        3/205 0 0 0 0 -6/41 -3/205 -3/41 3/41 6/41 0 0 0
        -1777/4100 0 0 -341/164 4496/1025 -289/82 2193/4100 51/82 33/164 12/41 0 1 0"""),
                  b=_parse("41/840 0 0 0 0 34/105 9/35 9/35 9/280 9/280 41/840 0 0\n"
                           "0 0 0 0 0 34/105 9/35 9/35 9/280 9/280 0 41/840 41/840"),
                  c=[Fr(0), Fr(2, 27), Fr(1, 9), Fr(1, 6), Fr(5, 12), Fr(1, 2), Fr(5, 6), Fr(1, 6), Fr(2, 3),
                     Fr(1, 3), Fr(1), Fr(0), Fr(1)]),
}
TABLEAUS["ode45_dp"] = TABLEAUS["ode45"]
# EXTENSION, not in the reference source (README.md:57 names it): Cash & Karp, ACM TOMS 16 (1990); order (5,4) like
# bt_dopri5 -- the solver advances with the 5th-order row
TABLEAUS["ode45_ck"] = dict(order=(5, 4), a=_parse("""
        0 0 0 0 0 0
        1/5 0 0 0 0 0
        3/40 9/40 0 0 0 0
        3/10 -9/10 6/5 0 0 0
        -11/54 5/2 -70/27 35/27 0 0
        1631/55296 175/512 575/13824 44275/110592 253/4096 0"""),
                            b=_parse("37/378 0 250/621 125/594 0 512/1771\n"
                                     "2825/27648 0 18575/48384 13525/55296 277/14336 1/4"),
                            c=[Fr(0), Fr(1, 5), Fr(3, 10), Fr(3, 5), Fr(1), Fr(7, 8)])


def _f(fr):
    return float(fr.numerator) / float(fr.denominator)  # Rational -> Float64


class Tab:
    def __init__(self, name):
        d = TABLEAUS[name]
        self.order = d["order"]
        self.a = [[_f(x) for x in r] for r in d["a"]]
        self.b = [[_f(x) for x in r] for r in d["b"]]
        self.c = [_f(x) for x in d["c"]]
        self.S = len(self.c)
        self.adaptive = len(self.b) == 2
        self.fsal = self.adaptive and self.a[-1] == self.b[0] and self.c[-1] == 1.0
        self.exact = d


# -- right-hand sides ----------------------------------------------------------
def rhs_factory(name, p):
    if name == "linear":
        return lambda t, y: [p[0] * y[0]]
    if name == "const":
        return lambda t, y: [p[0]]
    if name == "timelin":
        return lambda t, y: [p[0] * t]
    if name == "rotation":
        return lambda t, y: [-y[1], y[0]]
    if name == "lorenz":
        return lambda t, y: [p[0] * (y[1] - y[0]), y[0] * (p[1] - y[2]) - y[1], y[0] * y[1] - p[2] * y[2]]
    if name == "vdp":
        return lambda t, y: [y[1], p[0] * ((1.0 - y[0] * y[0]) * y[1]) - y[0]]
    if name == "robertson":
        def f(t, y):
            d1 = -p[0] * y[0] + p[2] * y[1] * y[2]
            d3 = p[1] * y[1] * y[1]
            return [d1, -d1 - d3, d3]
        return f
    if name == "kepler":
        def f(t, y):
            r2 = y[0] * y[0] + y[1] * y[1]
            r3 = r2 * math.sqrt(r2)
            return [y[2], y[3], -y[0] / r3, -y[1] / r3]
        return f
    if name == "reaction_diffusion":
        def f(t, y):
            n = len(y)
            return [p[0] * ((y[i - 1] - 2.0 * y[i]) + y[(i + 1) % n]) + (p[1] * y[i]) * (1.0 - y[i])
                    for i in range(n)]
        return f
    raise KeyError(name)


class Counted:
    def __init__(self, f):
        self.f, self.n = f, 0

    def __call__(self, t, y):
        self.n += 1
        return self.f(t, y)


# -- hinit ---------------------------------------------------------------------
def hinit(F, x0, t0, tend, p, reltol, abstol):
    tdir = jsign(tend - t0)
    if tdir == 0:
        raise ValueError("Zero time span")
    tau = jmax(reltol * norminf(x0), abstol)
    d0 = norminf(x0) / tau
    f0 = F(t0, x0)
    d1 = norminf(f0) / tau
    h0 = 1e-6 if (d0 < 1e-5 or d1 < 1e-5) else 0.01 * (d0 / d1)
    th = tdir * h0
    x1 = [a + th * b for a, b in zip(x0, f0)]
    f1 = F(t0 + tdir * h0, x1)
    d2 = norminf([a - b for a, b in zip(f1, f0)]) / (tau * h0)
    dm = jmax(d1, d2)
    if dm <= 1e-15:
        h1 = jmax(1e-6, 1e-3 * h0)
    else:
        pw = -(2.0 + math.log10(dm)) / (p + 1)
        h1 = jpow(10.0, pw)
    return tdir * jmin(jmin(100.0 * h0, h1), tdir * (tend - t0)), tdir, f0


def hermite(tq, t, dt, y0, y1, f0, f1):
    th = (tq - t) / dt
    return [((1.0 - th) * a + th * b) + (th * (th - 1.0)) *
            (((1.0 - 2.0 * th) * (b - a) + ((th - 1.0) * dt) * g0) + (th * dt) * g1)
            for a, b, g0, g1 in zip(y0, y1, f0, f1)]


def calc_next_k(ks, y, s, F, t, dt, bt):
    ytmp = list(y)
    for ss in range(s):
        a = bt.a[s][ss]
        k = ks[ss]
        for d in range(len(y)):
            ytmp[d] += dt * k[d] * a
    ks[s] = F(t + bt.c[s] * dt, ytmp)


def oderk_adapt(F, y0, tspan, method, reltol=1e-5, abstol=1e-8, minstep=None, maxstep=None, initstep=0.0,
                points="all", max_steps=10 ** 8):
    bt = Tab(method)
    assert bt.adaptive
    F = Counted(F)
    S = bt.S
    order = min(bt.order)
    tspan = [float(x) for x in tspan]
    span = abs(tspan[-1] - tspan[0])
    minstep = span / 1e18 if minstep is None else minstep
    maxstep = span / 2.5 if maxstep is None else maxstep
    dof = len(y0)
    t = tstart = tspan[0]
    tend = tspan[-1]
    y = [float(v) for v in y0]
    nfixed = len(tspan)
    if points == "specified":
        ys = [list(y)] + [[0.0] * dof for _ in range(nfixed - 1)]
        tout = list(tspan)
    else:
        ys = [list(y)]
        tout = [tstart]
    iter_fixed = 2
    ks = [None] * S
    dt, tdir, ks[0] = hinit(F, y, tstart, tend, order, reltol, abstol)
    if initstep != 0:
        if jsign(initstep) != tdir:
            raise ValueError("initstep has wrong sign.")
        dt = initstep
    islast = abs(t + dt - tend) <= jeps(tend)
    timeout = 0
    it = 2
    nacc = nrej = 0
    trace = []
    status = 0
    attempts = 0
    while True:
        if attempts >= max_steps:
            status = 2
            break
        attempts += 1
        ytrial = [0.0 + bt.b[0][0] * k for k in ks[0]]
        yerr = [0.0 + bt.b[1][0] * k for k in ks[0]]
        for s in range(1, S):
            calc_next_k(ks, y, s, F, t, dt, bt)
            for d in range(dof):
                ytrial[d] += bt.b[0][s] * ks[s][d]
                yerr[d] += bt.b[1][s] * ks[s][d]
        for d in range(dof):
            yerr[d] = dt * (ytrial[d] - yerr[d])
            ytrial[d] = y[d] + dt * ytrial[d]
        # stepsize_hw92!
        nanout = False
        for d in range(dof):
            if ytrial[d] != ytrial[d]:
                nanout = True
                break
            yerr[d] = yerr[d] / (abstol + jmax(abs(y[d]), abs(ytrial[d])) * reltol)
        if nanout:
            err, newdt, timeout = 10.0, dt * (1.0 / 5.0), 5
        else:
            err = norm2(yerr)
            nd = jmin(maxstep, (tdir * dt) * jmax(1.0 / 5.0, 0.8 * jpow(1.0 / err if err != 0 else INF,
                                                                          1.0 / (order + 1))))
            if timeout > 0:
                nd = jmin(nd, dt)
                timeout -= 1
            newdt = tdir * nd
        trace.append((t, dt, err, 1.0 if err <= 1 else 0.0))
        if err <= 1:
            nacc += 1
            f0 = ks[0]
            f1 = ks[S - 1] if bt.fsal else F(t + dt, ytrial)
            if points == "specified":
                while it - 1 < nfixed and (tdir * tspan[it - 1] < tdir * (t + dt) or islast):
                    ys[it - 1] = hermite(tspan[it - 1], t, dt, y, ytrial, f0, f1)
                    it += 1
            elif points == "all":
                while iter_fixed - 1 < nfixed and tdir * t < tdir * tspan[iter_fixed - 1] < tdir * (t + dt):
                    ys.append(hermite(tspan[iter_fixed - 1], t, dt, y, ytrial, f0, f1))
                    tout.append(tspan[iter_fixed - 1])
                    iter_fixed += 1
                    it += 1
                ys.append(list(ytrial))
                tout.append(t + dt)
                it += 1
            else:
                ys = [list(ytrial)]
                tout = [t + dt]
            ks[0] = f1
            if islast:
                break
            y, ytrial = ytrial, y
            t += dt
            dt = newdt
            if tdir * (t + dt + dt / 100) >= tdir * tend:
                dt = tend - t
                islast = True
        elif abs(newdt) < minstep:
            status = 1
            break
        else:
            islast = False
            nrej += 1
            dt = newdt
            timeout = 5
            if not math.isfinite(dt):
                status = 3
                break
    return dict(t=tout, y=ys, n_accept=nacc, n_reject=nrej, n_rhs=F.n, trace=trace, status=status)


def oderk_fixed(F, y0, tspan, method):
    bt = Tab(method)
    F = Counted(F)
    tspan = [float(x) for x in tspan]
    ys = [[float(v) for v in y0]]
    ks = [None] * bt.S
    for i in range(len(tspan) - 1):
        dt = tspan[i + 1] - tspan[i]
        yn = list(ys[i])
        for s in range(bt.S):
            calc_next_k(ks, ys[i], s, F, tspan[i], dt, bt)
            for d in range(len(yn)):
                yn[d] += dt * bt.b[0][s] * ks[s][d]
        ys.append(yn)
    return dict(t=tspan, y=ys, n_rhs=F.n)


# -- ode23s --------------------------------------------------------------------
def fdjacobian(F, x, t):
    ftx = F(t, x)
    n = len(x)
    J = [[0.0] * n for _ in range(n)]
    for j in range(n):
        dx = [0.0] * n
        dx[j] = (x[j] + (1.0 if x[j] == 0 else 0.0)) / 100
        fp = F(t, [a + b for a, b in zip(x, dx)])
        for i in range(n):
            J[i][j] = (fp[i] - ftx[i]) / dx[j]
    return J


def lu_factor(a):
    n = len(a)
    a = [list(r) for r in a]
    piv = []
    for j in range(n):
        p = j
        mx = abs(a[j][j])
        for i in range(j + 1, n):
            if abs(a[i][j]) > mx:
                mx, p = abs(a[i][j]), i
        piv.append(p)
        if a[p][j] == 0.0:
            raise ZeroDivisionError("singular")
        if p != j:
            a[j], a[p] = a[p], a[j]
        rp = 1.0 / a[j][j]
        for i in range(j + 1, n):
            a[i][j] = a[i][j] * rp
        for c in range(j + 1, n):
            u = a[j][c]
            for i in range(j + 1, n):
                a[i][c] = a[i][c] - a[i][j] * u
    return a, piv


def lu_solve(lu, b):
    a, piv = lu
    n = len(a)
    b = list(b)
    for j in range(n):
        if piv[j] != j:
            b[j], b[piv[j]] = b[piv[j]], b[j]
    for j in range(n):
        for i in range(j + 1, n):
            b[i] = b[i] - b[j] * a[i][j]
    for j in range(n - 1, -1, -1):
        b[j] = b[j] / a[j][j]
        for i in range(j):
            b[i] = b[i] - b[j] * a[i][j]
    return b


def ode23s(F, y0, tspan, reltol=1e-5, abstol=1e-8, minstep=None, maxstep=None, initstep=0.0, points="all",
           max_steps=10 ** 8, jac=None):
    F = Counted(F)
    jacf = (lambda yy, tt: jac(tt, yy)) if jac is not None else (lambda yy, tt: fdjacobian(F, yy, tt))
    tspan = [float(x) for x in tspan]
    span = abs(tspan[-1] - tspan[0])
    minstep = span / 1e18 if minstep is None else minstep
    maxstep = span / 2.5 if maxstep is None else maxstep
    d = 1.0 / (2.0 + math.sqrt(2.0))
    e32 = 6.0 + math.sqrt(2.0)
    t = tspan[0]
    tfinal = tspan[-1]
    y = [float(v) for v in y0]
    n = len(y)
    h = initstep
    if h == 0.0:
        h, tdir, F0 = hinit(F, y, t, tfinal, 3, reltol, abstol)
    else:
        tdir = jsign(tfinal - t)
        F0 = F(t, y)
    h = tdir * jmin(abs(h), maxstep)
    tout = [t]
    yout = [list(y)]
    J = jacf(y, t)
    nacc = nrej = 0
    trace = []
    attempts = 0
    status = 0
    while abs(t - tfinal) > 0 and minstep < abs(h):
        if attempts >= max_steps:
            status = 2
            break
        attempts += 1
        if abs(t - tfinal) < abs(h):
            h = tfinal - t
        hd = h * d
        W = [[(-(hd * J[i][j]) + 1.0) if i == j else -(hd * J[i][j]) for j in range(n)] for i in range(n)]
        lu = lu_factor(W)
        Fa = F(t + h / 100, y)
        T = [(hd * (a - b)) / (h / 100) for a, b in zip(Fa, F0)]
        k1 = lu_solve(lu, [a + b for a, b in zip(F0, T)])
        hh = 0.5 * h
        F1 = F(t + 0.5 * h, [a + hh * b for a, b in zip(y, k1)])
        k2 = lu_solve(lu, [a - b for a, b in zip(F1, k1)])
        k2 = [a + b for a, b in zip(k2, k1)]
        ynew = [a + h * b for a, b in zip(y, k2)]
        F2 = F(t + h, ynew)
        k3 = lu_solve(lu, [((F2[i] - e32 * (k2[i] - F1[i])) - 2 * (k1[i] - F0[i])) + T[i] for i in range(n)])
        err = (abs(h) / 6) * norm2([(k1[i] - 2 * k2[i]) + k3[i] for i in range(n)])
        delta = jmax(reltol * jmax(norm2(y), norm2(ynew)), abstol)
        acc = err <= delta
        trace.append((t, h, err / delta if delta != 0 else NAN, 1.0 if acc else 0.0))
        if acc:
            nacc += 1
            if points == "final":
                tout, yout = [t + h], [list(ynew)]
            else:
                for toi in tspan:
                    if toi > t and toi <= t + h:
                        s = (toi - t) / h
                        den = 1 - 2 * d
                        tout.append(toi)
                        yout.append([y[i] + h * (((k1[i] * s) * (1 - s)) / den + ((k2[i] * s) * (s - 2 * d)) / den)
                                     for i in range(n)])
                if points == "all" and tout[-1] != t + h:
                    tout.append(t + h)
                    yout.append(list(ynew))
            t = t + h
            y = ynew
            F0 = F2
            J = jacf(y, t)
        else:
            nrej += 1
        h = tdir * jmin(maxstep, (abs(h) * 0.8) * jpow(delta / err if err != 0 else INF, 1.0 / 3))
    if status == 0 and abs(t - tfinal) > 0:
        status = 1 if math.isfinite(h) else 3
    return dict(t=tout, y=yout, n_accept=nacc, n_reject=nrej, n_rhs=F.n, trace=trace, status=status)


# -- registered analytic Jacobians (same entry expressions as the C++ oracle) ----------
def jac_factory(name, p):
    if name == "linear":
        return lambda t, y: [[p[0]]]
    if name in ("const", "timelin"):
        return lambda t, y: [[0.0]]
    if name == "rotation":
        return lambda t, y: [[0.0, -1.0], [1.0, 0.0]]
    if name == "lorenz":
        return lambda t, y: [[-p[0], p[0], 0.0], [p[1] - y[2], -1.0, -y[0]], [y[1], y[0], -p[2]]]
    if name == "vdp":
        return lambda t, y: [[0.0, 1.0], [p[0] * ((-2.0 * y[0]) * y[1]) - 1.0, p[0] * (1.0 - y[0] * y[0])]]
    if name == "robertson":
        def J(t, y):
            a, b, c = p[2] * y[2], p[2] * y[1], (2.0 * p[1]) * y[1]
            return [[-p[0], a, b], [p[0], -a - c, -b], [0.0, c, 0.0]]
        return J
    if name == "kepler":
        def J(t, y):
            x, yy = y[0], y[1]
            r2 = x * x + yy * yy
            r3 = r2 * math.sqrt(r2)
            r5 = r3 * r2
            return [[0.0, 0.0, 1.0, 0.0], [0.0, 0.0, 0.0, 1.0],
                    [(3.0 * x) * x / r5 - 1.0 / r3, (3.0 * x) * yy / r5, 0.0, 0.0],
                    [(3.0 * x) * yy / r5, (3.0 * yy) * yy / r5 - 1.0 / r3, 0.0, 0.0]]
        return J
    raise KeyError(name)


ROS = {
    "ode4s_kr": (0.231,
                 [[0, 0, 0, 0], [2, 0, 0, 0], [4.452470820736, 4.16352878860, 0, 0], [4.452470820736, 4.16352878860, 0, 0]],
                 [3.95750374663, 4.62489238836, 0.617477263873, 1.28261294568],
                 [[0, 0, 0, 0], [-5.07167533877, 0, 0, 0], [6.02015272865, 0.1597500684673, 0, 0],
                  [-1.856343618677, -8.50538085819, -2.08407513602, 0]]),
    "ode4s_s": (0.5,
                [[0, 0, 0, 0], [2, 0, 0, 0], [48 / 25, 6 / 25, 0, 0], [48 / 25, 6 / 25, 0, 0]],
                [19 / 9, 1 / 2, 25 / 108, 125 / 108],
                [[0, 0, 0, 0], [-8, 0, 0, 0], [372 / 25, 12 / 5, 0, 0], [-112 / 125, -54 / 125, -2 / 5, 0]]),
}
ROS["ode4s"] = ROS["ode4s_s"]


def oderosenbrock(F, x0, tspan, method, jac=None):
    """src/ODE.jl:366-404."""
    gamma, a, b, c = ROS[method]
    F = Counted(F)
    tspan = [float(v) for v in tspan]
    xs_all = [[float(v) for v in x0]]
    n = len(x0)
    for st in range(len(tspan) - 1):
        ts, hs = tspan[st], tspan[st + 1] - tspan[st]
        xs = xs_all[st]
        J = jac(ts, xs) if jac is not None else fdjacobian(F, xs, ts)
        lam = 1.0 / (gamma * hs)
        W = [[(-J[i][j] + lam) if i == j else -J[i][j] for j in range(n)] for i in range(n)]
        lu = lu_factor(W)
        g = [lu_solve(lu, F(ts + b[0] * hs, xs))]
        xn = [xs[d] + b[0] * g[0][d] for d in range(n)]
        for i in range(1, 4):
            dx = [0.0] * n
            dF = [0.0] * n
            for j in range(i):
                for d in range(n):
                    dx[d] += a[i][j] * g[j][d]
                    dF[d] += c[i][j] * g[j][d]
            f = F(ts + b[i] * hs, [xs[d] + dx[d] for d in range(n)])
            g.append(lu_solve(lu, [f[d] + dF[d] / hs for d in range(n)]))
            for d in range(n):
                xn[d] += b[i] * g[i][d]
        xs_all.append(xn)
    return dict(t=tspan, y=xs_all, n_rhs=F.n)


def ms_coefficients(order):
    """b of ode_ms (src/ODE.jl:180-194): ms_coefficients4 (:440-443), and for order 5 the reference's floating-point
    evaluation through poly / polyint / polyval (restated: product over the roots, a_k/(k+1), Horner)."""
    from math import factorial
    b4 = [[1, 0, 0, 0], [-1 / 2, 3 / 2, 0, 0], [5 / 12, -4 / 3, 23 / 12, 0], [-9 / 24, 37 / 24, -59 / 24, 55 / 24]]
    if order <= 4:
        return b4
    b = [[0.0] * order for _ in range(order)]
    for i in range(4):
        for j in range(4):
            b[i][j] = float(b4[i][j])
    for s in range(5, order + 1):
        for j in range(s):
            pc = [1.0]
            for r in range(s):
                if r == j:
                    continue
                root = -float(r)
                q = [0.0] * (len(pc) + 1)
                for k, c in enumerate(pc):
                    q[k + 1] += c
                    q[k] += -root * c
                pc = q
            pi = [0.0] + [c / float(k + 1) for k, c in enumerate(pc)]
            v = pi[-1]
            for k in range(len(pi) - 2, -1, -1):
                v = pi[k] + 1.0 * v
            b[s - 1][s - j - 1] = (-1.0) ** j / factorial(j) / factorial(s - 1 - j) * v
    return b


def ode_ms(F, x0, tspan, order):
    """ode_ms, src/ODE.jl:175-209."""
    b = ms_coefficients(order)
    F = Counted(F)
    tspan = [float(v) for v in tspan]
    x = [[float(v) for v in x0]]
    xdot = []
    for i in range(len(tspan) - 1):
        h = tspan[i + 1] - tspan[i]
        so = min(i + 1, order)
        xdot.append(F(tspan[i], x[i]))
        xn = list(x[i])
        for j in range(so):
            xd = xdot[i - (so - 1) + j]
            for d in range(len(xn)):
                xn[d] += h * b[so - 1][j] * xd[d]
        x.append(xn)
    return dict(t=tspan, y=x, n_rhs=F.n)


def ode4ms(F, x0, tspan):
    return ode_ms(F, x0, tspan, 4)


MS_ORDER = dict(ode4ms=4, ode5ms=5, ode_ms1=1, ode_ms2=2, ode_ms3=3, ode_ms4=4, ode_ms5=5)


def solve(method, rhs, y0, tspan, params=None, jacobian=0, **kw):
    F = rhs_factory(rhs, params)
    if method in ROS:
        return oderosenbrock(F, y0, tspan, method, jac_factory(rhs, params) if jacobian else None)
    if method in MS_ORDER:
        return ode_ms(F, y0, tspan, MS_ORDER[method])
    if method == "ode23s":
        return ode23s(F, y0, tspan, jac=jac_factory(rhs, params) if jacobian else None, **kw)
    if Tab(method).adaptive:
        return oderk_adapt(F, y0, tspan, method, **kw)
    return oderk_fixed(F, y0, tspan, method)
