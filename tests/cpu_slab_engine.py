"""A CPU stand-in for the CUDA slab session, used ONLY by the gloo tests.

It follows the same exchange protocol as ode.jl_b200/csrc/large.cu (slots of the
XCHG buffer, ghost refresh on accept, every rank running the same controller on
the gathered partials) with numpy arithmetic in the reference's operation order,
so the multi-process driver (`integrate_slab`) can be exercised without a GPU.
Test infrastructure: it uses the oracle's tableau and pow, never the product's.
"""
from fractions import Fraction

import numpy as np

from oracle import oracle as O

XS_HI, XS_LO, XS_MAX, XS_NAN, XS_AUX0, XS_AUX1, XS_EDGE = 0, 1, 2, 3, 4, 5, 8


def jmax(a, b):
    if a != a or b != b:
        return float("nan")
    return a if a > b else b


def jmin(a, b):
    if a != a or b != b:
        return float("nan")
    return a if a < b else b


class CpuSlabEngine:
    def __init__(self, method, params, n_local, rank, world, tspan, send, recv, reltol=1e-5, abstol=1e-8,
                 max_steps=10 ** 8):
        self.tab = O.tableau(method)
        S = self.tab["S"]
        self.S = S
        self.fsal = self.tab["fsal"]
        self.G = S - 1 + (0 if self.fsal else 1)
        self.H = (self.G + 3) // 4 * 4
        self.n, self.rank, self.world = n_local, rank, world
        self.kap, self.r = params
        self.t0, self.tend = float(tspan[0]), float(tspan[-1])
        span = abs(self.tend - self.t0)
        self.reltol, self.abstol = reltol, abstol
        self.minstep, self.maxstep = span / 1e18, span / 2.5
        self.tdir = 1.0 if self.tend > self.t0 else -1.0
        self.send, self.recv = send, recv  # numpy views of the torch CPU tensors
        H = self.H
        self.Y = [np.zeros(n_local + 2 * H), np.zeros(n_local + 2 * H)]
        self.K = [np.zeros(n_local + 2 * H), np.zeros(n_local + 2 * H)]
        self.par = 0
        self.ctl = dict(done=False, status=0, n_accept=0, n_reject=0, t=self.t0, dt=0.0, err=0.0, t_final=self.t0,
                        attempts=0, timeout=0, islast=False)
        self.max_steps = max_steps
        self.trace = []

    # -- helpers -------------------------------------------------------------
    def F(self, u):
        """stencil RHS on u[1:-1] from a padded array (same association as the device functor)"""
        um, uc, up = u[:-2], u[1:-1], u[2:]
        return self.kap * ((um - 2.0 * uc) + up) + (self.r * uc) * (1.0 - uc)

    def set_state(self, y):
        self.Y[0][self.H:self.H + self.n] = y

    def get_state(self):
        return self.Y[self.par][self.H:self.H + self.n].copy()

    def _pack(self, a, slot):
        H, n = self.H, self.n
        self.send[XS_EDGE + 2 * H * slot:XS_EDGE + 2 * H * slot + H] = a[H:2 * H]
        self.send[XS_EDGE + 2 * H * slot + H:XS_EDGE + 2 * H * slot + 2 * H] = a[n:n + H]

    def _ghost(self, a, slot):
        H, n, w = self.H, self.n, self.world
        lr, rr = (self.rank - 1) % w, (self.rank + 1) % w
        a[:H] = self.recv[lr, XS_EDGE + 2 * H * slot + H:XS_EDGE + 2 * H * slot + 2 * H]
        a[n + H:n + 2 * H] = self.recv[rr, XS_EDGE + 2 * H * slot:XS_EDGE + 2 * H * slot + H]

    @staticmethod
    def _nmax(v):
        m = 0.0
        for x in np.abs(v):
            m = m if (m != m or m > x) else float(x)
        return m

    # -- hinit phases (src/ODE.jl:85-112) --------------------------------------
    def init_phase(self, phase, *_):
        H, n = self.H, self.n
        y, k = self.Y[0], self.K[0]
        if phase == 0:
            self._pack(y, 0)
        elif phase == 1:
            self._ghost(y, 0)
            k[H:H + n] = self.F(y[H - 1:H + n + 1])
            self.send[XS_AUX0] = self._nmax(y[H:H + n])
            self.send[XS_AUX1] = self._nmax(k[H:H + n])
            self._pack(k, 1)
        elif phase == 2:
            n0 = self._nmax(self.recv[:, XS_AUX0])
            n1 = self._nmax(self.recv[:, XS_AUX1])
            self.tau = jmax(self.reltol * n0, self.abstol)
            d0, self.d1 = n0 / self.tau, n1 / self.tau
            self.h0 = 1e-6 if (d0 < 1e-5 or self.d1 < 1e-5) else 0.01 * (d0 / self.d1)
            self._ghost(k, 1)
            th = self.tdir * self.h0
            x1 = y[H - 1:H + n + 1] + th * k[H - 1:H + n + 1]
            f1 = self.F(x1)
            self.send[XS_AUX0] = self._nmax(f1 - k[H:H + n])
        else:
            n2 = self._nmax(self.recv[:, XS_AUX0])
            d2 = n2 / (self.tau * self.h0)
            dm = jmax(self.d1, d2)
            L = O.lib()
            if dm <= 1e-15:
                h1 = jmax(1e-6, 1e-3 * self.h0)
            else:
                h1 = L.oracle_pow(10.0, -(2.0 + L.oracle_log10(dm, 0)) / (self.tab["order"] + 1), 0)
            dt = self.tdir * jmin(jmin(100.0 * self.h0, h1), self.tdir * (self.tend - self.t0))
            c = self.ctl
            c["dt"], c["t"] = dt, self.t0
            c["islast"] = abs(self.t0 + dt - self.tend) <= np.spacing(abs(self.tend))

    # -- one attempt on the slab (rk_embedded_step! with shrinking validity) -------
    def attempt(self, *_):
        c = self.ctl
        if c["done"]:
            return
        H, n, S = self.H, self.n, self.S
        a, b, dt = self.tab["a"], self.tab["b"], c["dt"]
        y, k1 = self.Y[self.par], self.K[self.par]
        ks = [k1]
        ytrial = b[0, 0] * k1 if b[0, 0] != 0 else np.zeros_like(k1)
        yerr = b[1, 0] * k1 if b[1, 0] != 0 else np.zeros_like(k1)
        for s in range(1, S):
            ytmp = y.copy()
            for ss in range(s):
                if a[s, ss] != 0:
                    ytmp = ytmp + (dt * ks[ss]) * a[s, ss]
            k = np.zeros_like(y)
            k[1:-1] = self.F(ytmp)  # edges lose validity by one point per stage
            ks.append(k)
            if b[0, s] != 0:
                ytrial = ytrial + b[0, s] * k
            if b[1, s] != 0:
                yerr = yerr + b[1, s] * k
        yerr = dt * (ytrial - yerr)
        ytrial = y + dt * ytrial
        knext = ks[S - 1]
        if not self.fsal:
            knext = np.zeros_like(y)
            knext[1:-1] = self.F(ytrial)
        sl = slice(H, H + n)
        yt, ye, yy = ytrial[sl], yerr[sl], y[sl]
        nanf = 1.0 if np.isnan(yt).any() else 0.0
        e = ye / (self.abstol + np.maximum(np.abs(yy), np.abs(yt)) * self.reltol)
        tot = sum((Fraction(float(x)) ** 2 for x in e), Fraction(0)) if nanf == 0.0 else Fraction(0)
        hi = float(tot)
        lo = float(tot - Fraction(hi))
        self.send[XS_HI], self.send[XS_LO] = hi, lo
        self.send[XS_MAX], self.send[XS_NAN] = self._nmax(e), nanf
        self.Y[self.par ^ 1][sl] = yt
        self.K[self.par ^ 1][sl] = knext[sl]
        self._pack(self.Y[self.par ^ 1], 0)
        self._pack(self.K[self.par ^ 1], 1)

    # -- controller on the gathered partials (identical on every rank) ------------
    def control(self, *_):
        c = self.ctl
        if c["done"]:
            return
        tot = Fraction(0)
        m = nf = 0.0
        for r in range(self.world):
            tot += Fraction(float(self.recv[r, XS_HI])) + Fraction(float(self.recv[r, XS_LO]))
            x = float(self.recv[r, XS_MAX])
            m = m if (m != m or m > x) else x
            nf = max(nf, float(self.recv[r, XS_NAN]))
        dt, t, tdir = c["dt"], c["t"], self.tdir
        timeout = c["timeout"]
        if nf > 0:
            err, newdt, timeout = 10.0, dt * 0.2, 5
        else:
            err = m if (m == 0.0 or m != m or np.isinf(m)) else float(np.sqrt(np.float64(float(tot))))
            pw = O.lib().oracle_pow(1.0 / err if err != 0 else float("inf"), 1.0 / (self.tab["order"] + 1), 0)
            nd = jmin(self.maxstep, (tdir * dt) * jmax(0.2, 0.8 * pw))
            if timeout > 0:
                nd = jmin(nd, dt)
                timeout -= 1
            newdt = tdir * nd
        c["attempts"] += 1
        c["err"] = err
        acc = err <= 1.0
        self.trace.append((t, dt, err, 1.0 if acc else 0.0))
        if acc:
            c["n_accept"] += 1
            self.par ^= 1
            if c["islast"]:
                c["t_final"], c["done"] = t + dt, True
            else:
                tn, dn, islast = t + dt, newdt, False
                if tdir * (tn + dn + dn / 100) >= tdir * self.tend:
                    dn, islast = self.tend - tn, True
                c["t"], c["dt"], c["islast"], c["t_final"] = tn, dn, islast, tn
                self._ghost(self.Y[self.par], 0)
                self._ghost(self.K[self.par], 1)
        elif abs(newdt) < self.minstep:
            c["done"], c["status"] = True, 1
        else:
            c["islast"] = False
            c["n_reject"] += 1
            c["dt"] = newdt
            timeout = 5
            if not np.isfinite(newdt):
                c["done"], c["status"] = True, 3
        c["timeout"] = timeout
        if not c["done"] and c["attempts"] >= self.max_steps:
            c["done"], c["status"] = True, 2

    def poll(self):
        return dict(self.ctl)
