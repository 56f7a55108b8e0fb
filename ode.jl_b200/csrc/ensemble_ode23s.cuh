// ensemble_ode23s.cuh -- modified Rosenbrock triple, one trajectory per thread.
//
// Replaces n_traj calls of ode23s (/root/reference src/ODE.jl:252-361) with the
// default forward-difference Jacobian fdjacobian (src/ODE.jl:232-243).
// W = I - h*d*J is factorised per attempt as an in-register LU with partial
// pivoting (first maximal |entry|, reciprocal-pivot scaling, right-looking
// update -- the operation order of LAPACK dgetf2 that Julia's `lu` reaches for a
// tiny matrix) and the three solves follow getrs (row swaps, unit-lower
// forward sweep, upper back substitution with division).  Row exchanges use
// compare-and-swap on register rows, so no local-memory indexing appears.
#pragma once

#include "common.cuh"
#include "rhs.cuh"

namespace odeb200 {

template <int D>
struct LuReg {
  double a[D][D];
  double rd[D];  // refined reciprocal of U[j][j] (recip_prepare), shared by 1/pivot and the back substitutions
  int piv[D];
  bool singular;

  __device__ __forceinline__ void factor() {
    bool unused = false;
    factor_t<true>(unused);
  }
  __device__ __forceinline__ void solve(double (&b)[D]) const {
    bool unused = false;
    solve_t<true>(b, unused);
  }
  // EXACT = false: straight-line code, divisions with the deferred acceptance check (div_nb) and no zero-pivot
  // branch -- a zero pivot makes the reciprocal non-finite, which sets `bad` like any other failed check, and the
  // caller repeats the attempt with EXACT = true, where the zero pivot is reported as `singular`.
  template <bool EXACT>
  __device__ __forceinline__ void factor_t(bool& bad) {
    singular = false;
#pragma unroll
    for (int j = 0; j < D; j++) {
      int pr = j;
      double mx = fabs(a[j][j]);
#pragma unroll
      for (int i = j + 1; i < D; i++) {
        double v = fabs(a[i][j]);
        if (v > mx) {
          mx = v;
          pr = i;
        }
      }
      piv[j] = pr;
#pragma unroll
      for (int i = j + 1; i < D; i++) {  // row exchange as selects (keeps the rows in registers)
        const bool sw = (pr == i);
#pragma unroll
        for (int c = 0; c < D; c++) {
          const double aj = a[j][c], ai = a[i][c];
          a[j][c] = sw ? ai : aj;
          a[i][c] = sw ? aj : ai;
        }
      }
      rd[j] = 0.0;
      if (EXACT && a[j][j] == 0.0) {
        singular = true;
      } else {
        Recip R = recip_prepare(a[j][j]);
        rd[j] = R.r;
        double rp = div_sel<EXACT>(1.0, R, bad);
#pragma unroll
        for (int i = j + 1; i < D; i++) a[i][j] = a[i][j] * rp;
#pragma unroll
        for (int c = j + 1; c < D; c++) {
#pragma unroll
          for (int i = j + 1; i < D; i++) a[i][c] = a[i][c] - a[i][j] * a[j][c];
        }
      }
    }
  }
  template <bool EXACT>
  __device__ __forceinline__ void solve_t(double (&b)[D], bool& bad) const {
#pragma unroll
    for (int j = 0; j < D; j++) {
#pragma unroll
      for (int i = j + 1; i < D; i++) {
        const bool sw = (piv[j] == i);
        const double bj = b[j], bi = b[i];
        b[j] = sw ? bi : bj;
        b[i] = sw ? bj : bi;
      }
    }
#pragma unroll
    for (int j = 0; j < D; j++) {
#pragma unroll
      for (int i = j + 1; i < D; i++) b[i] = b[i] - b[j] * a[i][j];
    }
#pragma unroll
    for (int j = D - 1; j >= 0; j--) {
      Recip R;
      R.b = a[j][j];
      R.r = rd[j];
      b[j] = div_sel<EXACT, true>(b[j], R, bad);
#pragma unroll
      for (int i = 0; i < j; i++) b[i] = b[i] - b[j] * a[i][j];
    }
  }
};

// fdjacobian, src/ODE.jl:232-243 (recomputes F(t,x), D+1 RHS calls).
// `known`: F(t, x) as the caller already holds it.  The reference evaluates F(t, x) again (:233), with exactly the
// arguments of a call it has just made -- F0 = F(t0, y0) of hinit (:91) for the first Jacobian (:294), F2 =
// F(t + h, ynew) (:320) for the one after an accepted step (:352) -- and a functor is a deterministic function of its
// arguments, so the stored values are the recomputed ones bit for bit.  The RHS-call counter still counts D + 1.
template <class Rhs, int D, int NPA, bool EXACT = true>
__device__ __forceinline__ void fd_jacobian(double t, const double (&x)[D], const double (&p)[NPA], double (&J)[D][D],
                                            bool& bad, const double* known = nullptr) {
  double ftx[D];
  if (ODEB200_S23_REUSE && known != nullptr) {
#pragma unroll
    for (int i = 0; i < D; i++) ftx[i] = known[i];
  } else {
    Rhs::eval(t, x, ftx, p);  // :233
  }
  const Recip R100 = recip_prepare(100.0);
#pragma unroll
  for (int j = 0; j < D; j++) {  // :236
    const double dxj = div_sel<EXACT>(x[j] + ((x[j] == 0.0) ? 1.0 : 0.0), R100, bad);  // :239
    const Recip Rdx = recip_prepare(dxj);
    double xp[D], fp[D];
#pragma unroll
    for (int i = 0; i < D; i++) xp[i] = x[i] + ((i == j) ? dxj : 0.0);
    Rhs::eval(t, xp, fp, p);
#pragma unroll
    for (int i = 0; i < D; i++) J[i][j] = div_sel<EXACT, true>(fp[i] - ftx[i], Rdx, bad);  // :240
  }
}
template <class Rhs, int D, int NPA>
__device__ __forceinline__ void fd_jacobian(double t, const double (&x)[D], const double (&p)[NPA], double (&J)[D][D]) {
  bool unused = false;
  fd_jacobian<Rhs, D, NPA, true>(t, x, p, J, unused);
}

// jac = jacobian(t, y): the registered analytic functor (JAC = 1) or fdjacobian (JAC = 0), :262-267
template <class Rhs, int JAC, int D, int NPA, bool EXACT = true>
__device__ __forceinline__ int eval_jacobian(double t, const double (&x)[D], const double (&p)[NPA], double (&J)[D][D],
                                             bool& bad, const double* known = nullptr) {
  if (JAC) {
    Rhs::jac(t, x, J, p);
    return 0;
  }
  fd_jacobian<Rhs, D, NPA, EXACT>(t, x, p, J, bad, known);
  return D + 1;
}
template <class Rhs, int JAC, int D, int NPA>
__device__ __forceinline__ int eval_jacobian(double t, const double (&x)[D], const double (&p)[NPA], double (&J)[D][D],
                                             const double* known = nullptr) {
  bool unused = false;
  return eval_jacobian<Rhs, JAC, D, NPA, true>(t, x, p, J, unused, known);
}

// hinit with p = 3 (src/ODE.jl:279-287): returns the first step (before the maxstep clip of :287), F0 and the RHS count
template <class Rhs, int D, int NPA>
__device__ __forceinline__ double hinit_ode23s(const EnsembleArgs& A, double tstart, double tfinal, double tdir,
                                               const double (&y)[D], const double (&p)[NPA], double (&F0)[D], int& nrhs) {
  double h = A.initstep;  // :279
  if (h == 0.0) {         // hinit, :282 (src/ODE.jl:85-112)
    double n0 = fabs(y[0]);
#pragma unroll
    for (int d = 1; d < D; d++) n0 = normInf_step(n0, fabs(y[d]));
    double tau = jl_max(A.reltol * n0, A.abstol);
    double d0 = n0 / tau;
    Rhs::eval(tstart, y, F0, p);
    double n1 = fabs(F0[0]);
#pragma unroll
    for (int d = 1; d < D; d++) n1 = normInf_step(n1, fabs(F0[d]));
    double d1 = n1 / tau;
    double h0 = (d0 < 1e-5 || d1 < 1e-5) ? 1e-6 : 0.01 * (d0 / d1);
    double x1[D], f1[D];
    double th = tdir * h0;
#pragma unroll
    for (int d = 0; d < D; d++) x1[d] = y[d] + th * F0[d];
    Rhs::eval(tstart + tdir * h0, x1, f1, p);
    double n2 = fabs(f1[0] - F0[0]);
#pragma unroll
    for (int d = 1; d < D; d++) n2 = normInf_step(n2, fabs(f1[d] - F0[d]));
    double d2 = n2 / (tau * h0);
    double dm = jl_max(d1, d2);
    double h1;
    if (dm <= 1e-15) {
      h1 = jl_max(1e-6, 1e-3 * h0);
    } else {
      double pw = -(2.0 + dm_log10(dm)) / 4.0;
      h1 = dm_pow(10.0, pw);
    }
    h = tdir * jl_min(jl_min(100.0 * h0, h1), tdir * (tfinal - tstart));
    nrhs = 2;
  } else {  // :284-285
    Rhs::eval(tstart, y, F0, p);
    nrhs = 1;
  }
  return h;
}

// Pre-pass for large ensembles (like ens_adapt_init_kernel): the first step, F0 and the first Jacobian of every
// trajectory with full warps -- inside the persistent kernel a lane that fetches a trajectory would run hinit and
// D+2 more right-hand sides alone while its 31 neighbours wait.  init[idx] = {h, F0[D], J[D][D], rhs count}; h also
// feeds the dispatch order.
template <class Rhs, int JAC>
__global__ void __launch_bounds__(S23_BLOCK) ens_ode23s_init_kernel(const EnsembleArgs A) {
  constexpr int D = Rhs::DOF, NPA = RhsNP<Rhs>::value, W = 2 + D + D * D;
  const double tstart = A.tspan[0];
  const double tfinal = A.tspan[A.n_tspan - 1];
  const double tdir = jl_sign(tfinal - tstart);
  for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < A.n_traj;
       idx += (long long)gridDim.x * blockDim.x) {
    double y[D], F0[D], p[NPA], Jt[D][D];
#pragma unroll
    for (int d = 0; d < D; d++) y[d] = A.y0[idx * D + d];
#pragma unroll
    for (int q = 0; q < NPA; q++) p[q] = (q < Rhs::NP) ? A.params[idx * A.params_stride + q] : 0.0;
    int nrhs = 0;
    double h = hinit_ode23s<Rhs, D, NPA>(A, tstart, tfinal, tdir, y, p, F0, nrhs);
    h = tdir * jl_min(fabs(h), A.maxstep);  // :287
    nrhs += eval_jacobian<Rhs, JAC, D, NPA>(tstart, y, p, Jt, F0);  // :294 (F(tstart, y) = F0)
    double* out = A.init + idx * W;
    out[0] = h;
#pragma unroll
    for (int d = 0; d < D; d++) out[1 + d] = F0[d];
#pragma unroll
    for (int i = 0; i < D; i++) {
#pragma unroll
      for (int j = 0; j < D; j++) out[1 + D + i * D + j] = Jt[i][j];
    }
    out[1 + D + D * D] = (double)nrhs;
  }
}

template <class Rhs, int MODE, int JAC>
#ifdef ODEB200_S23_MAXNREG
__global__ void __maxnreg__(ODEB200_S23_MAXNREG) ens_ode23s_kernel(const EnsembleArgs A) {
#else
__global__ void __launch_bounds__(S23_BLOCK, (Rhs::DOF <= 3) ? ODEB200_S23_MINB : 0) ens_ode23s_kernel(const EnsembleArgs A) {
#endif
  constexpr int D = Rhs::DOF, NPA = RhsNP<Rhs>::value;
  const double dd = 1.0 / (2.0 + sqrt(2.0));  // :270
  const double e32 = 6.0 + sqrt(2.0);         // :271
  const double tstart = A.tspan[0];
  const double tfinal = A.tspan[A.n_tspan - 1];
  const double tdir = jl_sign(tfinal - tstart);
  const int nfixed = A.n_tspan;

  long long idx = -1;
  double y[D], F0[D], p[NPA];
  // The Jacobian lives across attempts (it is rebuilt only after an accepted step) but is read once per attempt:
  // for small systems it is parked in shared memory, [entry][thread] so that a warp's accesses are conflict free,
  // which frees 2*D*D registers for occupancy.
  constexpr bool JS = (D <= ODEB200_S23_JSHARED_MAXD);
  __shared__ double Jsh[JS ? D * D : 1][JS ? S23_BLOCK : 1];
  double Jreg[JS ? 1 : D][JS ? 1 : D];
  double t = 0.0, h = 0.0, last_tout = 0.0;
  double ny = 0.0;  // norm(y), :324: y is the ynew of the last accepted step, whose norm that attempt already took
  constexpr bool AUTO = ODEB200_S23_AUTO && RhsAutonomous<Rhs>::value;
  int nacc = 0, nrej = 0, nrhs = 0, npts = 0;
  long long attempts = 0, ntrace = 0;

  while (true) {
    if (idx < 0) {
      idx = (long long)atomicAdd(A.work_counter, 1ULL);
      if (idx >= A.n_traj) break;
      if (A.order != nullptr) idx = A.order[idx];  // longest expected trajectory first (common.cuh)
#pragma unroll
      for (int d = 0; d < D; d++) y[d] = A.y0[idx * D + d];
#pragma unroll
      for (int q = 0; q < NPA; q++) p[q] = (q < Rhs::NP) ? A.params[idx * A.params_stride + q] : 0.0;
      t = tstart;
      nacc = nrej = nrhs = 0;
      attempts = 0;
      ntrace = 0;
      double Jt[D][D];
      if (MODE == 0 && A.init != nullptr) {  // computed by ens_ode23s_init_kernel with all lanes busy
        const double* in = A.init + idx * (2 + D + D * D);
        h = in[0];
#pragma unroll
        for (int d = 0; d < D; d++) F0[d] = in[1 + d];
#pragma unroll
        for (int i = 0; i < D; i++) {
#pragma unroll
          for (int j = 0; j < D; j++) Jt[i][j] = in[1 + D + i * D + j];
        }
        nrhs = (int)in[1 + D + D * D];
      } else {
        h = hinit_ode23s<Rhs, D, NPA>(A, tstart, tfinal, tdir, y, p, F0, nrhs);
        h = tdir * jl_min(fabs(h), A.maxstep);  // :287
        nrhs += eval_jacobian<Rhs, JAC, D, NPA>(t, y, p, Jt, F0);  // :294 (F(tstart, y) = F0)
      }
      if (ODEB200_S23_NORMY) ny = norm2_small<D>(y);
      if (MODE == 1) {  // tout = [t], yout = [y0] :290-291
        if (A.pts_cap > 0 && A.pts_y) {
          A.pts_t[idx * A.pts_cap] = tstart;
#pragma unroll
          for (int d = 0; d < D; d++) A.pts_y[(idx * A.pts_cap) * D + d] = y[d];
        }
        npts = 1;
        last_tout = tstart;
      }
#pragma unroll
      for (int i = 0; i < D; i++) {
#pragma unroll
        for (int j = 0; j < D; j++) {
          if constexpr (JS) Jsh[i * D + j][threadIdx.x] = Jt[i][j];
          else Jreg[i][j] = Jt[i][j];
        }
      }
    }

    int fin = -1;
    if (!(fabs(t - tfinal) > 0 && A.minstep < fabs(h))) {  // :296
      fin = (fabs(t - tfinal) > 0) ? ODE_B200_ST_MINSTEP : ODE_B200_ST_OK;
      if (fin == ODE_B200_ST_MINSTEP && !is_finite(h)) fin = ODE_B200_ST_NONFINITE;
    } else if (attempts >= A.max_steps) {
      fin = ODE_B200_ST_MAXSTEPS;
    } else {
      attempts++;
      if (fabs(t - tfinal) < fabs(h)) h = tfinal - t;  // :297-299
      const double hd = h * dd;
      // One attempt (:301-324) as straight-line arithmetic.  EXACT = false: every division carries a deferred
      // acceptance check (div_nb) and there is no zero-pivot branch; if any check failed (`bad`) the attempt is
      // repeated with EXACT = true -- the branchy, always-correct divisions -- which produces the same bits whenever
      // the fast pass would have been accepted, so the two passes never disagree.  With ODEB200_S23_SPECJ the
      // Jacobian of the NEXT step, J(t + h, ynew) (:352), is part of the attempt and simply dropped on a rejection.
      double k1[D], k2[D], ynew[D], F2[D], Jn[D][D];
      double err = 0.0, delta = 0.0, nyn = 0.0;
      bool singular = false;
      int jac_rhs = 0;
      auto attempt = [&](auto EX, bool& bad) {
        constexpr bool EXACT = decltype(EX)::value != 0;
        LuReg<D> W;  // :301-310
#pragma unroll
        for (int i = 0; i < D; i++) {
#pragma unroll
          for (int j = 0; j < D; j++) {
            double Jij;
            if constexpr (JS) Jij = Jsh[i * D + j][threadIdx.x];
            else Jij = Jreg[i][j];
            double m = -(hd * Jij);
            W.a[i][j] = (i == j) ? (m + 1.0) : m;
          }
        }
        W.template factor_t<EXACT>(bad);
        singular = W.singular;
        if (EXACT && singular) return;
        double T[D], k3[D], F1[D], tmp[D];
        if constexpr (AUTO && !EXACT) {
          // :313 for a functor that ignores t: F(t + h/100, y) is F(., y) again, and F0 IS F(., y) -- from hinit or,
          // after an accepted step, F2 = F(t + h, ynew) with y = ynew -- so F(t + h/100, y) - F0 = +0 exactly and
          // T = (hd * +0) / (h/100) = +0 (hd and h/100 carry the sign of h), provided F0 is finite and h/100 does not
          // underflow to zero; otherwise the attempt is repeated literally (EXACT).  The additions of +0 stay: they
          // turn a -0 into +0 like the reference's.
          unsigned int ex = 0u;
#pragma unroll
          for (int i = 0; i < D; i++) {
            const unsigned int e = (unsigned int)__double2hiint(F0[i]) & 0x7ff00000u;
            ex = ex > e ? ex : e;
          }
          bad = bad || (ex == 0x7ff00000u) || (((unsigned int)__double2hiint(h) & 0x7ff00000u) < 0x01000000u);
#pragma unroll
          for (int i = 0; i < D; i++) T[i] = 0.0;
        } else {
          double Fa[D];
          const Recip R100 = recip_prepare(100.0);
          const double h100 = div_sel<EXACT>(h, R100, bad);
          Rhs::eval(t + h100, y, Fa, p);  // :313
          const Recip Rh100 = recip_prepare(h100);
#pragma unroll
          for (int i = 0; i < D; i++) T[i] = div_sel<EXACT, true>(hd * (Fa[i] - F0[i]), Rh100, bad);
        }
#pragma unroll
        for (int i = 0; i < D; i++) k1[i] = F0[i] + T[i];  // :316
        W.template solve_t<EXACT>(k1, bad);
        const double hh = 0.5 * h;
#pragma unroll
        for (int i = 0; i < D; i++) tmp[i] = y[i] + hh * k1[i];  // :317
        Rhs::eval(t + 0.5 * h, tmp, F1, p);
#pragma unroll
        for (int i = 0; i < D; i++) k2[i] = F1[i] - k1[i];  // :318
        W.template solve_t<EXACT>(k2, bad);
#pragma unroll
        for (int i = 0; i < D; i++) k2[i] = k2[i] + k1[i];
#pragma unroll
        for (int i = 0; i < D; i++) ynew[i] = y[i] + h * k2[i];  // :319
        Rhs::eval(t + h, ynew, F2, p);                           // :320
#pragma unroll
        for (int i = 0; i < D; i++) k3[i] = ((F2[i] - e32 * (k2[i] - F1[i])) - 2 * (k1[i] - F0[i])) + T[i];  // :321
        W.template solve_t<EXACT>(k3, bad);
        if (ODEB200_S23_SPECJ) jac_rhs = eval_jacobian<Rhs, JAC, D, NPA, EXACT>(t + h, ynew, p, Jn, bad, F2);  // :352, ahead of time (F(t + h, ynew) = F2)
#pragma unroll
        for (int i = 0; i < D; i++) tmp[i] = (k1[i] - 2 * k2[i]) + k3[i];
        err = div_sel<EXACT>(fabs(h), recip_prepare(6.0), bad) * norm2_sel<EXACT, D>(tmp, bad);  // :323
        nyn = norm2_sel<EXACT, D>(ynew, bad);
        const double nyv = (ODEB200_S23_NORMY && !EXACT) ? ny : norm2_sel<EXACT, D>(y, bad);
        delta = jl_max(A.reltol * jl_max(nyv, nyn), A.abstol);  // :324
      };
      bool bad = false;
      if (ODEB200_S23_DEFER) {
        attempt(iconst<0>{}, bad);
#ifndef ODEB200_S23_NOEXACT
        if (bad) attempt(iconst<1>{}, bad);
#endif
      } else {
        attempt(iconst<1>{}, bad);
      }
      if (singular) {
        fin = ODE_B200_ST_SINGULAR;
      } else {
        nrhs += 3;
        const bool accept = err <= delta;  // :327
        if (MODE == 1 && idx == A.trace_traj) {
          if (ntrace < A.trace_cap) {
            A.trace[ntrace * 4 + 0] = t;
            A.trace[ntrace * 4 + 1] = h;
            A.trace[ntrace * 4 + 2] = err / delta;
            A.trace[ntrace * 4 + 3] = accept ? 1.0 : 0.0;
          }
          ntrace++;
        }
        if (accept) {
          nacc++;
          if (MODE == 1 && A.points != ODE_B200_POINTS_FINAL) {
            double* py = A.pts_y + (idx * A.pts_cap) * D;
            double* pt = A.pts_t + idx * A.pts_cap;
            for (int q = 0; q < nfixed; q++) {  // :332 tspan points in (t, t+h]
              const double toi = A.tspan[q];
              if (toi > t && toi <= t + h) {
                const double s = (toi - t) / h;  // :334
                const double den = 1 - 2 * dd;
                if (npts < A.pts_cap && A.pts_y) {
                  pt[npts] = toi;
#pragma unroll
                  for (int i = 0; i < D; i++)  // :338
                    py[(long long)npts * D + i] =
                        y[i] + h * (((k1[i] * s) * (1 - s)) / den + ((k2[i] * s) * (s - 2 * dd)) / den);
                }
                npts++;
                last_tout = toi;
              }
            }
            if (A.points == ODE_B200_POINTS_ALL && last_tout != t + h) {  // :341-345
              if (npts < A.pts_cap && A.pts_y) {
                pt[npts] = t + h;
#pragma unroll
                for (int i = 0; i < D; i++) py[(long long)npts * D + i] = ynew[i];
              }
              npts++;
              last_tout = t + h;
            }
          }
          t = t + h;  // :348
#pragma unroll
          for (int i = 0; i < D; i++) {
            y[i] = ynew[i];  // :349
            F0[i] = F2[i];   // :351
          }
          ny = nyn;
          if (!ODEB200_S23_SPECJ) jac_rhs = eval_jacobian<Rhs, JAC, D, NPA>(t, y, p, Jn, F0);  // :352
          nrhs += jac_rhs;
#pragma unroll
          for (int i = 0; i < D; i++) {
#pragma unroll
            for (int j = 0; j < D; j++) {
              if constexpr (JS) Jsh[i * D + j][threadIdx.x] = Jn[i][j];
              else Jreg[i][j] = Jn[i][j];
            }
          }
        } else {
          nrej++;
        }
        h = tdir * jl_min(A.maxstep, (fabs(h) * 0.8) * dm_pow(delta / err, 1.0 / 3));  // :357
      }
    }

    if (fin >= 0) {
      if (MODE == 1 && npts > A.pts_cap && A.pts_y) fin = ODE_B200_ST_OVERFLOW;
      if (A.t_final) A.t_final[idx] = t;
      if (A.y_final) {
#pragma unroll
        for (int d = 0; d < D; d++) A.y_final[idx * D + d] = y[d];
      }
      if (A.n_accept) A.n_accept[idx] = nacc;
      if (A.n_reject) A.n_reject[idx] = nrej;
      if (A.n_rhs) A.n_rhs[idx] = nrhs;
      if (A.status) A.status[idx] = fin;
      if (MODE == 1) {
        if (A.pts_n) A.pts_n[idx] = npts;
        if (idx == A.trace_traj && A.trace_n) *A.trace_n = ntrace;
      }
      idx = -1;
    }
  }
}

}  // namespace odeb200
