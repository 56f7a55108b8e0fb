// field.cuh -- one large system with a GENERAL right-hand side: F_i(t, y) may read any component of y.
//
// Replaces oderk_adapt (/root/reference src/runge_kutta.jl:232-369) on a dof = N state for a registered "field"
// functor  rhs(i, n, y, p, t) -> dy_i/dt  (a gather over the whole state: N-body sums, networks, non-local
// couplings, 2-D grids flattened by the caller ...).  The stencil path (large.cuh) fuses all stages of an attempt
// into one kernel because a stencil's inputs are local; a general gather needs every stage's argument
//   ytmp_s = y + sum_{ss<s} (dt*k_ss)*a[s,ss]                       (calc_next_k! :452-455)
// complete in memory before k_s = F(t + c_s*dt, ytmp_s) can be evaluated anywhere, so here an attempt is one
// kernel per stage (SURVEY.md section 8d's stage-per-kernel design):
//   field_first_kernel       ytmp_2 = y + (dt*k1)*a21
//   field_stage_kernel<s>    k_s = F(ytmp_s);  s < S: ytmp_{s+1} from y, k_1..k_s;
//                            s = S: ytrial, yerr, the scaled error and its double-double partials (:384-398,:430-435)
//   field_f1_kernel          non-FSAL tableaus: k_next = F(t+dt, ytrial) (:320)
// with the reference's arithmetic per component (no FMA, left-associated).  Zero tableau coefficients are
// dropped unless ctl->redo asks for the literal repeat of a non-finite attempt (then Inf*0 = NaN as in the
// reference).  The session, the control kernel (accept/reject, new dt, pointer-parity swap), hinit's scalar
// phases, the Hermite output kernel and the error reduction are those of the stencil path with zero ghost cells.
// HBM traffic for dopri5: 48 array passes = 384 B per component per attempt (the gather of ytmp counted once).
#pragma once

#include "large.cuh"

namespace odeb200 {

constexpr int FIELD_THREADS = 256;
constexpr int FIELD_MAX_STAGES = 13;

struct FieldArgs {
  double* Y[2];                   // y / ytrial by parity (ctl->par), like the stencil path; no ghost cells
  double* K[2];                   // k1 / k_next by parity
  double* KS[FIELD_MAX_STAGES];   // KS[s] = k_{s+1} for the middle stages s = 1 .. S-2 (0-based)
  double* T[2];                   // ytmp ping-pong
  LargeCtl* ctl;
  double* part;                   // per-CTA error partials [grid][4]
  double* xsend;                  // [XCHG]
  long long n;
  StencilCtx sc;                  // parameters (p), n_global = n, i0 = 0
};

// gather version of the reaction-diffusion stencil: the same expression as RhsReactionDiffusion, reading its
// neighbours from the global state (compiled in so that the templates below are part of the library build, and
// as the like-for-like comparison of the stage-per-kernel design with the fused kernel)
struct FieldReactionDiffusion {
  static constexpr int NP = 2, ID = ODE_B200_RHS_RD_FIELD;
  __device__ __forceinline__ static double rhs(long long i, long long n, const double* __restrict__ y, const double (&p)[8],
                                               double) {
    const double um = y[i == 0 ? n - 1 : i - 1], u = y[i], up = y[i == n - 1 ? 0 : i + 1];
    return p[0] * ((um - 2.0 * u) + up) + (p[1] * u) * (1.0 - u);
  }
};

// k of reference stage ss (1-based) at component i: k1 lives in K[par], the last stage in K[par^1], the rest in KS
template <int S>
__device__ __forceinline__ const double* stage_array(const FieldArgs& A, int par, int ss0) {  // ss0 = 0-based
  if (ss0 == 0) return par ? A.K[1] : A.K[0];
  if (ss0 == S - 1) return par ? A.K[0] : A.K[1];
  return A.KS[ss0];
}

// ytmp_2 = y + (dt*k1)*a21   (first RHS argument of the attempt)
template <class Tab>
__global__ void __launch_bounds__(FIELD_THREADS) field_first_kernel(const __grid_constant__ FieldArgs A) {
  const LargeCtl* ctl = A.ctl;
  if (ctl->done) return;
  const bool lit = ctl->redo != 0;
  const int par = ctl->par;
  const double dt = ctl->dt;
  const double* __restrict__ y = par ? A.Y[1] : A.Y[0];
  const double* __restrict__ k1 = par ? A.K[1] : A.K[0];
  double* __restrict__ out = A.T[0];
  constexpr double a = Tab::S > 1 ? Tab::a(Tab::S > 1 ? 1 : 0, 0) : 0.0;
  for (long long i = (long long)blockIdx.x * FIELD_THREADS + threadIdx.x; i < A.n; i += (long long)gridDim.x * FIELD_THREADS) {
    double v = y[i];
    if (a != 0.0 || lit) v = v + (dt * k1[i]) * a;
    out[i] = v;
  }
}

// Stage s (1-based reference numbering, 2 <= s <= S).
template <class Tab, class Rhs, int s>
__global__ void __launch_bounds__(FIELD_THREADS) field_stage_kernel(const __grid_constant__ FieldArgs A) {
  constexpr int S = Tab::S, T = FIELD_THREADS;
  LargeCtl* ctl = A.ctl;
  if (ctl->done) return;
  const bool lit = ctl->redo != 0;
  const int par = ctl->par;
  const double dt = ctl->dt, t = ctl->t;
  const double reltol = ctl->reltol, abstol = ctl->abstol;
  const double* __restrict__ y = par ? A.Y[1] : A.Y[0];
  const double* __restrict__ tin = A.T[s & 1];          // ytmp_s (stage 2 reads T[0])
  double* __restrict__ tout = A.T[(s + 1) & 1];         // ytmp_{s+1}
  double* __restrict__ ks_out = (s == S) ? (par ? A.K[0] : A.K[1]) : A.KS[s - 1];
  double* __restrict__ ytrial_out = par ? A.Y[0] : A.Y[1];
  constexpr double c = Tab::c(s - 1);
  const double tt = t + c * dt;  // :456

  dm_dd acc;
  acc.hi = 0.0;
  acc.lo = 0.0;
  double mx = 0.0;
  bool anynan = false;
  for (long long i = (long long)blockIdx.x * T + threadIdx.x; i < A.n; i += (long long)gridDim.x * T) {
    const double ks = Rhs::rhs(i, A.n, tin, A.sc.p, tt);
    if constexpr (s < S) {
      ks_out[i] = ks;
      double v = y[i];  // calc_next_k! :452-455 for stage s+1
      static_for<s>([&](auto J) {
        constexpr int ss = decltype(J)::value;  // 0-based earlier stage
        constexpr double a = Tab::a(s, ss);     // a[s+1, ss+1] in the reference's numbering
        if (a != 0.0 || lit) {
          const double kv = (ss == s - 1) ? ks : stage_array<S>(A, par, ss)[i];
          v = v + (dt * kv) * a;
        }
      });
      tout[i] = v;
    } else {
      // rk_embedded_step! :384-398 with every stage of this component, then the scaling of stepsize_hw92! :430-434
      constexpr double b10 = Tab::b(0, 0), b20 = Tab::b(1, 0);
      const double k1v = stage_array<S>(A, par, 0)[i];
      double ytr = lit ? (0.0 + b10 * k1v) : ((b10 != 0.0) ? b10 * k1v : 0.0);
      double yer = lit ? (0.0 + b20 * k1v) : ((b20 != 0.0) ? b20 * k1v : 0.0);
      static_for<S - 1>([&](auto J) {
        constexpr int ss = decltype(J)::value + 1;  // 0-based stage 1 .. S-1
        constexpr double b1 = Tab::b(0, ss), b2 = Tab::b(1, ss);
        if (b1 != 0.0 || b2 != 0.0 || lit) {
          const double kv = (ss == S - 1) ? ks : A.KS[ss][i];
          if (b1 != 0.0 || lit) ytr = ytr + b1 * kv;
          if (b2 != 0.0 || lit) yer = yer + b2 * kv;
        }
      });
      const double yi = y[i];
      yer = dt * (ytr - yer);
      ytr = yi + dt * ytr;
      ytrial_out[i] = ytr;
      ks_out[i] = ks;  // FSAL: this is k1 of the next step (:320); otherwise field_f1_kernel overwrites it
      anynan = anynan || (ytr != ytr);
      const double ay = fabs(yi), at = fabs(ytr);
      const double e = yer / (abstol + ((ay > at) ? ay : at) * reltol);  // :433
      acc = dd_add_sq(acc, e);
      mx = normInf_step(mx, fabs(e));
    }
  }
  if constexpr (s == S) reduce_error_partials<T>(acc, mx, anynan, A.part, ctl, A.xsend);
}

// non-FSAL tableaus: f1 = F(t + dt, ytrial) -> k_next (:320); evaluated for every attempt (used when it is accepted)
template <class Rhs>
__global__ void __launch_bounds__(FIELD_THREADS) field_f1_kernel(const __grid_constant__ FieldArgs A) {
  const LargeCtl* ctl = A.ctl;
  if (ctl->done) return;
  const int par = ctl->par;
  const double tt = ctl->t + ctl->dt;
  const double* __restrict__ ytrial = par ? A.Y[0] : A.Y[1];
  double* __restrict__ kn = par ? A.K[0] : A.K[1];
  for (long long i = (long long)blockIdx.x * FIELD_THREADS + threadIdx.x; i < A.n; i += (long long)gridDim.x * FIELD_THREADS)
    kn[i] = Rhs::rhs(i, A.n, ytrial, A.sc.p, tt);
}

// ---- fixed-step explicit Runge-Kutta on a dof = N state (oderk_fixed, src/runge_kutta.jl:176-212) ---------------
// A stencil functor seen as a field functor (periodic gather of its window), so that the stage-per-kernel
// drivers below also serve registered stencils.
template <class St>
struct FieldFromStencil {
  static constexpr int NP = St::NP, ID = St::ID;
  __device__ __forceinline__ static double rhs(long long i, long long n, const double* __restrict__ y, const double (&p)[8],
                                               double t) {
    constexpr int R = St::RADIUS;
    double w[2 * R + 1];
#pragma unroll
    for (int k = 0; k < 2 * R + 1; k++) {
      long long j = i + (k - R);
      if (j < 0) j += n;
      else if (j >= n) j -= n;
      w[k] = y[j];
    }
    return St::point(w, p, t, i, n);
  }
};

struct FixedArgs {
  const double* y;               // state at tspan[i]
  double* ynew;                  // state at tspan[i+1]
  double* KS[FIELD_MAX_STAGES];  // KS[s] = k_{s+1}, s = 0 .. S-2
  double* T[2];                  // stage-argument ping-pong
  long long n;
  double t, dt;                  // tspan[i], tspan[i+1] - tspan[i] (:202)
  StencilCtx sc;
};

// Stage s (1-based, 1 <= s <= S): k_s = F(t + c_s*dt, ytmp_s) -- every stage calls F, the first included (:205) --
// then ytmp_{s+1} (every coefficient multiplied, zeros too, like the reference), or for s = S the update
// ys[i+1] = y + sum (dt*b[s])*k_s (:207).
template <class Tab, class Rhs, int s>
__global__ void __launch_bounds__(FIELD_THREADS) field_fixed_stage_kernel(const __grid_constant__ FixedArgs A) {
  constexpr int S = Tab::S;
  const double dt = A.dt;
  const double* __restrict__ tin = (s == 1) ? A.y : A.T[s & 1];
  double* __restrict__ tout = A.T[(s + 1) & 1];
  constexpr double c = Tab::c(s - 1);
  const double tt = A.t + c * dt;  // :456
  for (long long i = (long long)blockIdx.x * FIELD_THREADS + threadIdx.x; i < A.n; i += (long long)gridDim.x * FIELD_THREADS) {
    const double ks = Rhs::rhs(i, A.n, tin, A.sc.p, tt);
    if constexpr (s < S) {
      A.KS[s - 1][i] = ks;
      double v = A.y[i];
      static_for<s>([&](auto J) {
        constexpr int ss = decltype(J)::value;
        constexpr double a = Tab::a(s, ss);
        const double kv = (ss == s - 1) ? ks : A.KS[ss][i];
        v = v + (dt * kv) * a;  // (dt*k)*a :454
      });
      tout[i] = v;
    } else {
      double yn = A.y[i];  // :203
      static_for<S>([&](auto J) {
        constexpr int ss = decltype(J)::value;
        constexpr double b = Tab::b(0, ss);
        const double kv = (ss == S - 1) ? ks : A.KS[ss][i];
        yn = yn + (dt * b) * kv;  // :207
      });
      A.ynew[i] = yn;
    }
  }
}

// ---- ode_ms: Adams-Bashforth, order ramping up to `order` <= 5 (src/ODE.jl:175-213) on a dof = N state ----------
struct MsArgs {
  const double* y;        // x[i]
  double* ynew;           // x[i+1]
  const double* hold[ODE_B200_MS_MAX_ORDER - 1];  // xdot[i-4] .. xdot[i-1] (oldest first; unused ones may be anything)
  double* hnew;           // xdot[i] = F(t_i, x[i]) is written here
  long long n;
  double t, h;            // tspan[i], tspan[i+1] - tspan[i]
  int so;                 // steporder = min(i, order), 1-based step number i (:200)
  double hb[ODE_B200_MS_MAX_ORDER];  // h[i]*b[steporder, j] for hold[0..3], hnew (zero where unused): (h*b)*xdot (:205)
  StencilCtx sc;
};
template <class Rhs>
__global__ void __launch_bounds__(FIELD_THREADS) field_ms_kernel(const __grid_constant__ MsArgs A) {
  constexpr int MO = ODE_B200_MS_MAX_ORDER;
  const int so = A.so;
  for (long long i = (long long)blockIdx.x * FIELD_THREADS + threadIdx.x; i < A.n; i += (long long)gridDim.x * FIELD_THREADS) {
    const double hn = Rhs::rhs(i, A.n, A.y, A.sc.p, A.t);  // :201
    A.hnew[i] = hn;
    double v = A.y[i];  // :203-206, oldest derivative first
#pragma unroll
    for (int q = 0; q < MO - 1; q++)
      if (q >= MO - so) v = v + A.hb[q] * A.hold[q][i];
    v = v + A.hb[MO - 1] * hn;
    A.ynew[i] = v;
  }
}

// ---- hinit (src/ODE.jl:85-112) ---------------------------------------------------------------------
// f0 = F(t0, y) -> K[0]; partial max|y|, max|f0|
template <class Rhs>
__global__ void __launch_bounds__(256) field_hinit_f0_kernel(const double* __restrict__ y, double* f0, const LargeCtl* ctl,
                                                             double* part, long long n, const __grid_constant__ StencilCtx sc) {
  __shared__ double sh[8];
  const double t0 = ctl->tstart;
  double my = 0.0, mf = 0.0;
  for (long long i = blockIdx.x * 256LL + threadIdx.x; i < n; i += (long long)gridDim.x * 256) {
    const double f = Rhs::rhs(i, n, y, sc.p, t0);
    f0[i] = f;
    my = normInf_step(my, fabs(y[i]));
    mf = normInf_step(mf, fabs(f));
  }
  double a = block_max_nanprop<256>(my, sh);
  double b = block_max_nanprop<256>(mf, sh);
  if (threadIdx.x == 0) {
    part[blockIdx.x * 2] = a;
    part[blockIdx.x * 2 + 1] = b;
  }
}
// x1 = y + (tdir*h0)*f0 (:100), materialised because F gathers
__global__ void __launch_bounds__(256) field_hinit_x1_kernel(const double* __restrict__ y, const double* __restrict__ f0,
                                                             const LargeCtl* ctl, double* x1, long long n);
// f1 = F(t0 + tdir*h0, x1); partial max|f1 - f0|
template <class Rhs>
__global__ void __launch_bounds__(256) field_hinit_f1_kernel(const double* __restrict__ x1, const double* __restrict__ f0,
                                                             const LargeCtl* ctl, double* part, long long n,
                                                             const __grid_constant__ StencilCtx sc) {
  __shared__ double sh[8];
  const double t1 = ctl->tstart + ctl->tdir * ctl->h0;  // :101
  double md = 0.0;
  for (long long i = blockIdx.x * 256LL + threadIdx.x; i < n; i += (long long)gridDim.x * 256) {
    const double f1 = Rhs::rhs(i, n, x1, sc.p, t1);
    md = normInf_step(md, fabs(f1 - f0[i]));
  }
  double a = block_max_nanprop<256>(md, sh);
  if (threadIdx.x == 0) part[blockIdx.x * 2] = a;
}

}  // namespace odeb200
